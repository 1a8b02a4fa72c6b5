cd oracle/_ref
for b in ${BINS:-hb_mcmc_ref hb_mcmc_ref_shim hb_mcmc_ref_shim25}; do
  s=$(date +%s.%N); ./$b ${1:-2000} 102289966 0.7960497 9 > scratch/run_$b.log 2>&1; e=$(date +%s.%N)
  python3 -c "print('$b', ${1:-2000}, 'iterations', round($e-$s,2), 's ->', round(${1:-2000}/($e-$s),1), 'steps/s')"
  grep -E "logL=" scratch/run_$b.log | tail -1 | cut -c1-160
done
