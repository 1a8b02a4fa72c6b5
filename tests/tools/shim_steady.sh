#!/bin/bash
# steady-state step rate of the unmodified reference driver linked against the shim, memo on / off: iterations over
# the span the shim reports (HB_SHIM_STATS: end of its first device batch to end of its last; start-up excluded)
cd oracle/_ref
IT=${1:-3000}
for memo in 1 0 1 0; do
  HB_SHIM_MEMO=$memo HB_SHIM_STATS=1 ./hb_mcmc_ref_shim $IT 102289966 0.7960497 9 > scratch/a.log 2>&1
  span=$(grep -o "span [0-9.]*" scratch/a.log | cut -d' ' -f2)
  python3 -c "print('memo', $memo, ':', round($IT/$span,1), 'steps/s over', $span, 's')"
  grep "loglikelihood calls" scratch/a.log | cut -c1-220
done
