#!/bin/bash
# usage: build_variants.sh name "-DHB_EVAL_THREADS=512 ..." [extra nvcc flags]
cd /root/repo/hb_mcmc_b200/csrc
name=$1; shift
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -Xptxas -v "$@" -o /root/repo/tools/variants/lib_$name.so hb_kernels.cu hb_capi.cu hb_pt.cu hb_gaia_pt.cu hb_comm.cu -ldl 2>&1 | grep -A2 "k_chain_evalILi" | grep -E "Used|spill" | sed "s/^/[$name] /"
