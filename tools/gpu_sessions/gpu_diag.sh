#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/gpu_diag.log
: > $L
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "cta_pair" >> $L 2>&1
for bn in 256 1256 1128; do
  timeout 120 python tools/gemm_one.py 0 6464 2304 768 $bn 7 >> $L 2>&1
  timeout 120 python tools/gemm_one.py 0 6464 2304 6144 $bn 7 >> $L 2>&1
  timeout 120 python tools/gemm_one.py 2 3072 768 6464 $bn 7 >> $L 2>&1
  timeout 120 python tools/gemm_one.py 1 6464 768 3072 $bn 7 >> $L 2>&1
done
cat $L | grep -E "mode|rror|passed|failed"
