#!/bin/bash
# round 2, session B: ncu --set full of the pair kernel, single-wave launch with the dense-output epilogue (bias + dropout + LN residual)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/r2b_ncu_pair.log
: > $L
for v in resid plain; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_pair192 -s 9 -c 1 -f -o gpurun_out/r2b_pair_$v python tools/gemm_trace.py 6464 768 768 $v >> $L 2>&1
done
ls -la gpurun_out >> $L
tail -30 $L
