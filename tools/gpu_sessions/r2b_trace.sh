#!/bin/bash
# round 2, session B: per-CTA timeline of the pair kernel (where the non-mainloop time of a single-wave launch goes)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/r2b_trace.log
: > $L
for v in plain bias resid gelu dgelu; do timeout 120 python tools/gemm_trace.py 6464 768 768 $v >> $L 2>&1; done
timeout 120 python tools/gemm_trace.py 6464 768 3072 plain >> $L 2>&1
timeout 120 python tools/gemm_trace.py 6464 768 3072 resid >> $L 2>&1
VLB_PAIR192=2 timeout 120 python tools/gemm_trace.py 6464 3072 768 gelu >> $L 2>&1
cat $L | grep -v Warn
