#!/bin/bash
# round 2, session B: per-item timeline of the persistent 128x256 kernel on the multi-round GEMMs (mainloop vs epilogue per round)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/r2b_trace2.log
: > $L
for v in bias gelu dgelu; do VLB_PAIR192=0 timeout 120 python tools/gemm_trace.py 6464 3072 768 $v --persistent >> $L 2>&1; done
VLB_PAIR192=0 timeout 120 python tools/gemm_trace.py 6464 2304 768 bias --persistent >> $L 2>&1
VLB_PAIR192=0 VLB_CG2=1 timeout 120 python tools/gemm_trace.py 6464 3072 768 bias --persistent >> $L 2>&1
VLB_PAIR192=0 VLB_CG2=1 timeout 120 python tools/gemm_trace.py 6464 3072 768 gelu --persistent >> $L 2>&1
grep -v Warn $L
