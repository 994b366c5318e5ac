#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
# full-set capture (with source correlation) of the GEMM launches of one forward+backward layer in situ
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel|gemm_grouped" -s 96 -c 10 -f -o gpurun_out/prof_gemm_v4 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --skip-e2e --no-graph > gpurun_out/ncu_gemm_v4.log 2>&1
tail -3 gpurun_out/ncu_gemm_v4.log | cut -c1-300
