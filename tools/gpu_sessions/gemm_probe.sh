#!/bin/bash
# GPU bring-up driver for tools/gemm_probe.py: every section in its own process (a trapped kernel
# poisons the CUDA context) and under its own timeout.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 300 python tools/gemm_probe.py nt
for g in "0 0 0" "1024 8192 2048" "8192 1024 4096" "1024 2048 1024" "2048 1024 2048" "8192 1024 32" "1024 8192 32"; do
  timeout 120 python tools/gemm_probe.py geom $g
done
timeout 300 python tools/gemm_probe.py mn 0 0 0
timeout 300 python tools/gemm_probe.py time_nt
timeout 300 python tools/gemm_probe.py time_mn 0 0 0
} > gpurun_out/gemm_probe.log 2>&1
tail -150 gpurun_out/gemm_probe.log
