#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/gpu_final.log
: > $L
echo "=== ncu full: implicit conv GEMMs (fprop / wgrad / dgrad of the res5 3x3)" >> $L
timeout 110 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel --launch-skip 3 -c 3 -f -o gpurun_out/r01_ncu_full_conv_gemm python tools/conv_one.py 288 2 >> $L 2>&1
echo "--- exit $?" >> $L
echo "=== pytest -m gpu (frontend + kernels subset)" >> $L
timeout 150 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_frontend.py -q -x -m gpu -k "gemm or conv or bottleneck or roi" >> $L 2>&1
echo "--- exit $?" >> $L
tail -5 $L
