#!/bin/bash
# round 2, session B: what each fused epilogue costs (pair kernel, single wave), packed-fp32 GELU check
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/r2b_epi_probe.log
: > $L
echo "=== pytest gemm epilogues" >> $L
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "epilogues or pair192" >> $L 2>&1; echo "--- exit $?" >> $L
echo "=== probe 6464 768 768" >> $L
timeout 300 python tools/epilogue_cost_probe.py 6464 768 768 >> $L 2>&1
echo "=== probe 6464 768 768 PAIR192=0" >> $L
VLB_PAIR192=0 timeout 300 python tools/epilogue_cost_probe.py 6464 768 768 >> $L 2>&1
echo "=== probe 6464 3072 768 (old kernel: 4.13 rounds)" >> $L
timeout 300 python tools/epilogue_cost_probe.py 6464 3072 768 >> $L 2>&1
echo "=== probe 6464 3072 768 PAIR192=2" >> $L
VLB_PAIR192=2 timeout 300 python tools/epilogue_cost_probe.py 6464 3072 768 >> $L 2>&1
cat $L | grep -v Warning | tail -70
