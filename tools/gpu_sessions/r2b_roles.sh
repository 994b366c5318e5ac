#!/bin/bash
# round 2, session B: TMA producer / MMA issuer as the highest warps of the CTA (scheduler priority) -- A/B
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/r2b_roles.log
: > $L
echo "=== pytest gemm" >> $L
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x -k "gemm" 2>&1 | tail -2 >> $L
for r in 0 1; do
echo "=== layer gemm bench VLB_ROLES_HIGH=$r" >> $L
VLB_ROLES_HIGH=$r VLB_BENCH_CUBLAS=0 timeout 300 python tools/layer_gemm_bench.py 2>&1 | grep -v '^{' >> $L
done
echo "=== layer gemm bench VLB_ROLES_HIGH=1 VLB_GELU_EW16=1" >> $L
VLB_ROLES_HIGH=1 VLB_GELU_EW16=1 VLB_BENCH_CUBLAS=0 timeout 300 python tools/layer_gemm_bench.py 2>&1 | grep -E "ffn-up|sum" >> $L
for r in 0 1; do
echo "=== bench VLB_ROLES_HIGH=$r" >> $L
VLB_ROLES_HIGH=$r timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager --no-other-configs 2>> $L | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms/step %.3f value %.0f e2e %.0f gemm frac %.3f'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac']),{k:round(v['ms_per_step'],3) for k,v in d['kernel_profile'].items()})
" >> $L
done
grep -v Warn $L | tail -40
