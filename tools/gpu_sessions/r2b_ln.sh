#!/bin/bash
# round 2, session B: lean LayerNorm forward (one warp per row, 48 registers, whole matrix resident)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/r2b_ln.log
: > $L
echo "=== pytest layernorm + modules + parity" >> $L
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_modules.py tests/test_gpu_parity_bf16.py tests/test_gpu_dropout.py -q -x -k "not gemm" 2>&1 | tail -4 >> $L
for lean in 0 1; do
echo "=== bench VLB_LN_LEAN=$lean" >> $L
VLB_LN_LEAN=$lean timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager --no-other-configs 2>> $L | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms/step %.3f value %.0f e2e %.0f gemm frac %.3f'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac']),{k:round(v['ms_per_step'],3) for k,v in d['kernel_profile'].items()})
" >> $L
done
grep -v Warn $L | tail -12
