#!/bin/bash
# round 2, session B: GELU launch with 16 epilogue warps on 128x192 tiles
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/r2b_ew16.log
: > $L
echo "=== pytest gelu wide + epilogues" >> $L
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "gelu_wide or epilogues" 2>&1 | tail -3 >> $L
for e in 0 1; do
echo "=== layer gemm bench VLB_GELU_EW16=$e" >> $L
VLB_GELU_EW16=$e VLB_BENCH_CUBLAS=0 timeout 300 python tools/layer_gemm_bench.py 2>&1 | grep -E "ffn-up|sum" >> $L
echo "=== bench VLB_GELU_EW16=$e" >> $L
VLB_GELU_EW16=$e timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager --no-other-configs 2>> $L | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms/step %.3f value %.0f e2e %.0f gemm frac %.3f'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac']),{k:round(v['ms_per_step'],3) for k,v in d['kernel_profile'].items()})
" >> $L
done
VLB_GELU_EW16=1 VLB_PAIR192=0 timeout 120 python tools/gemm_trace.py 6464 3072 768 gelu --persistent >> $L 2>&1
grep -v Warn $L | tail -24
