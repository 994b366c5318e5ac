#!/bin/bash
# round 2, session B: explicit shared-space accesses in the epilogue transposition (was generic LD.E / ST.E)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/r2b_lds.log
: > $L
echo "=== pytest gemm + dropout + modules" >> $L
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_dropout.py tests/test_gpu_modules.py -q -x 2>&1 | tail -3 >> $L
echo "=== layer gemm bench" >> $L
VLB_BENCH_CUBLAS=0 timeout 300 python tools/layer_gemm_bench.py 2>&1 | grep -v '^{' >> $L
echo "=== bench" >> $L
timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager --no-other-configs > gpurun_out/r2b_bench_lds.json 2>> $L; echo "--- exit $?" >> $L
grep -v Warn $L | tail -20
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2b_bench_lds.json').read().strip().splitlines()[-1])
print('ms/step %.3f value %.0f e2e %.0f gemm frac %.3f'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac']),{k:round(v['ms_per_step'],3) for k,v in d['kernel_profile'].items()})
PY
