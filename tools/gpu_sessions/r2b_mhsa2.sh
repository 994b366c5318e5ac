#!/bin/bash
# round 2, session B: attention kernels with trimmed key columns + packed fp32 softmax (forward and backward)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/r2b_mhsa2.log
: > $L
echo "=== pytest mhsa + dropout + modules + parity" >> $L
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_dropout.py tests/test_gpu_modules.py tests/test_gpu_parity_bf16.py tests/test_gpu_training_loop.py -q -x 2>&1 | tail -8 >> $L
echo "=== mhsa timing" >> $L
timeout 300 python tools/mhsa_one.py 64 101 12 --drop 0 >> $L 2>&1
timeout 300 python tools/mhsa_one.py 64 101 12 >> $L 2>&1
timeout 300 python tools/mhsa_one.py 64 121 12 --drop 0 >> $L 2>&1
timeout 300 python tools/mhsa_one.py 64 165 16 --drop 0 >> $L 2>&1
echo "=== bench" >> $L
timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager --no-other-configs > gpurun_out/r2b_bench_mhsa2.json 2>> $L; echo "--- exit $?" >> $L
grep -v Warn $L | tail -24
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2b_bench_mhsa2.json').read().strip().splitlines()[-1])
print('ms/step %.3f value %.0f e2e %.0f gemm frac %.3f'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac']),{k:round(v['ms_per_step'],3) for k,v in d['kernel_profile'].items()})
PY
