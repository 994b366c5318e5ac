#!/bin/bash
# round 2, session B: the complete default bench line (all configs, eager baseline, CPU arm) + the reference arm
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/r2b_full_bench.log
: > $L
echo "=== pytest gemm" >> $L
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x -k gemm 2>&1 | tail -3 >> $L
echo "=== bench (default)" >> $L
timeout 900 python bench.py > gpurun_out/r2b_bench_full.json 2>> $L; echo "--- exit $?" >> $L
echo "=== bench --impl reference" >> $L
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2b_bench_reference.json 2>> $L; echo "--- exit $?" >> $L
grep -v Warn $L | tail -20
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2b_bench_full.json').read().strip().splitlines()[-1])
print('ms/step %.3f value %.0f e2e %.0f gemm frac %.3f'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac']),{k:round(v['ms_per_step'],3) for k,v in d['kernel_profile'].items()})
for k,v in d.get('other_configs',{}).items(): print(k, {kk:(round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','ms_per_step','unit')}, v.get('roofline',{}).get('frac'))
print(d.get('gpu_eager_baseline'))
print(d.get('cpu_baseline'))
PY
