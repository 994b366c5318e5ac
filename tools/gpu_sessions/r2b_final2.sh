#!/bin/bash
# round 2, session B: final validation -- new pre-training-loss workload, full suite, the driver's default bench line
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/r2b_final2.log
: > $L
echo "=== bench --config 6" >> $L
timeout 300 python bench.py --config 6 --no-cpu-baseline --no-gpu-eager --no-other-configs 2>> $L | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('config6: ms/step %.3f value %.0f e2e %.0f gemm frac %.3f last_loss %s'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac'],d['e2e'].get('last_loss')))
" >> $L
echo "=== pytest -m gpu" >> $L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $L
echo "=== bench (default)" >> $L
timeout 900 python bench.py > gpurun_out/r2b_bench_final.json 2>> $L; echo "--- exit $?" >> $L
grep -v Warn $L | tail -14
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2b_bench_final.json').read().strip().splitlines()[-1])
print('ms/step %.3f value %.0f e2e %.0f gemm frac %.3f'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac']),{k:round(v['ms_per_step'],3) for k,v in d['kernel_profile'].items()})
for k,v in d.get('other_configs',{}).items(): print(k, {kk:(round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','ms_per_step','unit','error')}, v.get('roofline',{}).get('frac'))
print(d.get('gpu_eager_baseline',{}).get('bf16_autocast'), d.get('cpu_baseline',{}).get('value'), d.get('clocks'))
PY
