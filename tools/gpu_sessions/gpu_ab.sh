#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/gpu_ab.log
: > $L
for i in 1 2; do
echo "=== old tree (a3cac68) run $i" >> $L; (cd tools/_old_tree && timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --skip-e2e) >> $L 2>&1
echo "=== new tree run $i" >> $L; timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --skip-e2e >> $L 2>&1
done
python - <<'PY'
import json
for line in open('gpurun_out/gpu_ab.log'):
    if line.startswith('==='): hdr=line.strip()
    if line.startswith('{"metric"'):
        d=json.loads(line)
        print(hdr[:90]); print('  ms/step %.3f value %.0f gemm frac %.3f ' % (d['ms_per_step'], d['value'], d['roofline']['frac']), {k:round(v['ms_per_step'],3) for k,v in d['kernel_profile'].items()})
PY
grep -i "error\|Traceback" $L | head -5
