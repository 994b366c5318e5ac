#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/gpu_round7.log
: > $L
run() { echo "=== $*" >> $L; timeout "$1" "${@:2}" >> $L 2>&1; echo "--- exit $?" >> $L; }
run 1200 python -m pytest tests/ -x -q -m gpu
run 300 python -c "import __graft_entry__ as g; g.smoke()"
echo "=== VLB_PDL=0 graph" >> $L; VLB_PDL=0 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --skip-e2e >> $L 2>&1
echo "=== VLB_PDL=1 graph" >> $L; VLB_PDL=1 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --skip-e2e >> $L 2>&1
echo "=== VLB_PDL=1 eager" >> $L; VLB_PDL=1 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --skip-e2e --no-graph >> $L 2>&1
run 600 python bench.py
grep -E "^===|^---|passed|failed|rror|smoke" $L | head -40
python - <<'PY'
import json
for line in open('gpurun_out/gpu_round7.log'):
    if line.startswith('==='): hdr=line.strip()
    if line.startswith('{"metric"'):
        d=json.loads(line)
        if d['ms_per_step']<1000:
            print(hdr[:90]); print('  ms/step %.3f value %.0f e2e %.0f graph %s gemm frac %.3f ' % (d['ms_per_step'], d['value'], d['e2e']['value'], d.get('cuda_graph'), d['roofline']['frac']), {k:round(v['ms_per_step'],3) for k,v in d['kernel_profile'].items()}, d.get('cpu_baseline'), d.get('clocks'))
PY
