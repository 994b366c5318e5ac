#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/gpu_round3.log
: > $L
run() { echo "=== $*" >> $L; timeout "$1" "${@:2}" >> $L 2>&1; echo "--- exit $?" >> $L; }
run 600 python -m pytest tests/test_gpu_kernels.py -q -x
run 600 python -m pytest tests/test_gpu_modules.py -q -x
run 300 python -c "import __graft_entry__ as g; g.smoke()"
run 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-graph
run 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
for v in 1 3; do
  echo "=== VLB_LN_BWD_BLOCKS_PER_SM=$v" >> $L
  VLB_LN_BWD_BLOCKS_PER_SM=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --skip-e2e --no-graph >> $L 2>&1
done
run 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --skip-e2e --no-graph
grep -E "^===|^---|passed|failed|rror|smoke" $L | head -40
python - <<'PY'
import json
for line in open('gpurun_out/gpu_round3.log'):
    if line.startswith('==='): hdr=line.strip()
    if line.startswith('{"metric"'):
        d=json.loads(line)
        if d['ms_per_step']<1000:
            print(hdr[:60]); print('  ms/step %.3f value %.0f e2e %.0f graph %s gemm frac %.3f ' % (d['ms_per_step'], d['value'], d['e2e']['value'], d.get('cuda_graph'), d['roofline']['frac']), {k:round(v['ms_per_step'],3) for k,v in d['kernel_profile'].items()})
PY
python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt; head -22 gpurun_out/launches_summary.txt
