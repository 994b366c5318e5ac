#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/gpu_cg2.log
: > $L
run() { echo "=== $*" >> $L; timeout "$1" "${@:2}" >> $L 2>&1; echo "--- exit $?" >> $L; }
run 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "cta_pair"
run 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "gemm and not cta_pair"
run 600 python -m pytest tests/test_gpu_modules.py -q -x
for cg in 0 1 -1; do
  echo "=== VLB_CG2=$cg" >> $L
  VLB_CG2=$cg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --skip-e2e >> $L 2>&1
done
VLB_CG2=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cg2.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --skip-e2e >> $L 2>&1
grep -E "^===|^---|passed|failed|rror" $L | head -30
python - <<'PY'
import json
for line in open('gpurun_out/gpu_cg2.log'):
    if line.startswith('=== VLB_CG2'): print(line.strip())
    if line.startswith('{"metric"'):
        d=json.loads(line)
        if d['ms_per_step']<1000:
            print('  ms/step %.3f  value %.0f  gemm frac %.3f  ' % (d['ms_per_step'], d['value'], d['roofline']['frac']), {k:round(v['ms_per_step'],3) for k,v in d['kernel_profile'].items()})
PY
python tools/summarize_launches.py gpurun_out/launches_cg2.csv | head -24
