#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/gpu_tune.log
: > $L
run() { echo "=== $*" >> $L; timeout "$1" "${@:2}" >> $L 2>&1; echo "--- exit $?" >> $L; }
run 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "gemm or layernorm"
run 600 python -m pytest tests/test_gpu_modules.py -q -x
for bn in 0 128 192 256; do
  echo "=== VLB_FORCE_BN=$bn" >> $L
  VLB_FORCE_BN=$bn timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --skip-e2e >> $L 2>&1
done
run 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --skip-e2e
grep -E "^===|^---|passed|failed|rror" $L | head -30
python - <<'PY'
import json,re
for line in open('gpurun_out/gpu_tune.log'):
    if line.startswith('=== VLB_FORCE_BN'): print(line.strip())
    if line.startswith('{"metric"'):
        d=json.loads(line)
        if d['ms_per_step']<1000:
            print('  ms/step %.3f  value %.0f  gemm frac %.3f  ' % (d['ms_per_step'], d['value'], d['roofline']['frac']), {k:round(v['ms_per_step'],3) for k,v in d['kernel_profile'].items()})
PY
python tools/summarize_launches.py gpurun_out/launches.csv | tee gpurun_out/launches_summary.txt | head -24
