#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/gpu_round8.log
: > $L
run() { echo "=== $*" >> $L; timeout "$1" "${@:2}" >> $L 2>&1; echo "--- exit $?" >> $L; }
run 400 python -m pytest tests/test_gpu_kernels.py -q -x -k "cta_pair"
for bn in 256 2256 128 2128; do
  timeout 120 python tools/gemm_one.py 0 6464 2304 768 $bn 7 >> $L 2>&1
  timeout 120 python tools/gemm_one.py 0 6464 2304 6144 $bn 7 >> $L 2>&1
  timeout 120 python tools/gemm_one.py 1 6464 768 3072 $bn 7 >> $L 2>&1
done
echo "=== VLB_MC2=0" >> $L; VLB_MC2=0 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --skip-e2e >> $L 2>&1
echo "=== VLB_MC2=1" >> $L; VLB_MC2=1 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --skip-e2e >> $L 2>&1
grep -E "^===|^---|passed|failed|rror|mode " $L | head -40
python - <<'PY'
import json
for line in open('gpurun_out/gpu_round8.log'):
    if line.startswith('==='): hdr=line.strip()
    if line.startswith('{"metric"'):
        d=json.loads(line)
        if d['ms_per_step']<1000:
            print(hdr[:90]); print('  ms/step %.3f value %.0f graph %s gemm frac %.3f ' % (d['ms_per_step'], d['value'], d.get('cuda_graph'), d['roofline']['frac']), {k:round(v['ms_per_step'],3) for k,v in d['kernel_profile'].items()})
PY
