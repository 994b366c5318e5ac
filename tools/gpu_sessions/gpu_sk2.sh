#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/gpu_sk2.log
: > $L
run() { echo "=== $*" >> $L; timeout "$1" "${@:2}" >> $L 2>&1; echo "--- exit $?" >> $L; }
run 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "streamk or epilogues"
echo "=== VLB_SK_MIN_KB=4 pytest gemm" >> $L; VLB_SK_MIN_KB=4 timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x -k "gemm" >> $L 2>&1; echo "--- exit $?" >> $L
for cfg in "VLB_STREAMK=0" "VLB_STREAMK=1" "VLB_STREAMK=1 VLB_SK_MIN_KB=8"; do
  echo "=== $cfg" >> $L
  for shape in "0 6464 768 3072 0" "1 6464 768 3072 0" "1 6464 768 2304 0" "0 6464 768 768 0" "0 6464 3072 768 0" "1 6464 3072 768 0"; do
    env $cfg timeout 120 python tools/gemm_one.py $shape 9 >> $L 2>&1
  done
done
echo "=== bench VLB_STREAMK=0" >> $L; VLB_STREAMK=0 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --skip-e2e >> $L 2>&1
echo "=== bench VLB_STREAMK=1" >> $L; VLB_STREAMK=1 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --skip-e2e >> $L 2>&1
grep -E "^===|^---|passed|failed|rror|mode " $L | head -60
python - <<'PY'
import json
for line in open('gpurun_out/gpu_sk2.log'):
    if line.startswith('==='): hdr=line.strip()
    if line.startswith('{"metric"'):
        d=json.loads(line)
        print(hdr[:90]); print('  ms/step %.3f value %.0f graph %s gemm frac %.3f ' % (d['ms_per_step'], d['value'], d.get('cuda_graph'), d['roofline']['frac']), {k:round(v['ms_per_step'],3) for k,v in d['kernel_profile'].items()})
PY
