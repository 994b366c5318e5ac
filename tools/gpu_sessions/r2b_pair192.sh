#!/bin/bash
# round 2, session B: bring-up of the CTA-pair kernel with 192-row tiles (csrc/gemm_pair192.cuh) + state of the suite / bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/r2b_pair192.log
: > $L
echo "=== pytest pair192" >> $L
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "pair192" >> $L 2>&1; rc=$?; echo "--- exit $rc" >> $L
PAIR=1; [ $rc -ne 0 ] && PAIR=0
for mode in 0 $PAIR 2; do
  echo "=== layer gemm bench VLB_PAIR192=$mode" >> $L
  VLB_PAIR192=$mode VLB_BENCH_CUBLAS=$([ $mode = 0 ] && echo 1 || echo 0) timeout 300 python tools/layer_gemm_bench.py >> $L 2>&1
done
echo "=== layer gemm bench VLB_PAIR192=1 NN_BN=256" >> $L
VLB_PAIR192=1 VLB_PAIR192_NN_BN=256 VLB_BENCH_CUBLAS=0 timeout 300 python tools/layer_gemm_bench.py >> $L 2>&1
echo "=== pytest -m gpu (VLB_PAIR192=$PAIR)" >> $L
VLB_PAIR192=$PAIR timeout 1500 python -m pytest tests -m gpu -x -q >> $L 2>&1; echo "--- exit $?" >> $L
for mode in 0 $PAIR; do
  echo "=== bench VLB_PAIR192=$mode" >> $L
  VLB_PAIR192=$mode timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager --no-other-configs > gpurun_out/r2b_bench_pair$mode.json 2>> $L; echo "--- exit $?" >> $L
done
grep -E "^===|^---|passed|failed|rror|us  " $L | head -120
python - <<'PY'
import json
for m in (0,1):
    try:
        d=json.loads(open('gpurun_out/r2b_bench_pair%d.json'%m).read().strip().splitlines()[-1])
        print(m,'ms/step %.3f value %.0f e2e %.0f gemm frac %.3f'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac']),{k:round(v['ms_per_step'],3) for k,v in d['kernel_profile'].items()})
    except Exception as e: print(m,'no bench',e)
PY
