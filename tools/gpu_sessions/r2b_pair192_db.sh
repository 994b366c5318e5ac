#!/bin/bash
# round 2, session B: pair kernel with double-buffered acc1 (BN = 192) on multi-wave shapes, epilogue L2 prefetch A/B
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/r2b_pair192_db.log
: > $L
echo "=== pytest pair192" >> $L
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "pair192" >> $L 2>&1; rc=$?; echo "--- exit $rc" >> $L
run() { echo "=== layer gemm bench $*" >> $L; env "$@" VLB_BENCH_CUBLAS=0 timeout 300 python tools/layer_gemm_bench.py --config ${CFG:-2} 2>&1 | grep -v '^{' >> $L; }
run VLB_PAIR192=1 VLB_EPI_PREFETCH=0
run VLB_PAIR192=1 VLB_EPI_PREFETCH=1
run VLB_PAIR192=2 VLB_EPI_PREFETCH=1
CFG=3 run VLB_PAIR192=1
CFG=3 run VLB_PAIR192=2
CFG=4 run VLB_PAIR192=1
CFG=4 run VLB_PAIR192=2
for mode in 1 2; do
  echo "=== bench VLB_PAIR192=$mode" >> $L
  VLB_PAIR192=$mode timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager --no-other-configs > gpurun_out/r2b_bench_db_pair$mode.json 2>> $L; echo "--- exit $?" >> $L
done
grep -E "^===|^---|passed|failed|rror|us  " $L | head -150
python - <<'PY'
import json
for m in (1,2):
    try:
        d=json.loads(open('gpurun_out/r2b_bench_db_pair%d.json'%m).read().strip().splitlines()[-1])
        print(m,'ms/step %.3f value %.0f e2e %.0f gemm frac %.3f'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac']),{k:round(v['ms_per_step'],3) for k,v in d['kernel_profile'].items()})
    except Exception as e: print(m,'no bench',e)
PY
