#!/bin/bash
# round 2, session B: smem-staged epilogue operands (pair kernel, last tile), coalesced D prologue of the attention backward
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/r2b_stage.log
: > $L
echo "=== pytest -m gpu" >> $L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> $L; echo "--- exit $?" >> $L
for st in 1 0; do
echo "=== probe 6464 768 768 VLB_EPI_STAGE=$st" >> $L
VLB_EPI_STAGE=$st timeout 300 python tools/epilogue_cost_probe.py 6464 768 768 2>&1 | grep -E "residual|plain -> bf16" >> $L
done
echo "=== layer gemm bench" >> $L
VLB_BENCH_CUBLAS=0 timeout 300 python tools/layer_gemm_bench.py 2>&1 | grep -v '^{' >> $L
echo "=== mhsa" >> $L
timeout 300 python tools/mhsa_one.py 64 101 12 --drop 0 >> $L 2>&1
timeout 300 python tools/mhsa_one.py 64 121 12 --drop 0 >> $L 2>&1
echo "=== trace resid" >> $L
timeout 120 python tools/gemm_trace.py 6464 768 768 resid >> $L 2>&1
echo "=== bench" >> $L
timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager --no-other-configs > gpurun_out/r2b_bench_stage.json 2>> $L; echo "--- exit $?" >> $L
grep -v Warn $L | tail -60
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2b_bench_stage.json').read().strip().splitlines()[-1])
print('ms/step %.3f value %.0f e2e %.0f gemm frac %.3f'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac']),{k:round(v['ms_per_step'],3) for k,v in d['kernel_profile'].items()})
PY
