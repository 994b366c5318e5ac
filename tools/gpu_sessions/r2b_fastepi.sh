#!/bin/bash
# round 2, session B: interior fast-path epilogue (packed fp32) -- suite, epilogue cost probe, layer GEMMs, step
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/r2b_fastepi.log
: > $L
echo "=== pytest -m gpu" >> $L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 >> $L; echo "--- exit $?" >> $L
echo "=== probe 6464 768 768" >> $L
timeout 300 python tools/epilogue_cost_probe.py 6464 768 768 >> $L 2>&1
echo "=== probe 6464 3072 768" >> $L
timeout 300 python tools/epilogue_cost_probe.py 6464 3072 768 >> $L 2>&1
echo "=== probe 6464 3072 768 PAIR192=2" >> $L
VLB_PAIR192=2 timeout 300 python tools/epilogue_cost_probe.py 6464 3072 768 >> $L 2>&1
for mode in 1 2; do
echo "=== layer gemm bench VLB_PAIR192=$mode" >> $L
VLB_PAIR192=$mode VLB_BENCH_CUBLAS=0 timeout 300 python tools/layer_gemm_bench.py 2>&1 | grep -v '^{' >> $L
done
for mode in 1 2; do
  echo "=== bench VLB_PAIR192=$mode" >> $L
  VLB_PAIR192=$mode timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager --no-other-configs > gpurun_out/r2b_bench_fastepi_pair$mode.json 2>> $L; echo "--- exit $?" >> $L
done
grep -v Warn $L | tail -100
python - <<'PY'
import json
for m in (1,2):
    try:
        d=json.loads(open('gpurun_out/r2b_bench_fastepi_pair%d.json'%m).read().strip().splitlines()[-1])
        print(m,'ms/step %.3f value %.0f e2e %.0f gemm frac %.3f'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac']),{k:round(v['ms_per_step'],3) for k,v in d['kernel_profile'].items()})
    except Exception as e: print(m,'no bench',e)
PY
