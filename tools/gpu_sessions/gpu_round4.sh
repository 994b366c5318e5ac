#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/gpu_round4.log
: > $L
run() { echo "=== $*" >> $L; timeout "$1" "${@:2}" >> $L 2>&1; echo "--- exit $?" >> $L; }
run 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "grouped"
run 600 python -m pytest tests/test_gpu_modules.py -q -x
run 300 python -c "import __graft_entry__ as g; g.smoke()"
run 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
for cfgv in "128 1" "128 3" "256 1" "256 3"; do
  set -- $cfgv
  echo "=== VLB_WGRAD_BN=$1 VLB_WGRAD_SPLIT=$2" >> $L
  VLB_WGRAD_BN=$1 VLB_WGRAD_SPLIT=$2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --skip-e2e >> $L 2>&1
done
run 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --skip-e2e --text 20 --regions 100
grep -E "^===|^---|passed|failed|rror|smoke" $L | head -40
python - <<'PY'
import json
for line in open('gpurun_out/gpu_round4.log'):
    if line.startswith('==='): hdr=line.strip()
    if line.startswith('{"metric"'):
        d=json.loads(line)
        if d['ms_per_step']<1000:
            print(hdr[:70]); print('  ms/step %.3f value %.0f e2e %.0f graph %s gemm frac %.3f ' % (d['ms_per_step'], d['value'], d['e2e']['value'], d.get('cuda_graph'), d['roofline']['frac']), {k:round(v['ms_per_step'],3) for k,v in d['kernel_profile'].items()})
PY
