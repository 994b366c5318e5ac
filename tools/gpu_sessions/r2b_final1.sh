#!/bin/bash
# round 2, session B: full suite + default bench on the current tree
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/r2b_final1.log
: > $L
echo "=== pytest -m gpu" >> $L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 >> $L
for z in 0 1; do
echo "=== bench VLB_ZERO_GRADS_IN_FORWARD=$z" >> $L
VLB_ZERO_GRADS_IN_FORWARD=$z timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager --no-other-configs 2>> $L | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms/step %.3f value %.0f e2e %.0f gemm frac %.3f launches %s'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac'],d.get('gpu_launches')),{k:round(v['ms_per_step'],3) for k,v in d['kernel_profile'].items()})
" >> $L
done
echo "=== smoke" >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 >> $L
grep -v Warn $L | tail -14
