#!/bin/bash
# round 2, session B: the driver's scaling commands at N = 8 (and N = 4) on the current tree
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/r2b_n8.log
: > $L
for n in 8 4; do
echo "=== N=$n" >> $L
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager --no-other-configs 2>> $L | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms/step %.3f value %.0f e2e %.0f n_gpus %d gemm frac %.3f'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['n_gpus'],d['roofline']['frac']))
" >> $L
echo "--- exit $?" >> $L
done
grep -E "^===|ms/step|---|rror|Traceback" $L | head -20
