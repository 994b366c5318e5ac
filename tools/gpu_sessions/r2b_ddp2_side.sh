#!/bin/bash
# round 2, session B: two ranks, gradient casts / collective waits on a side stream vs on the compute stream
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/r2b_ddp2_side.log
: > $L
echo "=== pytest ddp" >> $L
timeout 600 python -m pytest tests/test_gpu_ddp.py -q -x >> $L 2>&1; echo "--- exit $?" >> $L
for side in 1 0; do
echo "=== N=2 VLB_DDP_SIDE_STREAM=$side" >> $L
VLB_DDP_SIDE_STREAM=$side timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager --no-other-configs 2>> $L | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms/step %.3f value %.0f e2e %.0f'%(d['ms_per_step'],d['value'],d['e2e']['value']))
" >> $L
done
grep -E "^===|ms/step|passed|failed|---" $L
