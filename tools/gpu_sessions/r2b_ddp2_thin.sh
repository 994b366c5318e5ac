#!/bin/bash
# round 2, session B: two ranks, per-layer reductions on a CTA-capped second communicator
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/r2b_ddp2_thin.log
: > $L
for c in 0 4 8 16; do
echo "=== N=2 VLB_DDP_LAYER_CTAS=$c" >> $L
VLB_DDP_LAYER_CTAS=$c timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager --no-other-configs --skip-e2e 2>> $L | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms/step %.3f value %.0f gemm frac %.3f'%(d['ms_per_step'],d['value'],d['roofline']['frac']))
" >> $L
done
grep -E "^===|ms/step|thin layer|rror" $L
