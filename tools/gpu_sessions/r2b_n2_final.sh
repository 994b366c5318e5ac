#!/bin/bash
# round 2, session B: the final tree at N = 2 (driver command)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 20 --warmup 5 2> gpurun_out/r2b_n2_final.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('N=2 final: ms/step %.3f value %.0f e2e %.0f n_gpus %d gemm frac %.3f'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['n_gpus'],d['roofline']['frac']))
"
echo "exit ${PIPESTATUS[0]}"; tail -3 gpurun_out/r2b_n2_final.err | cut -c1-300
