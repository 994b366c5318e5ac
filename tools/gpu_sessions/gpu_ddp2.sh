#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/gpu_ddp2.log
: > $L
nvidia-smi --query-gpu=index,name --format=csv >> $L
echo "=== N=2 torchrun bench" >> $L
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
   bench.py --gpus 2 --steps 10 --warmup 3 >> $L 2>&1
echo "--- exit $?" >> $L
echo "=== N=2 torchrun bench, CUDA graph incl. NCCL" >> $L
VLB_GRAPH_DDP=1 NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 \
   bench.py --gpus 2 --steps 10 --warmup 3 >> $L 2>&1
echo "--- exit $?" >> $L
echo "=== N=2 gradient equivalence" >> $L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tools/ddp_equiv.py >> $L 2>&1
echo "--- exit $?" >> $L
grep -E "^===|^---|rror|metric|equiv" $L | cut -c1-420
