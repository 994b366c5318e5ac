#!/bin/bash
# N=2 data-parallel tuning: does the overlapped NCCL all-reduce steal SMs from the persistent GEMM grids?
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/gpu_ddp3.log
: > $L
port=29530
run() {  # run "<env assignments>"
  echo "=== $1" >> $L
  env $1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port \
     bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --skip-e2e >> $L 2>&1
  echo "--- exit $?" >> $L
  port=$((port+1))
}
run "NCCL_DEBUG=WARN"
run "NCCL_DEBUG=WARN VLB_SM_LIMIT=132"
run "NCCL_DEBUG=WARN VLB_SM_LIMIT=140"
run "NCCL_DEBUG=WARN NCCL_MAX_NCHANNELS=4"
run "NCCL_DEBUG=WARN NCCL_MAX_NCHANNELS=4 VLB_SM_LIMIT=144"
run "NCCL_DEBUG=WARN NCCL_MAX_NCHANNELS=8 VLB_SM_LIMIT=140"
NCCL_DEBUG=INFO timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29599 \
     bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --skip-e2e 2>&1 | grep -i -E "channels|nvls|Connected all|P2P|coll" | head -12 >> $L
python - <<'PY'
import json
hdr=None
for line in open('gpurun_out/gpu_ddp3.log'):
    if line.startswith('==='): hdr=line.strip()
    if line.startswith('{"metric"'):
        d=json.loads(line)
        print(hdr[:90], '| ms/step %.3f value %.0f' % (d['ms_per_step'], d['value']))
PY
tail -12 $L | cut -c1-200
