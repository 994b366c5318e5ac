#!/bin/bash
# N=4: CUDA-graph step with captured NCCL vs eager
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/gpu_ddp5.log
: > $L
N=${1:-4}
port=29730
run() {
  echo "=== N=$N $1" >> $L
  local t0=$(date +%s)
  NCCL_DEBUG=WARN timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
     bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline $1 >> $L 2>&1
  echo "--- exit $? wall $(( $(date +%s) - t0 ))s" >> $L
  port=$((port+1))
}
run ""
run "--no-graph"
python - <<'PY'
import json
for line in open('gpurun_out/gpu_ddp5.log'):
    if line.startswith('===') or line.startswith('---'): print(line.strip()[:120])
    if line.startswith('{"metric"'):
        d=json.loads(line)
        print('   ms/step %.3f value %.0f e2e %.0f graph %s' % (d['ms_per_step'], d['value'], d['e2e']['value'], d.get('cuda_graph')))
PY
grep -i "rror\|Traceback\|hang" $L | head
