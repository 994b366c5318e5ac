#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/gpu_ddp4.log
: > $L
port=29630
run() {
  echo "=== $1 :: $2" >> $L
  local t0=$(date +%s)
  env $1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port \
     bench.py --gpus 2 --steps 20 --warmup 5 $2 >> $L 2>&1
  echo "--- exit $? wall $(( $(date +%s) - t0 ))s" >> $L
  port=$((port+1))
}
run "NCCL_DEBUG=WARN" "--no-cpu-baseline"
run "NCCL_DEBUG=WARN" "--no-cpu-baseline --no-graph"
echo "=== N=2 gradient equivalence" >> $L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29699 tools/ddp_equiv.py >> $L 2>&1
echo "--- exit $?" >> $L
python - <<'PY'
import json
hdr=None
for line in open('gpurun_out/gpu_ddp4.log'):
    if line.startswith('===') or line.startswith('---'): print(line.strip()[:120])
    if line.startswith('{"metric"'):
        d=json.loads(line)
        print('   ms/step %.3f value %.0f e2e %.0f graph %s' % (d['ms_per_step'], d['value'], d['e2e']['value'], d.get('cuda_graph')))
PY
grep -i "equiv\|rror" $L | head
