#!/bin/bash
# round 2, session B: two ranks -- the cluster-launched pair kernel inside the captured step next to NCCL's channel CTAs
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/r2b_ddp2.log
: > $L
for mode in 1 0; do
echo "=== N=2 VLB_PAIR192=$mode" >> $L
VLB_PAIR192=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager --no-other-configs > gpurun_out/r2b_ddp2_pair$mode.json 2>> $L; echo "--- exit $?" >> $L
done
echo "=== pytest ddp" >> $L
timeout 600 python -m pytest tests/test_gpu_ddp.py -q -x >> $L 2>&1; echo "--- exit $?" >> $L
grep -v Warn $L | tail -30
python - <<'PY'
import json
for m in (1,0):
    try:
        d=json.loads(open('gpurun_out/r2b_ddp2_pair%d.json'%m).read().strip().splitlines()[-1])
        print(m,'ms/step %.3f value %.0f e2e %.0f gemm frac %.3f'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac']),{k:round(v['ms_per_step'],3) for k,v in d['kernel_profile'].items()})
    except Exception as e: print(m,'no bench',e)
PY
