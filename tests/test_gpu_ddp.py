"""Multi-GPU: gradient equivalence of the overlapped NCCL exchange (vl-bert_b200/ddp.py) -- the average of the per-shard
gradients after LayerGradReducer equals the single-GPU gradient of the concatenated batch (tools/ddp_equiv.py, one process
per GPU under torch.distributed.run) -- and a clean teardown (destroy_process_group returns).  Self-skips below 2 GPUs
(the driver's `pytest -m gpu` box has one; `gpurun --gpus 2` runs it)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("wire", ["bf16", "fp32"])
def test_ddp_gradients_equal_the_full_batch_gradients(wire):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (found %d)" % n)
    world = 2
    env = dict(os.environ, VLB_DDP_WIRE=wire)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tools", "ddp_equiv.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ddp gradient equivalence" in r.stdout and "teardown ok" in r.stdout, r.stdout[-2000:]
