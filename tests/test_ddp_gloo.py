"""CPU, world_size 2, gloo: the data-parallel gradient exchange (vl-bert_b200/ddp.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import vlbert_b200
        red = vlbert_b200.ddp.LayerGradReducer()
        assert red.world == world and not red.avg
        # per-layer flat buffers, launched in backward order, drained once
        layers = [torch.full((1000,), float(rank + 1) * (l + 1)) for l in range(3)]
        for l in (2, 1, 0):
            red.launch(layers[l])
        red.drain()
        mean = sum(range(1, world + 1)) / world
        ok = all(torch.allclose(layers[l], torch.full((1000,), mean * (l + 1))) for l in range(3))
        # remaining parameters
        ps = [torch.nn.Parameter(torch.zeros(7, 3)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(2))]
        ps[0].grad = torch.full((7, 3), float(rank))
        ps[1].grad = torch.full((5,), 10.0 * rank)
        big = torch.nn.Parameter(torch.zeros(1200, 1000))  # >= 1M elements: reduced in place, not coalesced
        big.grad = torch.full((1200, 1000), 3.0 * rank)
        ps.append(big)
        red.reduce_params(ps)  # ps[2] has no grad: skipped
        ok = ok and torch.allclose(big.grad, torch.full((1200, 1000), 1.5 * (world - 1)))
        ok = ok and torch.allclose(ps[0].grad, torch.full((7, 3), (world - 1) / 2)) and torch.allclose(ps[1].grad, torch.full((5,), 5.0 * (world - 1)))
        ok = ok and ps[2].grad is None
        # gradient equivalence: mean of per-shard gradients == gradient of the mean loss over the concatenated batch
        torch.manual_seed(0)
        w = torch.randn(4, 4, requires_grad=True)
        x = torch.randn(8, 4)
        shard = x[rank * 4:(rank + 1) * 4]
        (shard @ w).pow(2).mean().backward()
        g = w.grad.clone()
        red.launch(g)
        red.drain()
        w2 = w.detach().clone().requires_grad_(True)
        (x @ w2).pow(2).mean().backward()
        ok = ok and torch.allclose(g, w2.grad, atol=1e-6)
        # bf16 wire format with the depth-2 pipeline (the NCCL default): five buffers launched back to back, each result is
        # cast back over the fp32 gradients two launches later / at drain; values chosen exactly representable in bf16
        red16 = vlbert_b200.ddp.LayerGradReducer(wire_dtype=torch.bfloat16, pipeline_depth=2)
        bufs = [torch.full((4096,), float(rank + 1) * (l + 1)) for l in range(5)]
        for l in range(5):
            red16.launch(bufs[l])
            assert len(red16.pending) <= 2
        red16.drain()
        ok = ok and all(torch.allclose(bufs[l], torch.full((4096,), mean * (l + 1))) for l in range(5))
        ok = ok and len(red16._staging) == 5 and all(b.dtype == torch.bfloat16 for b in red16._staging.values())
        bufs[0].fill_(float(rank + 1))
        red16.launch(bufs[0])           # same buffer again: the staging buffer is re-used (static addresses under CUDA graphs)
        red16.drain()
        ok = ok and len(red16._staging) == 5 and torch.allclose(bufs[0], torch.full((4096,), mean))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_layer_grad_reducer_world2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
