"""The fused attention kernels alone at a BASELINE config's shape (timing with CUDA events over graph-replayed launches, and an
ncu target): python tools/mhsa_one.py [B S heads] [--drop 0.1] [--reps 24]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import vlbert_b200
    VF = vlbert_b200.functional
    pos = [a for a in sys.argv[1:] if not a.startswith("--")]
    B, S, heads = [int(a) for a in pos[:3]] if len(pos) >= 3 else (64, 101, 12)
    p = float(sys.argv[sys.argv.index("--drop") + 1]) if "--drop" in sys.argv else 0.1
    reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 24
    H = heads * 64
    dev, bf = "cuda", torch.bfloat16
    g = torch.Generator(device=dev).manual_seed(0)
    sets = []
    rng = torch.tensor([1234, 1], dtype=torch.int64, device=dev)
    for _ in range(4):
        qkv = torch.randn(B * S, 3 * H, device=dev, generator=g).to(bf)
        dctx = torch.randn(B * S, H, device=dev, generator=g).to(bf)
        mask = torch.zeros(B, S, device=dev)
        drop = VF.DropSite(p, 1, rng) if p > 0 else None
        ctx, lse = VF.mhsa_forward(qkv, mask, B, S, H, heads, drop=drop)
        sets.append((qkv, dctx, mask, drop, ctx, lse, torch.zeros(3 * H, device=dev)))
    torch.cuda.synchronize()

    def fwd(s):
        VF.mhsa_forward(s[0], s[2], B, S, H, heads, drop=s[3])

    def bwd(s):
        VF.mhsa_backward(s[0], s[2], s[4], s[5], s[1], B, S, H, heads, drop=s[3], dbias=s[6])

    side = torch.cuda.Stream()
    for name, fn, flops in (("mhsa forward", fwd, 4.0), ("mhsa backward", bwd, 8.0)):
        with torch.cuda.stream(side):
            for s in sets:
                fn(s)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                for r in range(reps):
                    fn(sets[r % 4])
            graph.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(side)
            for _ in range(3):
                graph.replay()
            e1.record(side)
            torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (3 * reps)
        print("%-14s B=%d S=%d heads=%d p=%.2f  %8.2f us   %6.1f TFLOP/s (unpadded)" % (name, B, S, heads, p, us, flops * B * heads * S * S * 64 / us / 1e6))


if __name__ == "__main__":
    main()
