"""Mainloop ceilings without wave quantisation: shapes whose tile count is an exact multiple of the grid (M = 148 x 128 rows),
through the library in its three scheduling modes (single CTA, cta_group::2 pairs, B-multicast clusters) and through cuBLAS
(torch.matmul).  Graph-replayed, rotating operand sets.   python tools/gemm_ceiling.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlbert_b200  # noqa: E402

VF = vlbert_b200.functional
dev, bf = "cuda", torch.bfloat16
g = torch.Generator(device=dev).manual_seed(0)
REPS, SETS = 12, 3


def timed(fn):
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for i in range(SETS):
            fn(i)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            for r in range(REPS):
                fn(r % SETS)
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(side)
        for _ in range(3):
            graph.replay()
        e1.record(side)
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (3 * REPS)


M = 148 * 128
for (N, K) in ((256, 768), (256, 3072), (1024, 768), (1024, 3072), (768, 3072), (3072, 768)):
    A = [torch.randn(M, K, device=dev, generator=g).to(bf) for _ in range(SETS)]
    W = [(torch.randn(N, K, device=dev, generator=g) * 0.03).to(bf) for _ in range(SETS)]
    Wt = [w.t().contiguous() for w in W]
    O = [torch.empty(M, N, device=dev, dtype=bf) for _ in range(SETS)]
    fl = 2.0 * M * N * K
    row = ["%dx%dx%d" % (M, N, K)]
    for name, bn in (("single256", 256), ("single128", 128), ("cg2_256", 1256), ("mc2_256", 2256)):
        try:
            us = timed(lambda i: VF.gemm(0, A[i], W[i], O[i], force_bn=bn))
            row.append("%s %.1f us %.0f TF" % (name, us, fl / us / 1e6))
        except Exception as e:  # noqa
            row.append("%s ERR %s" % (name, str(e)[:40]))
    us = timed(lambda i: torch.matmul(A[i], Wt[i], out=O[i]))
    row.append("cublas %.1f us %.0f TF" % (us, fl / us / 1e6))
    print(" | ".join(row))
