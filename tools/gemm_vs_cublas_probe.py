"""One launch of each N=768 encoder GEMM shape through the library and through torch.matmul (cuBLASLt), for an ncu capture that
shows which kernels / tile shapes / grids cuBLAS picks for these shapes next to ours.
    ncu --set full --clock-control none -o gpurun_out/x python tools/gemm_vs_cublas_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlbert_b200  # noqa: E402

VF = vlbert_b200.functional
dev, bf = "cuda", torch.bfloat16
g = torch.Generator(device=dev).manual_seed(0)
M = 6464
for (N, K) in ((768, 3072), (768, 768), (3072, 768), (2304, 768)):
    a = torch.randn(M, K, device=dev, generator=g).to(bf)
    w = (torch.randn(N, K, device=dev, generator=g) * 0.03).to(bf)
    wt = w.t().contiguous()
    out = torch.empty(M, N, device=dev, dtype=bf)
    for _ in range(3):
        VF.gemm(0, a, w, out)                    # ours, NT, plain bf16 epilogue
        torch.matmul(a, wt, out=out)             # cuBLASLt
    torch.cuda.synchronize()
print("done")
