"""One GEMM shape, repeated (for ncu captures / timing): python tools/gemm_one.py MODE M N K BN [REPS]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import vlbert_b200
VF = vlbert_b200.functional
mode, M, N, K, bn = [int(x) for x in sys.argv[1:6]]
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 5
dev = "cuda"
if mode == 0:
    A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
elif mode == 1:
    A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(K, N, device=dev).bfloat16()
else:
    A = torch.randn(K, M, device=dev).bfloat16(); B = torch.randn(K, N, device=dev).bfloat16()
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16 if mode != 2 else torch.float32)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    VF.gemm(mode, A, B, out, force_bn=bn)
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); VF.gemm(mode, A, B, out, force_bn=bn); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort()
print("mode %d %dx%dx%d bn %d: median %.1f us (L2 flushed)  %.0f TFLOP/s" % (mode, M, N, K, bn, ts[len(ts) // 2], 2.0 * M * N * K / ts[len(ts) // 2] / 1e6))
