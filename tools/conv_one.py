"""One res5-head 3x3 dilated convolution (implicit GEMM: fprop, wgrad, dgrad), a few repetitions -- for ncu captures.
    python tools/conv_one.py [rois] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import vlbert_b200
from vlbert_b200 import functional as VF
K = int(sys.argv[1]) if len(sys.argv) > 1 else 288
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = "cuda"
torch.manual_seed(0)
x = torch.randn(K, 14, 14, 512, device=dev).to(torch.bfloat16).requires_grad_(True)
w = (torch.randn(512, 512, 3, 3, device=dev) * 0.02).requires_grad_(True)
scale = torch.rand(512, device=dev) + 0.5
shift = torch.randn(512, device=dev) * 0.1
gy = torch.randn(K, 14, 14, 512, device=dev).to(torch.bfloat16)
for _ in range(reps):
    y = VF.conv_bn_act(x, w, scale, shift, stride=1, pad=2, dil=2, relu_mode=1)
    y.backward(gy)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
