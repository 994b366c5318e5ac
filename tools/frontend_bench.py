"""Times the end-to-end region-feature front end (BASELINE config 5 shape, per GPU: 8 images of 600x1000, 36 boxes each)
forward + backward on one B200, and -- for orientation -- the same network written with torch's cuDNN convolutions
(bf16 autocast, channels_last), which is what the reference's `common/fast_rcnn.py` runs on a GPU.

    python tools/frontend_bench.py [--images 8] [--h 600] [--w 1000] [--boxes 36] [--steps 5] [--torch]
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))


def torch_forward(m, images, boxes, box_mask, im_info):
    """the reference's module graph on torch ops (cuDNN), reading the library module's parameters"""
    import torchvision.ops as tvo

    def cba(x, conv, bn, relu=True):
        y = F.conv2d(x, conv.weight, None, conv.stride, conv.padding, conv.dilation)
        y = F.batch_norm(y, bn.running_mean, bn.running_var, bn.weight, bn.bias, False, 0.0, bn.eps)
        return F.relu(y) if relu else y

    def block(b, x):
        o = cba(cba(x, b.conv1, b.bn1), b.conv2, b.bn2)
        o = cba(o, b.conv3, b.bn3, relu=False)
        idn = x if b.downsample is None else cba(x, b.downsample[0], b.downsample[1], relu=False)
        return F.relu(o + idn)

    bb = m.backbone
    with torch.autocast("cuda", dtype=torch.bfloat16):
        with torch.no_grad():
            x = images.contiguous(memory_format=torch.channels_last)
            x = F.max_pool2d(cba(x, bb.conv1, bb.bn1), 3, 2, 1)
            for b in bb.layer1:
                x = block(b, x)
        for layer in (bb.layer2, bb.layer3):
            for b in layer:
                x = block(b, x)
        B, R = box_mask.shape
        bidx = torch.arange(B, device=x.device, dtype=torch.float32).view(B, 1, 1).expand(B, R, 1)
        rois = torch.cat((bidx, boxes[:, :, :4]), 2).view(B * R, 5)
        p = tvo.roi_align(x.float(), rois, (14, 14), 1.0 / 16, 1, aligned=False).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        for b in m.roi_head_feature_extractor:
            p = block(b, p)
        post = p.float().mean((2, 3))
    return post


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=8)
    ap.add_argument("--h", type=int, default=600)
    ap.add_argument("--w", type=int, default=1000)
    ap.add_argument("--boxes", type=int, default=36)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--torch", action="store_true")
    ap.add_argument("--graph", action="store_true", help="also time the step captured in a CUDA graph (vlbert_b200.GraphedStep)")
    a = ap.parse_args()
    import vlbert_b200
    from vlbert_b200 import _lib
    from synth import frontend_config
    dev = "cuda"
    torch.manual_seed(0)
    m = vlbert_b200.FastRCNN(frontend_config(101), True, 768, False).to(dev).eval()
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.weight.uniform_(0.3, 0.6)
                mod.running_var.uniform_(0.5, 1.5)
    m.compact_rois = False
    B, R = a.images, a.boxes
    g = torch.Generator().manual_seed(1)
    images = torch.randn(B, 3, a.h, a.w, generator=g).to(dev)
    x1 = torch.rand(B, R, generator=g) * a.w * 0.6
    y1 = torch.rand(B, R, generator=g) * a.h * 0.6
    boxes = torch.stack((x1, y1, x1 + 16 + torch.rand(B, R, generator=g) * a.w * 0.38, y1 + 16 + torch.rand(B, R, generator=g) * a.h * 0.38), -1).to(dev)
    box_mask = torch.ones(B, R, dtype=torch.bool, device=dev)
    im_info = torch.tensor([[float(a.w), float(a.h), 1, 1]] * B, device=dev)
    gw = torch.randn(B, R, 768, generator=g).to(dev)

    def ours():
        m.zero_grad(set_to_none=True)
        out = m(images=images, boxes=boxes, box_mask=box_mask, im_info=im_info)
        (out["obj_reps"] * gw).sum().backward()

    gp = torch.randn(B * R, 2048, generator=g).to(dev)

    def theirs():
        m.zero_grad(set_to_none=True)
        post = torch_forward(m, images, boxes, box_mask, im_info)
        (post * gp).sum().backward()

    res = {"shape": {"images": B, "h": a.h, "w": a.w, "boxes": R}}
    runs = [("library", ours)]
    if a.graph:
        def loss_fn(mod, im, bx, bm, info):
            return (mod(images=im, boxes=bx, box_mask=bm, im_info=info)["obj_reps"] * gw).sum()
        gstep = vlbert_b200.GraphedStep(m, loss_fn, (images, boxes, box_mask, im_info))
        runs.append(("library_cuda_graph", lambda: gstep()))
    if a.torch:
        runs.append(("torch_cudnn_bf16", theirs))
    for name, fn in runs:
        for _ in range(a.warmup):
            fn()
        torch.cuda.synchronize()
        n0 = _lib.lib().vlb_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        res[name] = {"ms_per_step": round(ms, 3), "images_per_s": round(B / ms * 1e3, 2),
                     "launches_per_step": (_lib.lib().vlb_launch_count() - n0) // a.steps,
                     "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}
    # per-category device time of one eager step (events around every library launch; serialises nothing)
    import ctypes
    lib = _lib.lib()
    ours(); torch.cuda.synchronize()
    lib.vlb_profile_enable(1)
    ours(); torch.cuda.synchronize()
    lib.vlb_profile_enable(0)
    pms, pwork, pcnt = (ctypes.c_double * 12)(), (ctypes.c_double * 12)(), (ctypes.c_int64 * 12)()
    _lib.check(lib.vlb_profile_collect(pms, pwork, pcnt))
    names = ["gemm_nt", "gemm_nn", "gemm_tn", "mhsa_fwd", "mhsa_bwd", "ln_fwd", "ln_bwd", "other", "im2col", "col2im", "conv_elt", "roi_nhwc"]
    res["profile"] = {n: {"ms": round(pms[i], 3), "launches": pcnt[i],
                          ("tflops" if i < 5 else "gbps"): round(pwork[i] / max(pms[i], 1e-9) / (1e9 if i < 5 else 1e6), 1)}
                      for i, n in enumerate(names) if pcnt[i]}
    res["profile_sum_ms"] = round(sum(pms), 3)
    # analytic conv MACs (fwd) for the roofline note: backbone trainable part x3 (fwd + dgrad + wgrad), frozen part x1
    print(json.dumps(res))


if __name__ == "__main__":
    main()
