"""TEST INFRASTRUCTURE ONLY.  fp32 torch stand-ins for the kernel-backed entry points of vlbert_b200.functional that the
front-end modules call, installed by a pytest fixture (monkeypatch) so that the HOST logic of vlbert_b200.resnet /
vlbert_b200.modules.FastRCNN -- which block gets which stride / dilation, residual wiring, frozen stages, RoI batching,
slot compaction, masking, the regularisation head -- can be checked on the CPU against the oracle.  The product never
imports this file and has no CPU path of its own.

Layout contract mirrored from the library: activations are NHWC (here fp32 instead of bf16)."""
import torch
import torch.nn.functional as F

import roi_align as roi_oracle
import vlbert_oracle as vo


def conv_bn_act(x, weight, scale, shift, resid=None, stride=1, pad=0, dil=1, relu_mode=1, w16=None, bag=None, role=None):
    y = F.conv2d(x.permute(0, 3, 1, 2), weight, stride=stride, padding=pad, dilation=dil)
    y = y * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    if resid is not None:
        y = y + resid.permute(0, 3, 1, 2)
    if relu_mode:
        y = torch.relu(y)
    return y.permute(0, 2, 3, 1).contiguous()


def maxpool3x3s2(x):
    return F.max_pool2d(x.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1).contiguous()


def nchw_to_nhwc(x):
    return x.float().permute(0, 2, 3, 1).contiguous()


def weight_to_gemm(weight, Kp=None):
    return weight


class _RoIAlignNHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, rois, ph, pw, scale, sr):
        f = feat.detach().permute(0, 3, 1, 2).contiguous()
        ctx.save_for_backward(rois)
        ctx.cfg = (ph, pw, scale, sr, tuple(f.shape))
        out = torch.from_numpy(roi_oracle.roi_align_forward(f.numpy(), rois.numpy(), scale, ph, pw, sr))
        return out.permute(0, 2, 3, 1).contiguous()

    @staticmethod
    def backward(ctx, g):
        (rois,) = ctx.saved_tensors
        ph, pw, scale, sr, (N, C, H, W) = ctx.cfg
        gin = roi_oracle.roi_align_backward(g.permute(0, 3, 1, 2).contiguous().numpy(), rois.numpy(), scale, ph, pw, N, C, H, W, sr)
        return torch.from_numpy(gin).permute(0, 2, 3, 1).contiguous(), None, None, None, None, None


class _AvgPool(object):
    @staticmethod
    def apply(x):
        return x.mean((1, 2))


class _Region(object):
    @staticmethod
    def apply(boxes, weight, bias, box_mask, im_info):
        return vo.fast_rcnn_precomputed(boxes, box_mask, im_info, weight, bias)


def install(monkeypatch):
    from vlbert_b200 import functional as VF
    monkeypatch.setattr(VF, "conv_bn_act", conv_bn_act)
    monkeypatch.setattr(VF, "maxpool3x3s2", maxpool3x3s2)
    monkeypatch.setattr(VF, "nchw_to_nhwc_bf16", nchw_to_nhwc)
    monkeypatch.setattr(VF, "weight_to_gemm", weight_to_gemm)
    monkeypatch.setattr(VF, "RoIAlignNHWCFn", _RoIAlignNHWC)
    monkeypatch.setattr(VF, "AvgPoolFn", _AvgPool)
    monkeypatch.setattr(VF, "RegionFn", _Region)
