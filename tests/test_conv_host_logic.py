"""CPU: the host-side algebra of the convolution front end (pure torch, no kernels): the GEMM formulations the library
launches must be the convolution / its gradients the reference's nn.Conv2d + frozen BatchNorm define
(common/backbone/resnet/resnet.py:98-118, common/fast_rcnn.py:122-126)."""
import pytest
import torch
import torch.nn.functional as F

import vlbert_b200
from vlbert_b200 import functional as VF
from vlbert_b200 import resnet as R

GEOMS = [(1, 1, 0, 1), (1, 2, 0, 1), (3, 1, 1, 1), (3, 1, 2, 2), (3, 2, 1, 1), (7, 2, 3, 1)]   # (k, stride, pad, dil)


def tap_major_cols(x, k, stride, pad, dil):
    """im2col matrix with rows = output pixels (n, ho, wo) and columns ordered (r, s, c): what vlb_im2col_nhwc writes and
    what a TMA im2col-mode load sequence over the k-blocks delivers"""
    N, C, H, W = x.shape
    u = F.unfold(x, k, dilation=dil, padding=pad, stride=stride)                 # [N, C*k*k, L], channel-major
    L = u.shape[-1]
    return u.view(N, C, k * k, L).permute(0, 3, 2, 1).reshape(N * L, k * k * C)


@pytest.mark.parametrize("k,stride,pad,dil", GEOMS)
def test_forward_is_a_gemm_on_tap_major_columns(k, stride, pad, dil):
    g = torch.Generator().manual_seed(k * 10 + stride)
    C, Cout = (3 if k == 7 else 16), 24
    x = torch.randn(2, C, 13, 11, generator=g)
    w = torch.randn(Cout, C, k, k, generator=g)
    w16 = VF.weight_to_gemm(w)                                                    # bf16 [Cout, Kp], zero padded to a multiple of 8
    assert w16.shape[1] % 8 == 0 and w16.shape[1] >= k * k * C
    assert bool((w16[:, k * k * C:] == 0).all())
    col = tap_major_cols(x, k, stride, pad, dil)
    y = col @ w16[:, :k * k * C].float().t()
    ref = F.conv2d(x, w.to(torch.bfloat16).float(), stride=stride, padding=pad, dilation=dil)
    Ho, Wo = VF._conv_out(13, k, stride, pad, dil), VF._conv_out(11, k, stride, pad, dil)
    assert ref.shape[2:] == (Ho, Wo)
    assert torch.allclose(y.view(2, Ho, Wo, Cout).permute(0, 3, 1, 2), ref, atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize("k,pad,dil", [(1, 0, 1), (3, 1, 1), (3, 2, 2), (3, 0, 1), (5, 2, 1)])
def test_data_gradient_is_a_convolution_with_the_flipped_filter(k, pad, dil):
    """stride-1 dgrad as launched by ConvBnActFn.backward: dX = conv(dY, weight_to_dgrad(W), pad' = dil*(k-1) - pad)."""
    g = torch.Generator().manual_seed(k + pad)
    C, Cout = 8, 16
    x = torch.randn(2, C, 12, 10, generator=g, requires_grad=True)
    w = torch.randn(Cout, C, k, k, generator=g).to(torch.bfloat16).float()
    y = F.conv2d(x, w, padding=pad, dilation=dil)
    gy = torch.randn(y.shape, generator=g)
    y.backward(gy)
    w16 = VF.weight_to_gemm(w)
    wd = VF.weight_to_dgrad(w16, Cout, k, k, C)                                    # [C, (r', s', cout)]
    assert wd.shape == (C, k * k * Cout)
    pad_t = dil * (k - 1) - pad
    col = tap_major_cols(gy, k, 1, pad_t, dil)                                     # rows: pixels of dX, columns (r', s', cout)
    dx = (col @ wd.float().t()).view(2, 12, 10, C).permute(0, 3, 1, 2)
    assert torch.allclose(dx, x.grad, atol=1e-4, rtol=1e-4)


def test_weight_gradient_layout_round_trip():
    """wgrad GEMM output [Cout, (r, s, c)] -> parameter layout [Cout, C, kh, kw] (ConvBnActFn.backward)"""
    g = torch.Generator().manual_seed(3)
    C, Cout, k = 8, 16, 3
    x = torch.randn(2, C, 9, 7, generator=g)
    w = torch.randn(Cout, C, k, k, generator=g, requires_grad=True)
    y = F.conv2d(x, w, padding=1)
    gy = torch.randn(y.shape, generator=g)
    y.backward(gy)
    col = tap_major_cols(x, k, 1, 1, 1)
    dwk = gy.permute(0, 2, 3, 1).reshape(-1, Cout).t() @ col                         # dY^T . col
    dw = dwk.view(Cout, k, k, C).permute(0, 3, 1, 2)
    assert torch.allclose(dw, w.grad, atol=1e-4, rtol=1e-4)


def test_frozen_batchnorm_is_the_cached_affine_map():
    bn = torch.nn.BatchNorm2d(12)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(12, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(12, generator=g))
        bn.running_mean.copy_(torch.randn(12, generator=g))
        bn.running_var.copy_(torch.rand(12, generator=g) + 0.3)
    bn.eval()
    scale, shift = R._bn_affine(bn)
    x = torch.randn(3, 12, 5, 4, generator=g)
    assert torch.allclose(x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1), bn(x), atol=1e-5, rtol=1e-5)
    s2, _ = R._bn_affine(bn)
    assert s2 is scale                       # cached
    with torch.no_grad():
        bn.running_var.mul_(2.0)             # a checkpoint load / in-place update invalidates the cache
    s3, _ = R._bn_affine(bn)
    assert s3 is not scale and not torch.equal(s3, scale)
    bn.train()
    with pytest.raises(NotImplementedError):
        R._bn_affine(bn)


def test_end_to_end_fastrcnn_construction_follows_the_cfg():
    from synth import frontend_config
    m = vlbert_b200.FastRCNN(frontend_config(101), True, 768, False)
    assert [len(getattr(m.backbone, "layer%d" % i)) for i in (1, 2, 3)] == [3, 4, 23]
    head = m.roi_head_feature_extractor
    assert head[0].conv2.dilation == (2, 2) and head[0].conv2.stride == (1, 1) and head[0].downsample[0].stride == (1, 1)   # IMAGE_C5_DILATED
    assert m.backbone.layer2[0].conv1.stride == (2, 2) and m.backbone.layer2[0].conv2.stride == (1, 1)                     # IMAGE_STRIDE_IN_1x1
    assert m.head[0] is head
    frozen = [n for n, p in m.named_parameters() if not p.requires_grad]
    assert all(n.startswith(("backbone.conv1", "backbone.bn1", "backbone.layer1")) or ".bn" in n or "downsample.1" in n for n in frozen)
    cfg = frontend_config(101)
    cfg.NETWORK.IMAGE_FROZEN_BACKBONE_STAGES = [1, 2, 5]
    m5 = vlbert_b200.FastRCNN(cfg, True, 768, False)
    assert not any(p.requires_grad for p in m5.roi_head_feature_extractor.parameters())
    cfg = frontend_config(34)
    with pytest.raises(NotImplementedError):
        vlbert_b200.FastRCNN(cfg, True, 768, False)
