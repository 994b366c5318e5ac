"""ResNet-C4 backbone + res5 RoI head of the region-feature front end on the library's conv path (NHWC bf16).

Parameter / buffer names are exactly the reference's (common/backbone/resnet/resnet.py:74-199, torchvision style:
`conv1.weight, bn1.{weight,bias,running_mean,running_var,num_batches_tracked}, layer{1,2,3}.{i}.conv{1,2,3}.weight,
...bn{1,2,3}.*, ...downsample.{0,1}.*`) so reference checkpoints load with strict=True.  BatchNorm runs in eval mode with
frozen statistics (NETWORK.IMAGE_FROZEN_BN, common/fast_rcnn.py:88-92,122-126): it is a per-channel affine map that is
fused into the convolution's GEMM epilogue.
"""
import torch
import torch.nn as nn

from . import functional as VF


class Bottleneck(nn.Module):
    """common/backbone/resnet/resnet.py:74-118 (conv1 1x1 -> conv2 3x3 -> conv3 1x1, expansion 4)."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, dilation=1, stride_in_1x1=False):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=1, stride=1 if not stride_in_1x1 else stride, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=stride if not stride_in_1x1 else 1, dilation=dilation,
                               padding=dilation, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * self.expansion, kernel_size=1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * self.expansion)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):  # x: NHWC bf16
        # identity blocks: the residual-branch gradient is added inside conv1's data-gradient GEMM (see ConvBnActFn.backward)
        fuse = self.downsample is None and self.conv1.stride[0] == 1 and x.requires_grad and torch.is_grad_enabled()
        bag = {} if fuse else None
        o = _cba(x, self.conv1, self.bn1, relu_mode=1, bag=bag, role="consume" if fuse else None)
        o = _cba(o, self.conv2, self.bn2, relu_mode=1)
        identity = x if self.downsample is None else _cba(x, self.downsample[0], self.downsample[1], relu_mode=0)
        return _cba(o, self.conv3, self.bn3, resid=identity, relu_mode=2, bag=bag, role="produce" if fuse else None)


def _bn_affine(bn):
    """eval-mode BatchNorm as y = x * scale + shift; cached until a buffer/parameter changes."""
    if bn.training:
        raise NotImplementedError("vlbert_b200: BatchNorm must be frozen / in eval mode on the fused conv path "
                                  "(NETWORK.IMAGE_FROZEN_BN: true, FastRCNN.bn_eval())")
    key = (bn.weight._version, bn.bias._version, bn.running_mean._version, bn.running_var._version, bn.weight.data_ptr())
    cache = getattr(bn, "_vlb_affine", None)
    if cache is None or cache[0] != key:
        with torch.no_grad():
            scale = (bn.weight.float() / torch.sqrt(bn.running_var.float() + bn.eps)).contiguous()
            shift = (bn.bias.float() - bn.running_mean.float() * scale).contiguous()
        bn._vlb_affine = cache = (key, scale, shift)
    return cache[1], cache[2]


def _cba(x, conv, bn, resid=None, relu_mode=1, bag=None, role=None):
    scale, shift = _bn_affine(bn)
    w16 = None
    if not conv.weight.requires_grad:  # frozen convolution: keep the bf16 GEMM operand
        key = (conv.weight._version, conv.weight.data_ptr())
        cache = getattr(conv, "_vlb_w16", None)
        if cache is None or cache[0] != key:
            conv._vlb_w16 = cache = (key, VF.weight_to_gemm(conv.weight))
        w16 = cache[1]
    return VF.conv_bn_act(x, conv.weight, scale, shift, resid=resid, stride=conv.stride[0], pad=conv.padding[0],
                          dil=conv.dilation[0], relu_mode=relu_mode, w16=w16, bag=bag, role=role)


def make_layer(inplanes, planes, blocks, stride=1, dilation=1, stride_in_1x1=False):
    """resnet.py:158-173 (_make_layer)"""
    downsample = None
    if stride != 1 or inplanes != planes * Bottleneck.expansion:
        downsample = nn.Sequential(nn.Conv2d(inplanes, planes * Bottleneck.expansion, kernel_size=1, stride=stride, bias=False),
                                   nn.BatchNorm2d(planes * Bottleneck.expansion))
    layers = [Bottleneck(inplanes, planes, stride, downsample, dilation, stride_in_1x1=stride_in_1x1)]
    inplanes = planes * Bottleneck.expansion
    for _ in range(1, blocks):
        layers.append(Bottleneck(inplanes, planes, dilation=dilation))
    return nn.Sequential(*layers), inplanes


class ResNetC4(nn.Module):
    """conv1-bn1-relu-maxpool, layer1..3 (resnet.py:121-199 with expose_stages=[4]); forward returns body4 in NHWC bf16."""

    def __init__(self, layers=(3, 4, 23), stride_in_1x1=True):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        inplanes = 64
        for i, (planes, blocks, stride) in enumerate(zip((64, 128, 256), layers, (1, 2, 2))):
            layer, inplanes = make_layer(inplanes, planes, blocks, stride=stride, stride_in_1x1=stride_in_1x1)
            setattr(self, "layer%d" % (i + 1), layer)
        self.inplanes = inplanes
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def frozen_parameters(self, frozen_stages=None, frozen_bn=False):
        """resnet.py:217-241"""
        if frozen_bn:
            for m in self.modules():
                if isinstance(m, nn.BatchNorm2d):
                    for p in m.parameters():
                        p.requires_grad = False
        for stage in (frozen_stages or []):
            mods = [self.conv1, self.bn1] if stage == 1 else [getattr(self, "layer%d" % (stage - 1))]
            for m in mods:
                for p in m.parameters():
                    p.requires_grad = False

    def forward(self, images):
        """images: [N, 3, H, W] float (NCHW, like the reference) -> {'body4': NHWC bf16 [N, H/16, W/16, 1024]}"""
        x = VF.nchw_to_nhwc_bf16(images)
        frozen_stem = not self.conv1.weight.requires_grad
        frozen_l1 = not any(p.requires_grad for p in self.layer1.parameters())
        if not (frozen_stem and frozen_l1):
            raise NotImplementedError("vlbert_b200: the stem and layer1 must be frozen (IMAGE_FROZEN_BACKBONE_STAGES: [1, 2]); "
                                      "the max-pool has no backward on the fused path")
        with torch.no_grad():
            x = _cba(x, self.conv1, self.bn1, relu_mode=1)
            x = VF.maxpool3x3s2(x)
            x = self.layer1(x)
        x = self.layer2(x)
        x = self.layer3(x)
        return {"body4": x}
