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
    def apply(boxes, weight, bias, box_mask, im_info, drop=None):
        assert drop is None, "the CPU stand-ins have no dropout: run the module in eval mode or with p = 0"
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


# ------------------------------------------------------------------------------------------------------------------
# encoder side: stand-ins with the call signatures of vlbert_b200.functional.PackIndex / EmbeddingFn / EncoderFn /
# GatherRowsFn / EncoderWeights (fp32 torch ops on the oracle's primitives)
# ------------------------------------------------------------------------------------------------------------------
class PackIndexShim(object):
    def check(self):
        pass

    def __init__(self, text_mask, object_mask, text_token_type_ids, S, pos_offset):
        kind, src, pos_id, te, oe, S_true = vo.pack_indices(text_mask.bool(), object_mask.bool())
        B, T = text_mask.shape
        R = object_mask.shape[1]
        assert S >= S_true, "max_length_hint smaller than the longest packed sequence"
        if S > S_true:                      # a hint larger than the true maximum: the extra rows are padding rows
            pad = S - S_true
            kind = torch.cat((kind, torch.full((B, pad), 3, dtype=kind.dtype)), 1)
            src = torch.cat((src, torch.zeros((B, pad), dtype=src.dtype)), 1)
            pos_id = torch.cat((pos_id, torch.arange(S_true, S).unsqueeze(0).expand(B, pad)), 1)
        self.B, self.T, self.R, self.S, self.pos_offset = B, T, R, S, int(pos_offset)
        self.kind, self.src = kind, src
        self.pos_id = pos_id + int(pos_offset)
        type_id = torch.zeros((B, S), dtype=torch.long)
        type_id = torch.where(kind == 0, text_token_type_ids.long().gather(1, src.clamp(max=T - 1)), type_id)
        self.type_id = torch.where((kind == 1) | (kind == 2), torch.full_like(type_id, 2), type_id)
        self.add_mask = torch.where(kind == 3, torch.tensor(-10000.0), torch.tensor(0.0))
        rank = torch.cumsum(object_mask.long(), 1) - 1
        row = torch.arange(B).unsqueeze(1) * S + te.unsqueeze(1) + rank
        self.obj_row = torch.where(object_mask.bool(), row, torch.full_like(row, -1))


class _EmbeddingShim(object):
    @staticmethod
    def apply(text_visual, object_vl, word, end, pos, typ, ln_w, ln_b, vt_w, vt_b, vo_w, vo_b, ids, pidx, eps, drop=None):
        assert drop is None, "the CPU stand-ins have no dropout: run the module in eval mode or with p = 0"
        H = word.shape[1]
        B, S = pidx.kind.shape
        text_vl = word[ids] + vo.layer_norm_tf(text_visual.float(), vt_w, vt_b, eps)
        obj_vl = object_vl[..., H:].float() + vo.layer_norm_tf(object_vl[..., :H].float(), vo_w, vo_b, eps)
        bidx = torch.arange(B).unsqueeze(1).expand(B, S)
        e = torch.zeros((B, S, H))
        e = torch.where((pidx.kind == 0).unsqueeze(-1), text_vl[bidx, pidx.src.clamp(max=text_vl.shape[1] - 1)], e)
        e = torch.where((pidx.kind == 1).unsqueeze(-1), obj_vl[bidx, pidx.src.clamp(max=obj_vl.shape[1] - 1)], e)
        e = torch.where((pidx.kind == 2).unsqueeze(-1), end[0].expand(B, S, H), e)
        e = e + pos[pidx.pos_id] + typ[pidx.type_id]
        out = vo.layer_norm_tf(e, ln_w, ln_b, eps)
        return out, out          # (bf16 GEMM operand, fp32 residual operand): one and the same tensor in the fp32 stand-in


_PER_LAYER = ("attention.self.query.weight", "attention.self.query.bias", "attention.self.key.weight", "attention.self.key.bias",
              "attention.self.value.weight", "attention.self.value.bias", "attention.output.dense.weight", "attention.output.dense.bias",
              "attention.output.LayerNorm.weight", "attention.output.LayerNorm.bias", "intermediate.dense.weight", "intermediate.dense.bias",
              "output.dense.weight", "output.dense.bias", "output.LayerNorm.weight", "output.LayerNorm.bias")


class _EncoderShim(object):
    @staticmethod
    def apply(emb, emb32, add_mask, meta, *params):
        assert len(params) == 16 * meta.L
        outs, h = [], emb
        am = add_mask.view(add_mask.shape[0], 1, 1, add_mask.shape[1])
        for l in range(meta.L):
            p = dict(zip(_PER_LAYER, params[16 * l: 16 * l + 16]))
            h = vo.bert_layer(h, am, p, meta.heads, meta.eps)
            if meta.all_layers or l == meta.L - 1:
                outs.append(h)
        return tuple(outs)


class _GatherRowsShim(object):
    @staticmethod
    def apply(src2d, idx, n_out):
        out = src2d.new_zeros((n_out, src2d.shape[1]))
        ok = idx >= 0
        return torch.where(ok.unsqueeze(-1), src2d[idx.clamp(min=0)], out)


class _LinearShim(object):
    @staticmethod
    def apply(x, weight, bias):
        return F.linear(x, weight, bias)


class _EncoderWeightsShim(object):
    def __init__(self, L, H, I, device):
        self.w_qkv = torch.zeros(1, device=device)


def install_encoder(monkeypatch):
    from vlbert_b200 import functional as VF
    monkeypatch.setattr(VF, "PackIndex", PackIndexShim)
    monkeypatch.setattr(VF, "EmbeddingFn", _EmbeddingShim)
    monkeypatch.setattr(VF, "EncoderFn", _EncoderShim)
    monkeypatch.setattr(VF, "GatherRowsFn", _GatherRowsShim)
    monkeypatch.setattr(VF, "EncoderWeights", _EncoderWeightsShim)
    monkeypatch.setattr(VF, "LinearFn", _LinearShim)
