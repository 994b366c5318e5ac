"""TEST INFRASTRUCTURE ONLY -- portable fp32 CPU restatement of the reference hot path.

This file is the ORACLE the CUDA path is checked against on the GPU box (where /root/reference does
not exist).  It restates, in plain torch fp32 ops, the algorithms of

  * common/visual_linguistic_bert.py:95-241        VisualLinguisticBert.forward / .embedding
  * external/pytorch_pretrained_bert/modeling.py   gelu :114-120, BertLayerNorm :218-235,
        BertSelfAttention :290-319, BertSelfOutput :329-333, BertIntermediate :361-364,
        BertOutput :374-378, BertLayer :388-397, BertEncoder :406-421, BertPooler :430-436
  * common/fast_rcnn.py:128-203 (precomputed-feature path), common/utils/bbox.py:33-65,
        common/utils/pad_sequence.py:4-17

It is PINNED against the real reference modules: tests/test_oracle_vs_reference.py imports the
unmodified reference (oracle/ref_shim.py) in the build container and checks equality, and
oracle/make_golden.py writes reference-generated fixtures to tests/golden/ which the oracle is
re-checked against on every CPU test run (those fixtures travel; the reference does not).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module.  The product (vl-bert_b200/) never does.
"""
import math
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F


def default_config(**over):
    """NETWORK.VLBERT defaults for BERT-base (reference: pretrain/function/config.py:87-115)."""
    cfg = dict(
        vocab_size=30522, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
        intermediate_size=3072, hidden_act="gelu", hidden_dropout_prob=0.0,
        attention_probs_dropout_prob=0.0, max_position_embeddings=512, type_vocab_size=3,
        initializer_range=0.02, visual_size=768, visual_scale_text_init=1.0, visual_scale_object_init=1.0,
        visual_ln=True, word_embedding_frozen=False, with_pooler=True, position_padding_idx=-1,
        obj_pos_id_relative=True,
    )
    cfg.update(over)
    return SimpleNamespace(**cfg)


# ------------------------------------------------------------------------------------------------
# bf16 comparator: the same graph with every GEMM evaluated the way a bf16 tensor-core path (or the reference under
# torch.autocast(bfloat16)) evaluates it -- operands rounded to bf16, fp32 accumulation, result rounded to bf16, in forward
# AND backward -- and everything else (softmax, LayerNorm, GELU, residual adds) left in fp32.  It is the yardstick for
# the floating-point tolerance: err(CUDA path, fp32 oracle) is asserted against err(this, fp32 oracle).  Real autocast is
# strictly worse (it also runs the Python LayerNorm's mean/pow in bf16), so this is the stricter comparator.
# ------------------------------------------------------------------------------------------------
_GEMM_BF16 = False


class bf16_gemms(object):
    """context manager: `with vo.bf16_gemms(): ...` evaluates the oracle with bf16-rounded GEMMs"""

    def __enter__(self):
        global _GEMM_BF16
        self.prev = _GEMM_BF16
        _GEMM_BF16 = True

    def __exit__(self, *a):
        global _GEMM_BF16
        _GEMM_BF16 = self.prev


class _RoundBF16(torch.autograd.Function):
    """value and gradient both pass through a bf16 rounding (what a bf16 tensor crossing a GEMM boundary does)"""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


def _rb(x):
    return _RoundBF16.apply(x) if _GEMM_BF16 else x


def linear(x, w, b):
    if not _GEMM_BF16:
        return F.linear(x, w, b)
    return _rb(F.linear(_rb(x), _rb(w), b))


def matmul(a, b):
    if not _GEMM_BF16:
        return torch.matmul(a, b)
    return _rb(torch.matmul(_rb(a), _rb(b)))


# ------------------------------------------------------------------------------------------------
# primitives
# ------------------------------------------------------------------------------------------------
def gelu_erf(x):
    """modeling.py:120 -- exact erf form."""
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def layer_norm_tf(x, weight, bias, eps=1e-12):
    """modeling.py:231-235 -- biased variance, eps inside the sqrt."""
    mu = x.mean(-1, keepdim=True)
    var = (x - mu).pow(2).mean(-1, keepdim=True)
    return weight * ((x - mu) / torch.sqrt(var + eps)) + bias


class DropSpec(object):
    """Dropout of one forward pass of the oracle.
    mode "philox": the library's counter-based contract (oracle/philox.py) -- masks reproducible bit for bit, used by the
    parity tests; mode "torch": torch.nn.functional.dropout on the global generator, i.e. what the reference itself executes
    (modeling.py:283,310,331,376) -- used only to TIME the reference's CPU path with its dropout work included."""

    def __init__(self, p_hidden, p_attn, seed=0, step=0, mode="philox"):
        self.p_hidden, self.p_attn, self.seed, self.step, self.mode = float(p_hidden), float(p_attn), int(seed), int(step), mode

    def apply(self, x, p, site):
        """x: [..., cols]; the mask is indexed over x viewed as [rows, cols]"""
        if p <= 0.0:
            return x
        if self.mode == "torch":
            return F.dropout(x, p, training=True)
        import philox
        cols = x.shape[-1]
        rows = x.numel() // cols
        keep = torch.from_numpy(philox.keep_mask_2d(rows, cols, p, self.seed, site, self.step)).view(x.shape).to(x.device)
        scale = float(1.0 / (1.0 - float(torch.tensor(p, dtype=torch.float32))))
        return x * (keep.to(x.dtype) * scale)


def self_attention(x, add_mask, wq, bq, wk, bk, wv, bv, num_heads, drop=None, site=0):
    """modeling.py:290-315.  x [B,S,H]; add_mask [B,1,1,S] additive (0 / -10000); drop: DropSpec or None (p = 0)."""
    B, S, H = x.shape
    d = H // num_heads

    def split(t):
        return t.view(B, S, num_heads, d).permute(0, 2, 1, 3)

    q, k, v = split(linear(x, wq, bq)), split(linear(x, wk, bk)), split(linear(x, wv, bv))
    scores = matmul(q, k.transpose(-1, -2)) / math.sqrt(d) + add_mask
    probs = torch.softmax(scores, dim=-1)
    if drop is not None:
        probs = drop.apply(probs, drop.p_attn, site)            # modeling.py:310
    ctx = matmul(probs, v).permute(0, 2, 1, 3).contiguous().view(B, S, H)
    return ctx


def bert_layer(x, add_mask, p, num_heads, eps=1e-12, drop=None, layer=0):
    """modeling.py:388-397.  `p` maps the reference's per-layer state_dict suffixes to tensors.
    drop: DropSpec or None; sites 1+3l (attention probabilities), 2+3l (:331), 3+3l (:376)."""
    ctx = self_attention(x, add_mask,
                         p["attention.self.query.weight"], p["attention.self.query.bias"],
                         p["attention.self.key.weight"], p["attention.self.key.bias"],
                         p["attention.self.value.weight"], p["attention.self.value.bias"], num_heads, drop, 1 + 3 * layer)
    d = linear(ctx, p["attention.output.dense.weight"], p["attention.output.dense.bias"])
    if drop is not None:
        d = drop.apply(d, drop.p_hidden, 2 + 3 * layer)
    a = d + x
    h = layer_norm_tf(a, p["attention.output.LayerNorm.weight"], p["attention.output.LayerNorm.bias"], eps)
    u = gelu_erf(linear(h, p["intermediate.dense.weight"], p["intermediate.dense.bias"]))
    d = linear(u, p["output.dense.weight"], p["output.dense.bias"])
    if drop is not None:
        d = drop.apply(d, drop.p_hidden, 3 + 3 * layer)
    y = d + h
    return layer_norm_tf(y, p["output.LayerNorm.weight"], p["output.LayerNorm.bias"], eps)


def pack_indices(text_mask, object_mask):
    """Integer index math of the left-packed [text ; regions ; END ; pad] sequence
    (visual_linguistic_bert.py:200-227,235).  Returns int64 tensors, all [B, S]:
      kind   0 text, 1 region, 2 END, 3 pad
      src    source column inside text (kind 0) / object (kind 1) arrays, else 0
      pos    position id minus (position_padding_idx + 1)
      and text_end [B], object_end [B], S (python int).
    The k-th True of text_mask[b] lands at packed position k, exactly what the reference's
    boolean-mask assignment `vl[grid_pos < text_end] = text_vl[text_mask]` does."""
    B = text_mask.shape[0]
    dev = text_mask.device
    text_end = text_mask.sum(1)
    object_end = text_end + object_mask.sum(1)
    S = int(object_end.max().item()) + 1
    pos = torch.arange(S, dtype=torch.long, device=dev).unsqueeze(0).expand(B, S)
    te, oe = text_end.unsqueeze(1), object_end.unsqueeze(1)
    kind = torch.full((B, S), 3, dtype=torch.long, device=dev)
    kind[pos < te] = 0
    kind[(pos >= te) & (pos < oe)] = 1
    kind[pos == oe] = 2
    src = torch.zeros((B, S), dtype=torch.long, device=dev)
    for b in range(B):
        t_idx = torch.nonzero(text_mask[b], as_tuple=False).flatten()
        o_idx = torch.nonzero(object_mask[b], as_tuple=False).flatten()
        src[b, : t_idx.numel()] = t_idx
        src[b, t_idx.numel(): t_idx.numel() + o_idx.numel()] = o_idx
    pos_id = pos.clone()
    pos_id = torch.where(kind == 1, te.expand(B, S), pos_id)
    pos_id = torch.where(kind == 2, (te + 1).expand(B, S), pos_id)
    return kind, src, pos_id, text_end, object_end, S


class _LN(nn.Module):
    def __init__(self, n, eps=1e-12):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(n))
        self.bias = nn.Parameter(torch.zeros(n))
        self.eps = eps

    def forward(self, x):
        return layer_norm_tf(x, self.weight, self.bias, self.eps)


class _SelfAtt(nn.Module):
    def __init__(self, H):
        super().__init__()
        self.query, self.key, self.value = nn.Linear(H, H), nn.Linear(H, H), nn.Linear(H, H)


class _SelfOut(nn.Module):
    def __init__(self, I, H):
        super().__init__()
        self.dense = nn.Linear(I, H)
        self.LayerNorm = _LN(H)


class _Att(nn.Module):
    def __init__(self, H):
        super().__init__()
        self.self = _SelfAtt(H)
        self.output = _SelfOut(H, H)


class _Inter(nn.Module):
    def __init__(self, H, I):
        super().__init__()
        self.dense = nn.Linear(H, I)


class _Layer(nn.Module):
    def __init__(self, H, I):
        super().__init__()
        self.attention = _Att(H)
        self.intermediate = _Inter(H, I)
        self.output = _SelfOut(I, H)


class _Encoder(nn.Module):
    def __init__(self, L, H, I):
        super().__init__()
        self.layer = nn.ModuleList([_Layer(H, I) for _ in range(L)])


class _Pooler(nn.Module):
    def __init__(self, H):
        super().__init__()
        self.dense = nn.Linear(H, H)


class VisualLinguisticBertOracle(nn.Module):
    """Same constructor config, forward signature, return structure and state_dict keys as the
    reference VisualLinguisticBert (common/visual_linguistic_bert.py:31-171) for the configuration
    space the shipped cfgs use (visual_ln on, no word_embedding_frozen, obj_pos_id_relative)."""

    def __init__(self, config, language_pretrained_model_path=None):
        super().__init__()
        assert language_pretrained_model_path is None
        assert config.visual_ln and not config.word_embedding_frozen and config.obj_pos_id_relative
        self.config = config
        H = config.hidden_size
        self.word_embeddings = nn.Embedding(config.vocab_size, H)
        self.end_embedding = nn.Embedding(1, H)
        self.position_embeddings = nn.Embedding(config.max_position_embeddings, H)
        self.token_type_embeddings = nn.Embedding(config.type_vocab_size, H)
        self.embedding_LayerNorm = _LN(H)
        self.visual_1x1_text = None
        self.visual_1x1_object = None
        if config.visual_size != H:
            self.visual_1x1_text = nn.Linear(config.visual_size, H)
            self.visual_1x1_object = nn.Linear(config.visual_size, H)
        self.visual_ln_text = _LN(H)
        self.visual_ln_object = _LN(H)
        self.encoder = _Encoder(config.num_hidden_layers, H, config.intermediate_size)
        if config.with_pooler:
            self.pooler = _Pooler(H)
        self.position_padding_idx = config.position_padding_idx
        #: dropout of the next forward passes: None = off (also when in eval mode or the config's probabilities are 0);
        #: ("philox", seed, step) = the library's contract with that state; ("torch",) = torch's generator (timing only)
        self.dropout_state = None
        self.reset_parameters()

    def _drop_spec(self):
        cfg = self.config
        if self.dropout_state is None or not self.training:
            return None
        if cfg.hidden_dropout_prob == 0 and cfg.attention_probs_dropout_prob == 0:
            return None
        if self.dropout_state[0] == "torch":
            return DropSpec(cfg.hidden_dropout_prob, cfg.attention_probs_dropout_prob, mode="torch")
        _, seed, step = self.dropout_state
        return DropSpec(cfg.hidden_dropout_prob, cfg.attention_probs_dropout_prob, seed, step)

    def reset_parameters(self):
        std = self.config.initializer_range
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Embedding)):
                m.weight.data.normal_(mean=0.0, std=std)
            if isinstance(m, nn.Linear):
                m.bias.data.zero_()
            if isinstance(m, _LN):
                m.weight.data.fill_(1.0)
                m.bias.data.zero_()
        self.visual_ln_text.weight.data.fill_(self.config.visual_scale_text_init)
        self.visual_ln_object.weight.data.fill_(self.config.visual_scale_object_init)

    # -- embedding (visual_linguistic_bert.py:173-241) ---------------------------------------------
    def embedding(self, text_input_ids, text_token_type_ids, text_visual_embeddings, text_mask,
                  object_vl_embeddings, object_mask, drop="auto"):
        cfg = self.config
        if drop == "auto":
            drop = self._drop_spec()
        VS = cfg.visual_size
        tv = text_visual_embeddings
        ov = object_vl_embeddings[:, :, :VS]
        if self.visual_1x1_text is not None:
            tv = self.visual_1x1_text(tv)
            ov = self.visual_1x1_object(ov)
        text_vl = self.word_embeddings(text_input_ids) + self.visual_ln_text(tv)
        obj_vl = object_vl_embeddings[:, :, VS:] + self.visual_ln_object(ov)

        kind, src, pos_id, text_end, object_end, S = pack_indices(text_mask, object_mask)
        B, H = text_vl.shape[0], text_vl.shape[-1]
        bidx = torch.arange(B, device=text_vl.device).unsqueeze(1).expand(B, S)
        vl = text_vl.new_zeros((B, S, H))
        is_t, is_o, is_e = kind == 0, kind == 1, kind == 2
        vl[is_t] = text_vl[bidx[is_t], src[is_t]]
        vl[is_o] = obj_vl[bidx[is_o], src[is_o]]
        vl[is_e] = self.end_embedding.weight[0]
        type_ids = torch.zeros((B, S), dtype=torch.long, device=text_vl.device)
        type_ids[is_t] = text_token_type_ids[bidx[is_t], src[is_t]]
        type_ids[is_o | is_e] = 2
        position_ids = pos_id + self.position_padding_idx + 1
        emb = vl + self.position_embeddings(position_ids) + self.token_type_embeddings(type_ids)
        emb = self.embedding_LayerNorm(emb)
        if drop is not None:
            emb = drop.apply(emb, drop.p_hidden, 0)              # embedding_dropout, visual_linguistic_bert.py:239
        mask = (kind != 3).to(text_mask.dtype)
        return emb, mask, is_t, is_o

    def _layer_params(self, i):
        prefix = "encoder.layer.%d." % i
        sd = dict(self.named_parameters())
        return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}

    def forward(self, text_input_ids, text_token_type_ids, text_visual_embeddings, text_mask,
                object_vl_embeddings, object_mask, output_all_encoded_layers=True,
                output_text_and_object_separately=False, output_attention_probs=False):
        assert not output_attention_probs
        drop = self._drop_spec()
        emb, mask, is_t, is_o = self.embedding(text_input_ids, text_token_type_ids, text_visual_embeddings,
                                               text_mask, object_vl_embeddings, object_mask, drop)
        add_mask = (1.0 - mask.to(emb.dtype)).unsqueeze(1).unsqueeze(2) * -10000.0
        layers = []
        h = emb
        for i in range(self.config.num_hidden_layers):
            h = bert_layer(h, add_mask, self._layer_params(i), self.config.num_attention_heads, drop=drop, layer=i)
            layers.append(h)
        pooled = torch.tanh(linear(h[:, 0], self.pooler.dense.weight, self.pooler.dense.bias)) if self.config.with_pooler else None
        encoded = layers if output_all_encoded_layers else layers[-1]
        if not output_text_and_object_separately:
            return encoded, pooled
        T, R = text_input_ids.shape[1], object_vl_embeddings.shape[1]
        lst = encoded if output_all_encoded_layers else [encoded]
        texts, objs = [], []
        for e in lst:
            texts.append(e[:, :T])
            o = e.new_zeros((e.shape[0], R, e.shape[2]))
            o[object_mask] = e[is_o]
            objs.append(o)
        if not output_all_encoded_layers:
            texts, objs = texts[0], objs[0]
        return texts, objs, pooled


# ------------------------------------------------------------------------------------------------
# region-feature front end, precomputed-feature path
# ------------------------------------------------------------------------------------------------
def coordinate_embeddings(boxes6, dim=256):
    """common/utils/bbox.py:33-65.  boxes6 [K,6] = (x1,y1,x2,y2,W,H) -> [K,4,2*dim]."""
    x1, y1, x2, y2, W, Hh = boxes6.unbind(1)
    pos = torch.stack(((x1 + x2) / 2 / W * 100, (y1 + y2) / 2 / Hh * 100, (x2 - x1) / W * 100, (y2 - y1) / Hh * 100), 1)
    dim_mat = 1000 ** (torch.arange(dim, dtype=boxes6.dtype) / dim)
    arg = pos.unsqueeze(-1) / dim_mat.view(1, 1, -1)
    return torch.cat((arg.sin(), arg.cos()), dim=-1)


def fast_rcnn_precomputed(boxes, box_mask, im_info, weight, bias, mvrc_ops=None, mask_visual_embed=None, drop=None):
    """common/fast_rcnn.py:136-193 with IMAGE_FEAT_PRECOMPUTED, no classes/segms.
    boxes [B,R,4+2048]; weight [final_dim,4096]; returns (obj_reps [B,R,final_dim], obj_reps_raw [B,R,2048]).
    drop = (p, seed, step) applies obj_downsample's Dropout (:104-109) under the library's contract: the mask is indexed
    over the [B*R, 4096] SLOT layout (row b*R + r of the box the operand row was built from), site 1000."""
    B, R = box_mask.shape
    idx = box_mask.nonzero()
    assert idx.shape[0] > 0
    feats = boxes[idx[:, 0], idx[:, 1]][:, 4:]
    coords = boxes[idx[:, 0], idx[:, 1]][:, :4]
    if mvrc_ops is not None and mask_visual_embed is not None:
        feats = feats.clone()
        feats[(mvrc_ops == 1)[idx[:, 0], idx[:, 1]]] = mask_visual_embed
    ce = coordinate_embeddings(torch.cat((coords, im_info[idx[:, 0], :2]), 1), 256)
    x = torch.cat((ce.reshape(ce.shape[0], -1), feats), -1)
    if drop is not None and drop[0] > 0:
        import philox
        p_, seed_, step_ = drop
        keep = torch.from_numpy(philox.keep_mask_2d(B * R, x.shape[1], p_, seed_, 1000, step_))[idx[:, 0] * R + idx[:, 1]]
        x = x * (keep.to(x.dtype) * float(1.0 / (1.0 - float(torch.tensor(p_, dtype=torch.float32)))))
    final = torch.relu(F.linear(x, weight, bias))
    # pad_sequence (common/utils/pad_sequence.py): the k-th valid box of sample b lands in slot k
    slot = torch.cumsum(box_mask.long(), 1) - 1
    obj_reps = final.new_zeros((B, R, final.shape[1]))
    raw = feats.new_zeros((B, R, feats.shape[1]))
    obj_reps[idx[:, 0], slot[idx[:, 0], idx[:, 1]]] = final
    raw[idx[:, 0], slot[idx[:, 0], idx[:, 1]]] = feats
    return obj_reps, raw
