"""Floating-point parity pinned the way north_star's "1e-3 relative for bf16" has to be read for a DEEP stack: one bf16
rounding is 2^-9 = 1.95e-3 (RMS 1.1e-3 relative), so no bf16 implementation -- the reference under torch.autocast included
-- can stay within 1e-3 of an fp32 run after dozens of chained GEMMs.  The measurable statement is

    err(CUDA path, fp32 oracle)  <=  C * err(bf16 comparator, fp32 oracle),     C = 1.25

for the outputs and for EVERY parameter gradient, where the comparator (oracle/vlbert_oracle.py: bf16_gemms) is the same
oracle graph with each GEMM evaluated like a bf16 tensor-core GEMM (operands and result rounded to bf16, fp32 accumulation,
forward and backward) and everything else in fp32.  Run at BASELINE config 2's FULL size (12 layers, batch 64, S = 101)
and at config 4's shape (VL-BERT-large width, S = 165) with 4 layers.  Absolute numbers are printed for DESIGN.md."""
import pytest
import torch

import vlbert_oracle as vo
from synth import seeded_state_dict, synth_vlbert_inputs, vlbert_loss

DEV = "cuda"
C_RATIO = 1.25


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _run(model, inputs, seed, dev):
    ids, types, tvis, tmask, ovl, omask = [t.to(dev) for t in inputs]
    tvis = tvis.clone().requires_grad_(True)
    ovl = ovl.clone().requires_grad_(True)
    layers, pooled = model(ids, types, tvis, tmask, ovl, omask, output_all_encoded_layers=True)
    loss = vlbert_loss(layers, pooled, seed)
    model.zero_grad()
    loss.backward()
    out = {"layer%02d" % i: l.detach() for i, l in enumerate(layers)}
    out["pooled"] = pooled.detach()
    out["grad:text_visual"] = tvis.grad
    out["grad:object_vl"] = ovl.grad
    for k, p in model.named_parameters():
        if p.grad is not None:
            out["grad:" + k] = p.grad.detach().clone()
    return out


def _errors(res, ref):
    e = {}
    for k, v in ref.items():
        if k.endswith("attention.self.key.bias"):
            # exactly zero in exact arithmetic (softmax shift invariance): measured against the sibling query.bias scale
            scale = ref[k.replace("key.bias", "query.bias")].double().norm().item()
            e[k] = (res[k].double().cpu() - v.double().cpu()).norm().item() / scale
        else:
            e[k] = rel(res[k], v)
    return e


def comparator_errors(cfg, sd, inputs, seed):
    ora = vo.VisualLinguisticBertOracle(cfg)
    ora.load_state_dict(sd)
    ref = _run(ora, inputs, seed, "cpu")
    with vo.bf16_gemms():
        cmp_ = _run(ora, inputs, seed, "cpu")
    return ref, _errors(cmp_, ref)


def test_comparator_is_a_bf16_sized_perturbation():
    """CPU: the yardstick itself -- per-GEMM bf16 rounding gives 1e-3 .. 1e-2 after two layers, nothing else changes."""
    cfg = vo.default_config(num_hidden_layers=2, vocab_size=500, max_position_embeddings=64)
    ora = vo.VisualLinguisticBertOracle(cfg)
    sd = seeded_state_dict(ora, 3)
    inputs = synth_vlbert_inputs(B=2, T=8, R=4, H=768, vocab=500, seed=4)
    ref, e = comparator_errors(cfg, sd, inputs, 5)
    assert all(2e-4 <= v <= 3e-2 for k, v in e.items() if not k.endswith("key.bias")), e
    assert not vo._GEMM_BF16          # the context manager restored full precision


def _check(cfg, B, T, R, seed):
    import vlbert_b200
    ora = vo.VisualLinguisticBertOracle(cfg)
    sd = seeded_state_dict(ora, seed)
    inputs = synth_vlbert_inputs(B=B, T=T, R=R, H=cfg.hidden_size, vocab=cfg.vocab_size, seed=seed + 1, ragged=True)
    ref, e_cmp = comparator_errors(cfg, sd, inputs, seed + 2)
    model = vlbert_b200.VisualLinguisticBert(cfg).to(DEV)
    model.load_state_dict(sd, strict=True)
    ours = _run(model, inputs, seed + 2, DEV)
    assert set(ours.keys()) == set(ref.keys())
    e_ours = _errors(ours, ref)
    rows = sorted(((e_ours[k] / max(e_cmp[k], 1e-12), k, e_ours[k], e_cmp[k]) for k in ref), reverse=True)
    outs = [r for r in rows if not r[1].startswith("grad:")]
    grads = [r for r in rows if r[1].startswith("grad:")]
    print("\nparity vs fp32 oracle (L=%d H=%d B=%d S<=%d): outputs worst %.2e (comparator %.2e), last layer %.2e (%.2e); "
          "gradients median %.2e (%.2e), worst %.2e (%.2e); worst ratio %.2f on %s" % (
              cfg.num_hidden_layers, cfg.hidden_size, B, T + R + 1, max(r[2] for r in outs), max(r[3] for r in outs),
              e_ours["layer%02d" % (cfg.num_hidden_layers - 1)], e_cmp["layer%02d" % (cfg.num_hidden_layers - 1)],
              sorted(r[2] for r in grads)[len(grads) // 2], sorted(r[3] for r in grads)[len(grads) // 2],
              max(r[2] for r in grads), max(r[3] for r in grads), rows[0][0], rows[0][1]))
    # attention.self.key.bias: the exact gradient is 0 (softmax shift invariance), so both sides measure pure rounding noise
    # of a 6000-row column sum against the sibling query.bias scale; it is bounded absolutely (1e-2 of that scale) instead of
    # by the ratio of two noise terms
    bad = [(k, "%.3e" % a, "%.3e" % b, "%.2f" % r) for r, k, a, b in rows
           if (a > 1e-2 if k.endswith("attention.self.key.bias") else a > C_RATIO * b + 1e-7)]
    assert not bad, bad


@pytest.mark.gpu
def test_config2_full_size_within_the_bf16_comparator():
    """BASELINE config 2 at full size: 12 layers, batch 64, 64 text + 36 region tokens (S = 101), ragged lengths."""
    _check(vo.default_config(num_hidden_layers=12), B=64, T=64, R=36, seed=101)


@pytest.mark.gpu
def test_config4_large_width_depth4_within_the_bf16_comparator():
    """BASELINE config 4 shape: VL-BERT-large width (H = 1024, 16 heads, I = 4096), 128 text + 36 regions (S = 165 > 128:
    multi-tile attention), 16 sequences (4 samples x 4 answer choices), 4 layers."""
    cfg = vo.default_config(num_hidden_layers=4, hidden_size=1024, num_attention_heads=16, intermediate_size=4096, visual_size=1024)
    _check(cfg, B=16, T=128, R=36, seed=202)


@pytest.mark.gpu
def test_config3_vqa_shape_depth12_within_the_bf16_comparator():
    """BASELINE config 3 token shape (20 text + 100 region tokens, S = 121), 12 layers, batch 16."""
    _check(vo.default_config(num_hidden_layers=12), B=16, T=20, R=100, seed=303)
