"""GPU parity of the fused masked-LM loss head (csrc/heads.cu + functional.MLMLossFn) against the torch head
of the same module (BertLMPredictionHead + F.cross_entropy(ignore_index=-1): the reference's formulation,
pretrain/modules/resnet_vlbert_for_pretraining.py:165-189) and against the reference-generated fixture
tests/golden/vlbert_tiny_pretrain_heads.npz.  Index work (compaction order, counts, arg-max) bit-exact; loss within 2e-3
(bf16 logits); gradients within 3e-2 relative L2 (the logits, hence softmax - onehot, are bf16-rounded before the GEMMs)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import vlbert_oracle as vo
from synth import synth_vlbert_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def test_label_compaction_is_exact():
    import vlbert_b200
    L = vlbert_b200._lib
    g = torch.Generator().manual_seed(0)
    for n, frac, cap in ((4096, 0.15, 1024), (77, 0.5, 80), (3000, 0.0, 16), (1500, 1.0, 1504), (5000, 0.3, 64)):
        labels = torch.where(torch.rand(n, generator=g) < frac, torch.randint(0, 30522, (n,), generator=g), torch.full((n,), -1))
        idx = torch.full((cap,), 7, dtype=torch.int32, device=DEV)
        lab = torch.full((cap,), 7, dtype=torch.int32, device=DEV)
        cnt = torch.zeros(1, dtype=torch.int32, device=DEV)
        L.check(L.lib().vlb_label_compact(labels.to(DEV).data_ptr(), n, -1, idx.data_ptr(), lab.data_ptr(), cap, cnt.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream))
        pos = (labels != -1).nonzero().flatten()
        assert int(cnt) == pos.numel()
        k = min(cap, pos.numel())
        assert torch.equal(idx.cpu()[:k].long(), pos[:k]) and torch.equal(lab.cpu()[:k].long(), labels[pos[:k]])
        assert bool((idx.cpu()[k:] == -1).all()) and bool((lab.cpu()[k:] == -1).all())


def _build(cfg, seed):
    import vlbert_b200
    torch.manual_seed(seed)
    m = vlbert_b200.VisualLinguisticBertForPretraining(cfg, None, False, True, False).to(DEV)
    with torch.no_grad():
        p = m.mlm_head.predictions
        p.bias.normal_(0, 0.02)
        p.transform.LayerNorm.weight.normal_(1, 0.1)
        p.transform.LayerNorm.bias.normal_(0, 0.05)
        p.transform.dense.bias.normal_(0, 0.02)
    return m


@pytest.mark.parametrize("shape", [dict(B=3, T=9, H=128, V=200, heads=2), dict(B=64, T=64, H=768, V=30522, heads=12)])
@pytest.mark.parametrize("hint", [False, True])
def test_fused_mlm_loss_against_the_torch_head(shape, hint):
    B, T, H, V = shape["B"], shape["T"], shape["H"], shape["V"]
    cfg = vo.default_config(vocab_size=V, hidden_size=H, num_hidden_layers=1, num_attention_heads=shape["heads"], intermediate_size=2 * H,
                            max_position_embeddings=128, visual_size=H, visual_region_classes=17)
    m = _build(cfg, 5)
    g = torch.Generator().manual_seed(6)
    text_out = torch.randn(B, T, H, generator=g).to(DEV)
    labels = torch.where(torch.rand(B, T, generator=g) < 0.15, torch.randint(0, V, (B, T), generator=g), torch.full((B, T), -1)).to(DEV)
    labels[0, 0] = V - 1                                   # the last vocabulary entry (next to the padding columns)
    params = {k: p for k, p in m.named_parameters() if k.startswith("mlm_head") or k == "word_embeddings.weight"}
    # reference formulation on the same module's torch head (fp32)
    t_ref = text_out.clone().requires_grad_(True)
    logits = m.mlm_head(t_ref)
    loss_ref = F.cross_entropy(logits.view(-1, V), labels.view(-1), ignore_index=-1)
    m.zero_grad()
    loss_ref.backward()
    g_ref = {k: p.grad.clone() for k, p in params.items()}
    keep = labels.view(-1) != -1
    acc_ref = int((logits.view(-1, V)[keep].argmax(1) == labels.view(-1)[keep]).sum())
    # fused
    t_our = text_out.clone().requires_grad_(True)
    m.zero_grad()
    loss, correct, count = m.mlm_loss(t_our, labels, max_labelled=(B * T if hint else None))
    (loss * 1.0).backward()
    assert int(count) == int(keep.sum())
    assert abs(float(loss) - float(loss_ref)) <= 2e-3 * abs(float(loss_ref)), (float(loss), float(loss_ref))
    assert abs(int(correct) - acc_ref) <= max(1, int(0.02 * int(keep.sum())))      # ties / bf16 near-ties may move a few arg-maxes
    errs = {k: rel(params[k].grad, g_ref[k]) for k in g_ref}
    errs["text_out"] = rel(t_our.grad, t_ref.grad)
    print("fused MLM loss %s hint=%s: loss %.6f (ref %.6f) correct %d (ref %d) grad errors %s" % (
        shape, hint, float(loss), float(loss_ref), int(correct), acc_ref, {k: "%.2e" % v for k, v in errs.items()}))
    assert all(v <= 3e-2 for v in errs.values()), errs
    assert bool((t_our.grad.view(-1, H)[~keep] == 0).all())                         # unlabelled positions get no gradient


def test_fused_mlm_loss_against_the_reference_fixture(golden_dir):
    """the reference's own mlm logits (tests/golden/vlbert_tiny_pretrain_heads.npz) -> cross-entropy on the host, against the
    fused loss computed from the library encoder's text output with the fixture's weights"""
    import vlbert_b200
    G = np.load(os.path.join(golden_dir, "vlbert_tiny_pretrain_heads.npz"))
    cfg = vo.default_config(vocab_size=200, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
                            max_position_embeddings=64, visual_size=128, visual_region_classes=17, pos_embedding_frozen=False)
    model = vlbert_b200.VisualLinguisticBertForPretraining(cfg, None, True, True, True).to(DEV)
    model.load_state_dict({k[3:]: torch.from_numpy(G[k]) for k in G.files if k.startswith("sd.")}, strict=True)
    inputs = [t.to(DEV) for t in synth_vlbert_inputs(B=3, T=9, R=5, H=128, vocab=200, seed=62)]
    with torch.no_grad():
        text_out, _, _ = vlbert_b200.VisualLinguisticBert.forward(model, *inputs, output_all_encoded_layers=False,
                                                                  output_text_and_object_separately=True)
    ref_logits = torch.from_numpy(G["mlm"])
    g = torch.Generator().manual_seed(9)
    labels = torch.where(torch.rand(3, 9, generator=g) < 0.4, torch.randint(0, 200, (3, 9), generator=g), torch.full((3, 9), -1))
    loss_ref = F.cross_entropy(ref_logits.view(-1, 200), labels.view(-1), ignore_index=-1)
    loss, correct, count = model.mlm_loss(text_out, labels.to(DEV))
    assert int(count) == int((labels != -1).sum())
    assert abs(float(loss) - float(loss_ref)) <= 1e-2 * abs(float(loss_ref)), (float(loss), float(loss_ref))
