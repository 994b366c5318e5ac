"""CPU: the host-side logic of vlbert_b200.VisualLinguisticBert / ...ForPretraining (packed length and its hint, parameter
flattening order, output structure, text/object split, heads) with the kernel entry points replaced by fp32 torch stand-ins
(tests/cpu_shim.py), against the reference fixtures -- and the drop-in boundary itself: the reference's own pre-training task
module (pretrain/modules/resnet_vlbert_for_pretraining.py) run twice on the same inputs and weights, once untouched and once
after `vlbert_b200.dropin.install()`."""
import os

import numpy as np
import pytest
import torch

import cpu_shim
import vlbert_oracle as vo
from synth import synth_vlbert_inputs, vlbert_loss


def _close(a, b, rel=2e-5, name=""):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (name, a.shape, b.shape)
    assert np.abs(a - b).max() <= rel * max(1e-6, np.abs(b).max()), (name, np.abs(a - b).max())


@pytest.fixture()
def tiny(monkeypatch, golden_dir):
    import vlbert_b200
    cpu_shim.install_encoder(monkeypatch)
    G = np.load(os.path.join(golden_dir, "vlbert_tiny.npz"))
    cfg = vo.default_config(vocab_size=200, hidden_size=128, num_hidden_layers=2, num_attention_heads=2,
                            intermediate_size=256, max_position_embeddings=64, visual_size=128)
    m = vlbert_b200.VisualLinguisticBert(cfg)
    m.load_state_dict({k[3:]: torch.from_numpy(G[k]) for k in G.files if k.startswith("sd.")}, strict=True)
    return m, G


def test_module_logic_against_reference_fixture(tiny):
    m, G = tiny
    ids, types, tvis, tmask, ovl, omask = synth_vlbert_inputs(B=3, T=9, R=5, H=128, vocab=200, seed=21)
    tvis = tvis.clone().requires_grad_(True)
    ovl = ovl.clone().requires_grad_(True)
    layers, pooled = m(ids, types, tvis, tmask, ovl, omask, output_all_encoded_layers=True)
    for i, l in enumerate(layers):
        _close(l.detach().numpy(), G["layer%d" % i], name="layer%d" % i)
    _close(pooled.detach().numpy(), G["pooled"], name="pooled")
    loss = vlbert_loss(layers, pooled, 31)
    m.zero_grad()
    loss.backward()
    _close(tvis.grad.numpy(), G["grad_text_visual"], name="grad_text_visual")
    _close(ovl.grad.numpy(), G["grad_object_vl"], name="grad_object_vl")
    for k, p in m.named_parameters():
        if p.grad is not None:
            _close(p.grad.numpy(), G["grad." + k], rel=5e-5, name=k)
    emb, mask, is_t, is_o = m.embedding(ids, types, tvis.detach(), tmask, ovl.detach(), omask)
    _close(emb.detach().numpy(), G["embedding"], name="embedding")
    assert np.array_equal(mask.numpy(), G["mask"]) and np.array_equal(is_t.numpy(), G["is_text"]) and np.array_equal(is_o.numpy(), G["is_object"])
    with torch.no_grad():
        tx, ob, _ = m(ids, types, tvis, tmask, ovl, omask, output_all_encoded_layers=False, output_text_and_object_separately=True)
        last, _ = m(ids, types, tvis, tmask, ovl, omask, output_all_encoded_layers=False)
    _close(tx.numpy(), G["split_text"], name="split_text")
    _close(ob.numpy(), G["split_object"], name="split_object")
    _close(last.numpy(), G["layer1"], name="last layer only")


def test_max_length_hint_only_adds_padding_rows(tiny):
    m, G = tiny
    inputs = synth_vlbert_inputs(B=3, T=9, R=5, H=128, vocab=200, seed=21)
    with torch.no_grad():
        ref, pooled_ref = m(*inputs, output_all_encoded_layers=False)
        m.max_length_hint = 9 + 5 + 1
        out, pooled = m(*inputs, output_all_encoded_layers=False)
    S = ref.shape[1]
    assert out.shape[1] == 15 and S <= 15
    valid = torch.from_numpy(G["mask"]).bool()
    _close(out[:, :S][valid].numpy(), ref[valid].numpy(), name="valid rows")
    _close(pooled.numpy(), pooled_ref.numpy(), name="pooled")


def test_pretraining_heads_against_reference_fixture(monkeypatch, golden_dir):
    import vlbert_b200
    cpu_shim.install_encoder(monkeypatch)
    G = np.load(os.path.join(golden_dir, "vlbert_tiny_pretrain_heads.npz"))
    cfg = vo.default_config(vocab_size=200, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
                            max_position_embeddings=64, visual_size=128, visual_region_classes=17, pos_embedding_frozen=False)
    m = vlbert_b200.VisualLinguisticBertForPretraining(cfg, None, True, True, True)
    m.load_state_dict({k[3:]: torch.from_numpy(G[k]) for k in G.files if k.startswith("sd.")}, strict=True)
    inputs = synth_vlbert_inputs(B=3, T=9, R=5, H=128, vocab=200, seed=62)
    with torch.no_grad():
        rel, mlm, mvrc = m(*inputs)
    _close(rel.numpy(), G["rel"], name="rel")
    _close(mlm.numpy(), G["mlm"], name="mlm")
    _close(mvrc.numpy(), G["mvrc"], name="mvrc")
