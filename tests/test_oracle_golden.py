"""CPU: pins the portable oracle (oracle/vlbert_oracle.py, oracle/roi_align_oracle.c) against the
golden fixtures that oracle/make_golden.py produced by executing the UNMODIFIED reference.
Tolerance: the oracle and the reference are both fp32 on CPU and differ only in op order, so
elementwise |diff| <= 2e-5 * max|ref| (fp32 re-association), bit-exact for index outputs."""
import os

import numpy as np
import pytest
import torch

import roi_align as roi_oracle
import vlbert_oracle as vo
import frontend_oracle as fo
from synth import (E2E_GRAD_SLICES, frontend_shapes, seeded_state_dict, synth_frontend_inputs, synth_vlbert_inputs, vlbert_loss)


def _close(a, b, rel=2e-5, name=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (name, a.shape, b.shape)
    tol = rel * max(1e-6, np.abs(b).max())
    assert np.abs(a - b).max() <= tol, (name, np.abs(a - b).max(), tol)


def _run(model, inputs, seed):
    ids, types, tvis, tmask, ovl, omask = inputs
    tvis = tvis.clone().requires_grad_(True)
    ovl = ovl.clone().requires_grad_(True)
    layers, pooled = model(ids, types, tvis, tmask, ovl, omask, output_all_encoded_layers=True)
    loss = vlbert_loss(layers, pooled, seed)
    model.zero_grad()
    loss.backward()
    grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    return layers, pooled, loss, grads, tvis.grad, ovl.grad


def test_tiny_forward_backward_matches_reference_fixture(golden_dir):
    G = np.load(os.path.join(golden_dir, "vlbert_tiny.npz"))
    cfg = vo.default_config(vocab_size=200, hidden_size=128, num_hidden_layers=2, num_attention_heads=2,
                            intermediate_size=256, max_position_embeddings=64, visual_size=128)
    model = vo.VisualLinguisticBertOracle(cfg)
    sd = {k[3:]: torch.from_numpy(G[k]) for k in G.files if k.startswith("sd.")}
    assert set(sd.keys()) == set(model.state_dict().keys())  # checkpoint ABI
    model.load_state_dict(sd, strict=True)
    inputs = synth_vlbert_inputs(B=3, T=9, R=5, H=128, vocab=200, seed=21)
    layers, pooled, loss, grads, gtv, gov = _run(model, inputs, 31)
    for i, l in enumerate(layers):
        _close(l.detach().numpy(), G["layer%d" % i], name="layer%d" % i)
    _close(pooled.detach().numpy(), G["pooled"], name="pooled")
    _close(loss.detach().numpy(), G["loss"], rel=1e-5, name="loss")
    _close(gtv.numpy(), G["grad_text_visual"], name="grad_text_visual")
    _close(gov.numpy(), G["grad_object_vl"], name="grad_object_vl")
    for k, g in grads.items():
        _close(g.numpy(), G["grad." + k], rel=5e-5, name=k)
    # embedding stage + masks are index work: masks bit-exact
    emb, mask, is_t, is_o = model.embedding(*inputs)
    _close(emb.detach().numpy(), G["embedding"], name="embedding")
    assert np.array_equal(mask.numpy(), G["mask"])
    assert np.array_equal(is_t.numpy(), G["is_text"])
    assert np.array_equal(is_o.numpy(), G["is_object"])
    with torch.no_grad():
        tx, ob, _ = model(*inputs, output_all_encoded_layers=False, output_text_and_object_separately=True)
    _close(tx.numpy(), G["split_text"], name="split_text")
    _close(ob.numpy(), G["split_object"], name="split_object")


def test_c1_base_width_matches_reference_fixture(golden_dir):
    """BASELINE config 1 (2 layers, 8 text + 4 region tokens, batch 2) at base width."""
    G = np.load(os.path.join(golden_dir, "vlbert_c1_base.npz"))
    cfg = vo.default_config(num_hidden_layers=2)
    model = vo.VisualLinguisticBertOracle(cfg)
    model.load_state_dict(seeded_state_dict(model, 12))
    inputs = synth_vlbert_inputs(B=2, T=8, R=4, H=768, vocab=30522, seed=22)
    layers, pooled, loss, grads, gtv, gov = _run(model, inputs, 32)
    for i, l in enumerate(layers):
        _close(l.detach().numpy(), G["layer%d" % i], name="layer%d" % i)
    _close(pooled.detach().numpy(), G["pooled"], name="pooled")
    _close(gtv.numpy(), G["grad_text_visual"], rel=5e-5)
    _close(gov.numpy(), G["grad_object_vl"], rel=5e-5)
    for k, g in grads.items():
        if "grad." + k in G.files:
            _close(g.numpy(), G["grad." + k], rel=1e-4, name=k)
        else:
            ref = float(G["gradnorm." + k])
            assert abs(g.double().norm().item() - ref) <= 1e-4 * max(ref, 1e-9), k
            _close(g.flatten()[:256].numpy(), G["gradhead." + k], rel=1e-4, name=k)


def test_fastrcnn_precomputed_matches_reference_fixture(golden_dir):
    G = np.load(os.path.join(golden_dir, "fastrcnn_prec.npz"))
    boxes = torch.from_numpy(G["boxes"]).requires_grad_(True)
    w = torch.from_numpy(G["weight"]).requires_grad_(True)
    b = torch.from_numpy(G["bias"]).requires_grad_(True)
    box_mask = torch.from_numpy(G["box_mask"])
    im_info = torch.from_numpy(G["im_info"])
    obj, raw = vo.fast_rcnn_precomputed(boxes, box_mask, im_info, w, b)
    _close(obj.detach().numpy(), G["obj_reps"], name="obj_reps")
    assert np.array_equal(raw.detach().numpy(), G["obj_reps_raw"])  # pure gather: bit-exact
    (obj * torch.from_numpy(G["grad_out"])).sum().backward()
    _close(w.grad.numpy(), G["grad_weight"], rel=5e-5)
    _close(b.grad.numpy(), G["grad_bias"], rel=5e-5)
    _close(boxes.grad.numpy(), G["grad_boxes"], rel=5e-5)
    idx = box_mask.nonzero()
    ce = vo.coordinate_embeddings(torch.cat((boxes.detach()[idx[:, 0], idx[:, 1]][:, :4], im_info[idx[:, 0], :2]), 1))
    _close(ce.numpy(), G["coord_embed"], rel=1e-5)


def test_frontend_oracle_matches_reference_fixture(golden_dir):
    """ResNet-101 C4 + RoIAlign + res5 + projection, forward and backward, against the reference's own run
    (oracle/make_golden.py:golden_fastrcnn_e2e).  fp32 both sides: 1e-4 of the tensor's max (conv re-association)."""
    torch.set_num_threads(8)
    G = np.load(os.path.join(golden_dir, "fastrcnn_e2e.npz"))
    sd = fo.synth_frontend_state(frontend_shapes(), 77)
    names = [n for n, _ in E2E_GRAD_SLICES]
    for n in names:
        sd[n] = sd[n].clone().requires_grad_(True)
    images, boxes, box_mask, im_info, gw = synth_frontend_inputs(78)
    body4 = fo.resnet_c4(sd, images)
    _close(body4.detach()[:, ::8].numpy(), G["body4_slice"], rel=1e-4, name="body4")
    obj, raw = fo.fast_rcnn_end2end(sd, images, boxes, box_mask, im_info)
    _close(obj.detach().numpy(), G["obj_reps"], rel=1e-4, name="obj_reps")
    _close(raw.detach().numpy(), G["obj_reps_raw"], rel=1e-4, name="raw")
    (obj * gw).sum().backward()
    for n, rows in E2E_GRAD_SLICES:
        g = sd[n].grad
        _close((g if rows is None else g[:rows]).numpy(), G["grad:" + n], rel=2e-4, name=n)
        assert abs(float(g.double().norm()) / float(G["gnorm:" + n]) - 1) < 1e-4, n


def test_roi_align_c_oracle_matches_reference_cpu_kernel_fixture(golden_dir):
    """Fixture came from the reference's own cpu/ROIAlign_cpu.cpp (oracle/_ref): bit-exact."""
    G = np.load(os.path.join(golden_dir, "roi_align_debug.npz"))
    for sr in (1, 2, 0):
        out = roi_oracle.roi_align_forward(G["debug_feature"], G["debug_rois"], 1.0, 3, 3, sr)
        assert np.array_equal(out, G["debug_out_sr%d" % sr]), sr
    for sr in (1, 2):
        out = roi_oracle.roi_align_forward(G["real_feature"], G["real_rois"], 1.0 / 16, 14, 14, sr)
        assert np.array_equal(out, G["real_out_sr%d" % sr]), sr


def test_roi_align_backward_oracle_is_adjoint_of_forward():
    """<forward(x), g> == <x, backward(g)> (linearity/adjoint property, size-independent)."""
    rng = np.random.RandomState(0)
    x = rng.randn(2, 4, 20, 30).astype(np.float32)
    rois = np.array([[0, 10, 20, 200, 150], [1, 0, 0, 479, 319], [1, 300, 100, 310, 104], [0, -20, -20, 40, 40]], np.float32)
    g = rng.randn(4, 4, 7, 7).astype(np.float32)
    y = roi_oracle.roi_align_forward(x, rois, 1 / 16.0, 7, 7, 2)
    gx = roi_oracle.roi_align_backward(g, rois, 1 / 16.0, 7, 7, 2, 4, 20, 30, 2)
    lhs = float((y.astype(np.float64) * g).sum())
    rhs = float((x.astype(np.float64) * gx).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs))


def test_pack_indices_edge_cases():
    """Ragged / empty-ish inputs of the packing index math (bit-exact integer work)."""
    tmask = torch.tensor([[1, 1, 1, 0], [1, 0, 0, 0], [1, 1, 1, 1]], dtype=torch.bool)
    omask = torch.tensor([[1, 0, 0], [1, 1, 1], [0, 0, 0]], dtype=torch.bool)
    kind, src, pos, te, oe, S = vo.pack_indices(tmask, omask)
    assert S == 5
    assert kind.tolist() == [[0, 0, 0, 1, 2], [0, 1, 1, 1, 2], [0, 0, 0, 0, 2]]
    assert pos.tolist() == [[0, 1, 2, 3, 4], [0, 1, 1, 1, 2], [0, 1, 2, 3, 5]]
    assert src[1].tolist()[:4] == [0, 0, 1, 2]
