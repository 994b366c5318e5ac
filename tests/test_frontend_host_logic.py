"""CPU: the host-side logic of the end-to-end front end (vlbert_b200.resnet, vlbert_b200.modules.FastRCNN) with the kernel
entry points replaced by fp32 torch stand-ins (tests/cpu_shim.py): every structural decision the Python layer makes -- stride /
dilation placement, residual wiring, frozen stages, RoI batching in both `compact_rois` modes, slot re-padding, MVRC masking,
the CNN regularisation head -- against the oracle and the reference fixture.  fp32 both sides: 1e-4 of max."""
import os

import numpy as np
import pytest
import torch

import cpu_shim
import frontend_oracle as fo
from synth import E2E_GRAD_SLICES, frontend_config, synth_frontend_inputs


def _close(a, b, rel=1e-4, name=""):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (name, a.shape, b.shape)
    assert np.abs(a - b).max() <= rel * max(1e-6, np.abs(b).max()), (name, np.abs(a - b).max(), np.abs(b).max())


@pytest.fixture()
def model(monkeypatch):
    import vlbert_b200
    cpu_shim.install(monkeypatch)
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    m = vlbert_b200.FastRCNN(frontend_config(101), average_pool=True, final_dim=64, enable_cnn_reg_loss=False)
    sd = fo.synth_frontend_state({k: v.shape for k, v in m.state_dict().items()}, 77)
    m.load_state_dict(sd, strict=True)
    return m.eval(), sd


@pytest.mark.parametrize("compact", [True, False])
def test_fastrcnn_module_logic_against_reference_fixture(model, golden_dir, compact):
    m, _ = model
    G = np.load(os.path.join(golden_dir, "fastrcnn_e2e.npz"))
    m.compact_rois = compact
    images, boxes, box_mask, im_info, gw = synth_frontend_inputs(78)
    body4 = m.backbone(images)["body4"].permute(0, 3, 1, 2)
    _close(body4.detach()[:, ::8].numpy(), G["body4_slice"], name="body4")
    out = m(images=images, boxes=boxes, box_mask=box_mask, im_info=im_info)
    _close(out["obj_reps"].detach().numpy(), G["obj_reps"], name="obj_reps")
    _close(out["obj_reps_raw"].detach().numpy(), G["obj_reps_raw"], name="obj_reps_raw")
    m.zero_grad()
    (out["obj_reps"] * gw).sum().backward()
    params = dict(m.named_parameters())
    for name, rows in E2E_GRAD_SLICES:
        g = params[name].grad
        _close((g if rows is None else g[:rows]).numpy(), G["grad:" + name], rel=3e-4, name=name)
    assert all(p.grad is None for n, p in params.items() if not p.requires_grad)


def test_mvrc_masking_and_unordered_mask(model):
    """mask_visual_embed replaces the pooled feature of masked regions (common/fast_rcnn.py:166-168); a box_mask with a hole
    re-pads the valid boxes to the front (pad_sequence, :177-178)"""
    m, sd = model
    images, boxes, box_mask, im_info, _ = synth_frontend_inputs(5, B=2, R=3, H=96, W=128)
    box_mask[0, 1] = False                      # hole in the middle
    box_mask[1, 2] = True
    boxes = boxes.abs() + 2.0
    boxes[..., 2:] = boxes[..., :2] + 25
    mvrc_ops = torch.zeros(2, 3, dtype=torch.long)
    mvrc_ops[0, 2] = 1
    mvrc_ops[1, 0] = 1
    emb = torch.randn(2048, generator=torch.Generator().manual_seed(1))
    for compact in (True, False):
        m.compact_rois = compact
        out = m(images=images, boxes=boxes, box_mask=box_mask, im_info=im_info, mvrc_ops=mvrc_ops, mask_visual_embed=emb)
        ref_obj, ref_raw = fo.fast_rcnn_end2end(sd, images, boxes, box_mask, im_info, mvrc_ops=mvrc_ops, mask_visual_embed=emb)
        _close(out["obj_reps"].detach().numpy(), ref_obj.detach().numpy(), name="obj_reps")


def test_cnn_regularisation_head(monkeypatch):
    """enable_cnn_reg_loss: logits / labels / loss over the valid boxes (common/fast_rcnn.py:160-163)"""
    import vlbert_b200
    cpu_shim.install(monkeypatch)
    torch.manual_seed(0)
    m = vlbert_b200.FastRCNN(frontend_config(50), average_pool=True, final_dim=32, enable_cnn_reg_loss=True).eval()
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.weight.fill_(0.3)
    images, boxes, box_mask, im_info, _ = synth_frontend_inputs(9, B=2, R=3, H=64, W=96)
    classes = torch.randint(0, 81, box_mask.shape)
    outs = []
    for compact in (True, False):
        m.compact_rois = compact
        outs.append(m(images=images, boxes=boxes, box_mask=box_mask, im_info=im_info, classes=classes))
    K = int(box_mask.sum())
    for o in outs:
        assert o["obj_logits"].shape == (K, 81) and torch.equal(o["obj_labels"], classes[box_mask])
        ce = torch.nn.functional.cross_entropy(o["obj_logits"], o["obj_labels"])
        assert torch.allclose(o["cnn_regularization_loss"], ce[None])
    assert torch.allclose(outs[0]["obj_logits"], outs[1]["obj_logits"], atol=1e-4, rtol=1e-4)
    assert torch.allclose(outs[0]["obj_reps"], outs[1]["obj_reps"], atol=1e-4, rtol=1e-4)


def test_instance_mask_weighted_pooling(model):
    """VCR passes per-box 14x14 instance masks (`segms`): the res5 map is multiplied by the mask before the mean pool
    (common/fast_rcnn.py:151-156, vcr/modules/resnet_vlbert_for_vcr.py:244-261)"""
    m, sd = model
    images, boxes, box_mask, im_info, gw = synth_frontend_inputs(11, B=2, R=3, H=96, W=128)
    box_mask[1, 2] = False
    g = torch.Generator().manual_seed(4)
    segms = (torch.rand(2, 3, 14, 14, generator=g) > 0.4).float()
    segms[0, 0] = 1.0                                                  # the whole-image box of VCR
    ref_obj, ref_raw = fo.fast_rcnn_end2end(sd, images, boxes, box_mask, im_info, segms=segms)
    plain_obj, _ = fo.fast_rcnn_end2end(sd, images, boxes, box_mask, im_info)
    assert not torch.allclose(ref_obj, plain_obj)
    for compact in (True, False):
        m.compact_rois = compact
        out = m(images=images, boxes=boxes, box_mask=box_mask, im_info=im_info, segms=segms)
        _close(out["obj_reps"].detach().numpy(), ref_obj.detach().numpy(), name="obj_reps")
        _close(out["obj_reps_raw"].detach().numpy(), ref_raw.detach().numpy(), name="raw")
    with pytest.raises(ValueError):
        m(images=images, boxes=boxes, box_mask=box_mask, im_info=im_info, segms=segms[:, :, :7, :7])
