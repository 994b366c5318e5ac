"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by executing the UNMODIFIED reference
(/root/reference, via oracle/ref_shim.py) on seeded synthetic inputs.  Run in the build container:

    python oracle/make_golden.py

Fixtures (kept small enough to commit):
  vlbert_tiny.npz        VisualLinguisticBert, H=128 (2 heads x 64), I=256, L=2, vocab 200: full
                         state_dict + inputs + every output + gradients of every parameter/input.
  vlbert_c1_base.npz     BASELINE config 1 shape (base width 768/12 heads, L=2, B=2, T=8, R=4, S<=13):
                         weights are re-generated from a seed by the test (too big to commit);
                         outputs + small gradients + norms of the big gradients are stored.
  fastrcnn_prec.npz      FastRCNN precomputed-feature path (final_dim=32) incl. coordinate embeddings.
  fastrcnn_e2e.npz       FastRCNN end to end (ResNet-101 C4 + RoIAlign + dilated res5 head, final_dim=64) on 2 images of
                         128x160: weights are re-generated from a seed (oracle/frontend_oracle.py:synth_frontend_state);
                         outputs, a slice of body4 and slices + norms of the conv-weight gradients are stored.
  adamw.npz              4 steps of the reference AdamW (two param groups, warm-up lr, clip_grad_norm_ 1.0) on seeded tensors.
  roi_align_debug.npz    common/lib/roi_pooling/debug.py inputs + a 38x63 realistic case through the
                         reference's own CPU kernel (oracle/_ref, built by oracle/build_ref.py).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

GOLD = os.path.join(HERE, "..", "tests", "golden")


def synth_vlbert_inputs(B, T, R, H, vocab, seed, ragged=True):
    """Shared with the tests (tests/synth.py re-implements the same calls in the same order)."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(min(1000, vocab // 2), vocab, (B, T), generator=g)
    types = torch.randint(0, 2, (B, T), generator=g)
    tvis = torch.randn(B, T, H, generator=g)
    ovl = torch.randn(B, R, 2 * H, generator=g)
    tmask = torch.ones(B, T, dtype=torch.bool)
    omask = torch.ones(B, R, dtype=torch.bool)
    if ragged:
        for b in range(B):
            tl = int(torch.randint(max(1, T // 2), T + 1, (1,), generator=g))
            ol = int(torch.randint(1, R + 1, (1,), generator=g))
            if b == 0:
                tl, ol = T, R  # one full sample so S = T + R + 1
            tmask[b, tl:] = False
            omask[b, ol:] = False
        ids = ids * tmask
    return ids, types, tvis, tmask, ovl, omask


def seeded_state_dict(model, seed, std=0.02):
    """Deterministic weights: N(0, std) for matrices/embeddings, LayerNorm weight ~ 1 + N(0, .1),
    biases N(0, .02) -- non-trivial values everywhere so that no term hides behind a zero."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, v in model.state_dict().items():
        if "LayerNorm.weight" in k or k.startswith("visual_ln") and k.endswith("weight"):
            sd[k] = 1.0 + 0.1 * torch.randn(v.shape, generator=g)
        elif k.endswith("bias"):
            sd[k] = 0.02 * torch.randn(v.shape, generator=g)
        else:
            sd[k] = std * torch.randn(v.shape, generator=g)
    return sd


def run_vlbert(model, inputs, seed):
    ids, types, tvis, tmask, ovl, omask = inputs
    tvis = tvis.clone().requires_grad_(True)
    ovl = ovl.clone().requires_grad_(True)
    layers, pooled = model(ids, types, tvis, tmask, ovl, omask, output_all_encoded_layers=True,
                           output_text_and_object_separately=False)
    g = torch.Generator().manual_seed(seed + 1)
    gw = [torch.randn(l.shape, generator=g) for l in layers]
    gp = torch.randn(pooled.shape, generator=g)
    loss = sum((l * w).sum() for l, w in zip(layers, gw)) * 0.5 + (pooled * gp).sum()
    # heavier weight on the last layer, all layers get gradient (output_all_encoded_layers=True)
    loss = loss + (layers[-1] * gw[-1]).sum() * 0.5
    model.zero_grad()
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    return layers, pooled, loss.detach(), grads, tvis.grad, ovl.grad


def golden_vlbert_tiny():
    from common.visual_linguistic_bert import VisualLinguisticBert
    cfg = ref_shim.vlbert_config(vocab_size=200, hidden_size=128, num_hidden_layers=2, num_attention_heads=2,
                                 intermediate_size=256, max_position_embeddings=64, visual_size=128)
    torch.manual_seed(0)
    model = VisualLinguisticBert(cfg).eval()
    sd = seeded_state_dict(model, 11, std=0.05)
    model.load_state_dict(sd)
    inputs = synth_vlbert_inputs(B=3, T=9, R=5, H=128, vocab=200, seed=21)
    layers, pooled, loss, grads, gtv, gov = run_vlbert(model, inputs, 31)
    # also the text/object split outputs
    with torch.no_grad():
        tx, ob, _ = model(*inputs, output_all_encoded_layers=False, output_text_and_object_separately=True)
        emb, mask, is_t, is_o = model.embedding(*inputs)
    out = {"sd." + k: v.numpy() for k, v in sd.items()}
    out.update({"grad." + k: v.numpy() for k, v in grads.items()})
    for i, l in enumerate(layers):
        out["layer%d" % i] = l.detach().numpy()
    out.update(pooled=pooled.detach().numpy(), loss=loss.numpy(), grad_text_visual=gtv.numpy(), grad_object_vl=gov.numpy(),
               split_text=tx.numpy(), split_object=ob.numpy(), embedding=emb.numpy(), mask=mask.numpy(),
               is_text=is_t.numpy(), is_object=is_o.numpy())
    np.savez_compressed(os.path.join(GOLD, "vlbert_tiny.npz"), **out)
    print("vlbert_tiny: loss", float(loss), "S", layers[0].shape[1])


def golden_vlbert_c1():
    from common.visual_linguistic_bert import VisualLinguisticBert
    cfg = ref_shim.vlbert_config(num_hidden_layers=2)
    torch.manual_seed(0)
    model = VisualLinguisticBert(cfg).eval()
    sd = seeded_state_dict(model, 12)
    model.load_state_dict(sd)
    inputs = synth_vlbert_inputs(B=2, T=8, R=4, H=768, vocab=30522, seed=22)
    layers, pooled, loss, grads, gtv, gov = run_vlbert(model, inputs, 32)
    out = {}
    for i, l in enumerate(layers):
        out["layer%d" % i] = l.detach().numpy()
    out.update(pooled=pooled.detach().numpy(), loss=loss.numpy(), grad_text_visual=gtv.numpy(), grad_object_vl=gov.numpy())
    for k, v in grads.items():
        if v.numel() <= 4096:
            out["grad." + k] = v.numpy()
        else:
            out["gradnorm." + k] = np.float64(v.double().norm().item())
            out["gradhead." + k] = v.flatten()[:256].numpy()
    np.savez_compressed(os.path.join(GOLD, "vlbert_c1_base.npz"), **out)
    print("vlbert_c1_base: loss", float(loss), "S", layers[0].shape[1])


def golden_pretrain_heads():
    """VisualLinguisticBertForPretraining (rel / MLM / MVRC heads, common/visual_linguistic_bert.py:312-380) on the tiny config."""
    from common.visual_linguistic_bert import VisualLinguisticBertForPretraining
    cfg = ref_shim.vlbert_config(vocab_size=200, hidden_size=128, num_hidden_layers=2, num_attention_heads=2,
                                 intermediate_size=256, max_position_embeddings=64, visual_size=128,
                                 visual_region_classes=17, pos_embedding_frozen=False)
    torch.manual_seed(0)
    model = VisualLinguisticBertForPretraining(cfg, None, with_rel_head=True, with_mlm_head=True, with_mvrc_head=True).eval()
    sd = seeded_state_dict(model, 61, std=0.05)
    sd["mlm_head.predictions.decoder.weight"] = sd["word_embeddings.weight"]  # tied (modeling.py:463-466)
    model.load_state_dict(sd)
    inputs = synth_vlbert_inputs(B=3, T=9, R=5, H=128, vocab=200, seed=62)
    rel, mlm, mvrc = model(*inputs)
    g = torch.Generator().manual_seed(63)
    loss = (rel * torch.randn(rel.shape, generator=g)).sum() + (mlm * torch.randn(mlm.shape, generator=g)).sum() * 0.1 \
        + (mvrc * torch.randn(mvrc.shape, generator=g)).sum()
    model.zero_grad()
    loss.backward()
    out = {"sd." + k: v.numpy() for k, v in sd.items()}
    out.update(rel=rel.detach().numpy(), mlm=mlm.detach().numpy(), mvrc=mvrc.detach().numpy(), loss=loss.detach().numpy())
    for k, p_ in model.named_parameters():
        if p_.grad is not None and (p_.numel() <= 40000):
            out["grad." + k] = p_.grad.numpy()
    np.savez_compressed(os.path.join(GOLD, "vlbert_tiny_pretrain_heads.npz"), **out)
    print("vlbert_tiny_pretrain_heads: rel", tuple(rel.shape), "mlm", tuple(mlm.shape), "mvrc", tuple(mvrc.shape))


def golden_fastrcnn():
    from easydict import EasyDict
    from common.fast_rcnn import FastRCNN
    cfg = EasyDict({"NETWORK": {"IMAGE_FEAT_PRECOMPUTED": True, "IMAGE_SEMANTIC": False}})
    torch.manual_seed(0)
    m = FastRCNN(cfg, average_pool=True, final_dim=32, enable_cnn_reg_loss=False).eval()
    g = torch.Generator().manual_seed(41)
    B, R = 3, 6
    w = 0.02 * torch.randn(32, 4096, generator=g)
    b = 0.02 * torch.randn(32, generator=g)
    m.obj_downsample[1].weight.data.copy_(w)
    m.obj_downsample[1].bias.data.copy_(b)
    im_info = torch.tensor([[600., 400., 1, 1], [512., 384., 1, 1], [1000., 600., 1, 1]])
    x1 = torch.rand(B, R, generator=g) * im_info[:, None, 0] * 0.6
    y1 = torch.rand(B, R, generator=g) * im_info[:, None, 1] * 0.6
    x2 = x1 + 8 + torch.rand(B, R, generator=g) * im_info[:, None, 0] * 0.39
    y2 = y1 + 8 + torch.rand(B, R, generator=g) * im_info[:, None, 1] * 0.39
    feats = torch.randn(B, R, 2048, generator=g)
    boxes = torch.cat((torch.stack((x1, y1, x2, y2), -1), feats), -1)
    box_mask = torch.ones(B, R, dtype=torch.bool)
    box_mask[1, 4:] = False
    box_mask[2, 1:] = False
    boxes[~box_mask] = -2.0  # pretrain/data/collate_batch.py:39 pads with -2
    boxes_in = boxes.clone().requires_grad_(True)
    out = m(images=None, boxes=boxes_in, box_mask=box_mask, im_info=im_info)
    gw = torch.randn(out["obj_reps"].shape, generator=g)
    (out["obj_reps"] * gw).sum().backward()
    from common.utils.bbox import coordinate_embeddings
    idx = box_mask.nonzero()
    ce = coordinate_embeddings(torch.cat((boxes[idx[:, 0], idx[:, 1]][:, :4], im_info[idx[:, 0], :2]), 1), 256)
    np.savez_compressed(os.path.join(GOLD, "fastrcnn_prec.npz"), weight=w.numpy(), bias=b.numpy(), boxes=boxes.numpy(),
                        box_mask=box_mask.numpy(), im_info=im_info.numpy(), obj_reps=out["obj_reps"].detach().numpy(),
                        obj_reps_raw=out["obj_reps_raw"].detach().numpy(), grad_out=gw.numpy(),
                        grad_weight=m.obj_downsample[1].weight.grad.numpy(), grad_bias=m.obj_downsample[1].bias.grad.numpy(),
                        grad_boxes=boxes_in.grad.numpy(), coord_embed=ce.numpy())
    print("fastrcnn_prec: obj_reps", tuple(out["obj_reps"].shape))


E2E_GRAD_SLICES = (("obj_downsample.1.bias", None), ("obj_downsample.1.weight", 8),
                   ("roi_head_feature_extractor.2.conv3.weight", 32), ("roi_head_feature_extractor.0.conv2.weight", 4),
                   ("roi_head_feature_extractor.0.downsample.0.weight", 16), ("backbone.layer3.22.conv3.weight", 32),
                   ("backbone.layer3.0.conv2.weight", 8), ("backbone.layer2.0.conv1.weight", None),
                   ("backbone.layer2.0.conv2.weight", 8), ("backbone.layer2.0.downsample.0.weight", 16))


def synth_frontend_inputs(seed, B=2, R=4, H=128, W=160):
    """Shared with the tests (tests/synth.py:synth_frontend_inputs is call-for-call identical)."""
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(B, 3, H, W, generator=g)
    im_info = torch.tensor([[float(W), float(H), 1.0, 1.0]] * B)
    x1 = torch.rand(B, R, generator=g) * W * 0.55
    y1 = torch.rand(B, R, generator=g) * H * 0.55
    x2 = x1 + 12 + torch.rand(B, R, generator=g) * W * 0.4
    y2 = y1 + 12 + torch.rand(B, R, generator=g) * H * 0.4
    boxes = torch.stack((x1, y1, x2.clamp(max=W - 1), y2.clamp(max=H - 1)), -1)
    box_mask = torch.ones(B, R, dtype=torch.bool)
    box_mask[1, R - 1:] = False
    boxes[~box_mask] = -2.0
    grad_out = torch.randn(B, R, 64, generator=g)
    return images, boxes, box_mask, im_info, grad_out


def golden_fastrcnn_e2e():
    sys.path.insert(0, os.path.join(HERE, ".."))
    from oracle.frontend_oracle import synth_frontend_state
    import warnings
    import torch.utils.model_zoo as model_zoo
    from easydict import EasyDict
    model_zoo.load_url = lambda *a, **k: {}          # no network; every tensor is overwritten below
    from common.fast_rcnn import FastRCNN
    cfg = EasyDict({"NETWORK": dict(IMAGE_FEAT_PRECOMPUTED=False, IMAGE_SEMANTIC=False, IMAGE_STRIDE_IN_1x1=True, IMAGE_C5_DILATED=True,
                                    IMAGE_NUM_LAYERS=101, IMAGE_PRETRAINED="", IMAGE_PRETRAINED_EPOCH=0, OUTPUT_CONV5=False,
                                    IMAGE_FROZEN_BN=True, IMAGE_FROZEN_BACKBONE_STAGES=[1, 2])})
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = FastRCNN(cfg, average_pool=True, final_dim=64, enable_cnn_reg_loss=False)
    m.load_state_dict(synth_frontend_state({k: v.shape for k, v in m.state_dict().items()}, 77), strict=True)
    m.eval()                                          # dropout off; BN on its frozen statistics (FastRCNN.bn_eval)
    images, boxes, box_mask, im_info, gw = synth_frontend_inputs(78)
    body4 = m.backbone(images)["body4"]
    out = m(images=images, boxes=boxes, box_mask=box_mask, im_info=im_info)
    (out["obj_reps"] * gw).sum().backward()
    params = dict(m.named_parameters())
    res = dict(obj_reps=out["obj_reps"].detach().numpy(), obj_reps_raw=out["obj_reps_raw"].detach().numpy(),
               body4_slice=body4.detach()[:, ::8].numpy())
    for name, rows in E2E_GRAD_SLICES:
        gfull = params[name].grad
        res["grad:" + name] = (gfull if rows is None else gfull[:rows]).numpy()
        res["gnorm:" + name] = np.float64(gfull.double().norm())
    np.savez_compressed(os.path.join(GOLD, "fastrcnn_e2e.npz"), **res)
    print("fastrcnn_e2e: obj_reps", tuple(out["obj_reps"].shape), "|raw|", float(out["obj_reps_raw"].abs().mean()),
          "|body4|", float(body4.abs().mean()))


def adamw_case(seed=5):
    """Shared with the tests: three parameter tensors in two groups (decay / no decay, different lr), 4 steps of seeded
    gradients, a linear-warmup learning rate and global-norm clipping at 1.0."""
    g = torch.Generator().manual_seed(seed)
    shapes = [(37, 19), (19,), (8, 3, 5)]
    params = [torch.randn(s, generator=g) for s in shapes]
    grads = [[torch.randn(s, generator=g) * (0.5 + k) for s in shapes] for k in range(4)]
    groups = [dict(idx=[0, 2], lr=2e-3, weight_decay=0.01), dict(idx=[1], lr=1e-3, weight_decay=0.0)]
    lr_scale = [0.25, 0.5, 0.75, 1.0]
    return params, grads, groups, lr_scale


def golden_adamw():
    from common.nlp.bert.optimization import AdamW
    import warnings
    params, grads, groups, lr_scale = adamw_case()
    ps = [torch.nn.Parameter(p.clone()) for p in params]
    opt = AdamW([dict(params=[ps[i] for i in g["idx"]], lr=g["lr"], weight_decay=g["weight_decay"]) for g in groups],
                lr=1e-3, betas=(0.9, 0.999), eps=1e-6)
    base = [g["lr"] for g in groups]
    out = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for k in range(4):
            for grp, b in zip(opt.param_groups, base):
                grp["lr"] = b * lr_scale[k]
            for p, gr in zip(ps, grads[k]):
                p.grad = gr.clone()
            out["norm%d" % k] = np.float64(torch.nn.utils.clip_grad_norm_(ps, 1.0))     # common/trainer.py:139-147
            opt.step()
            for i, p in enumerate(ps):
                out["p%d_step%d" % (i, k)] = p.detach().numpy().copy()
    for i, p in enumerate(ps):
        out["m%d" % i] = opt.state[p]["exp_avg"].numpy().copy()
        out["v%d" % i] = opt.state[p]["exp_avg_sq"].numpy().copy()
    np.savez_compressed(os.path.join(GOLD, "adamw.npz"), **out)
    print("adamw: 4 steps, norms", [float(out["norm%d" % k]) for k in range(4)])


def golden_roi_align():
    import build_ref
    ref = build_ref.load()
    out = {}
    f = torch.arange(81 * 2 * 3).view(2, 3, 9, 9).float()
    rois = torch.tensor([[0, 0, 0, 9, 9], [1, 0, 0, 9, 9], [1, 0, 0, 7, 7]]).float()
    out["debug_feature"], out["debug_rois"] = f.numpy(), rois.numpy()
    for sr in (1, 2, 0):
        out["debug_out_sr%d" % sr] = ref.roi_align_forward(f, rois, 1.0, 3, 3, sr).numpy()
    g = torch.Generator().manual_seed(51)
    f = torch.randn(2, 16, 38, 63, generator=g)
    K = 24
    x1 = torch.rand(K, generator=g) * 800
    y1 = torch.rand(K, generator=g) * 500
    w = torch.rand(K, generator=g) * 400 + 2
    h = torch.rand(K, generator=g) * 300 + 2
    rois = torch.stack([torch.randint(0, 2, (K,), generator=g).float(), x1, y1, x1 + w, y1 + h], 1)
    rois[0] = torch.tensor([0, -40., -30., 20., 10.])      # partially outside
    rois[1] = torch.tensor([1, 100., 100., 100.2, 100.1])  # degenerate -> clamped to 1x1 (ROIAlign_cuda.cu:92-93)
    rois[2] = torch.tensor([1, 990., 590., 1100., 700.])   # beyond the 1008x608 extent
    out["real_feature"], out["real_rois"] = f.numpy(), rois.numpy()
    out["real_out_sr1"] = ref.roi_align_forward(f, rois, 1.0 / 16, 14, 14, 1).numpy()
    out["real_out_sr2"] = ref.roi_align_forward(f, rois, 1.0 / 16, 14, 14, 2).numpy()
    np.savez_compressed(os.path.join(GOLD, "roi_align_debug.npz"), **out)
    print("roi_align_debug: from the reference's own CPU kernel (oracle/_ref)")


if __name__ == "__main__":
    ref_shim.install()
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    golden_vlbert_tiny()
    golden_vlbert_c1()
    golden_pretrain_heads()
    golden_fastrcnn()
    golden_fastrcnn_e2e()
    golden_adamw()
    golden_roi_align()
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)) // 1024, "KiB")
