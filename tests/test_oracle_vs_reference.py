"""CPU, build container only: the oracle against the LIVE unmodified reference (skipped where
/root/reference is absent, e.g. on the GPU box -- there the committed fixtures pin it)."""
import os
import sys

import numpy as np
import pytest
import torch

import ref_shim
import vlbert_oracle as vo
from synth import seeded_state_dict, synth_vlbert_inputs

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")


@pytest.mark.parametrize("ragged", [False, True])
def test_oracle_equals_reference_module(ragged):
    ref_shim.install()
    from common.visual_linguistic_bert import VisualLinguisticBert
    kw = dict(vocab_size=300, hidden_size=128, num_hidden_layers=3, num_attention_heads=2, intermediate_size=512,
              max_position_embeddings=80, visual_size=128, with_pooler=True)
    torch.manual_seed(0)
    ref = VisualLinguisticBert(ref_shim.vlbert_config(**kw)).eval()
    ora = vo.VisualLinguisticBertOracle(vo.default_config(**kw)).eval()
    sd = seeded_state_dict(ref, 5, std=0.05)
    ref.load_state_dict(sd)
    ora.load_state_dict(sd, strict=True)
    inputs = synth_vlbert_inputs(B=4, T=12, R=7, H=128, vocab=300, seed=9, ragged=ragged)
    with torch.no_grad():
        a = ref(*inputs, output_all_encoded_layers=True, output_text_and_object_separately=True)
        b = ora(*inputs, output_all_encoded_layers=True, output_text_and_object_separately=True)
    for la, lb in zip(a[0], b[0]):
        assert torch.allclose(la, lb, atol=2e-5, rtol=1e-5)
    for la, lb in zip(a[1], b[1]):
        assert torch.allclose(la, lb, atol=2e-5, rtol=1e-5)
    assert torch.allclose(a[2], b[2], atol=2e-5, rtol=1e-5)


def test_reference_cpu_roialign_equals_c_oracle():
    import build_ref
    import roi_align as ro
    ref = build_ref.load()
    g = torch.Generator().manual_seed(3)
    f = torch.randn(2, 8, 38, 63, generator=g)
    K = 40
    x1 = torch.rand(K, generator=g) * 900
    y1 = torch.rand(K, generator=g) * 550
    rois = torch.stack([torch.randint(0, 2, (K,), generator=g).float(), x1, y1, x1 + torch.rand(K, generator=g) * 300,
                        y1 + torch.rand(K, generator=g) * 300], 1)
    for sr in (1, 2, 0):
        a = ref.roi_align_forward(f, rois, 1 / 16.0, 14, 14, sr).numpy()
        b = ro.roi_align_forward(f.numpy(), rois.numpy(), 1 / 16.0, 14, 14, sr)
        assert np.array_equal(a, b)


def test_dropin_classes_have_the_reference_checkpoint_abi():
    """state_dict keys + shapes of the drop-in modules == the live reference modules."""
    ref_shim.install()
    import vlbert_b200
    from common.visual_linguistic_bert import VisualLinguisticBert, VisualLinguisticBertForPretraining
    kw = dict(vocab_size=300, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=512,
              max_position_embeddings=80, visual_size=128, visual_region_classes=11, pos_embedding_frozen=False)
    for ref_cls, our_cls, args in ((VisualLinguisticBert, vlbert_b200.VisualLinguisticBert, ()),
                                   (VisualLinguisticBertForPretraining, vlbert_b200.VisualLinguisticBertForPretraining, (None, True, True, True))):
        a = ref_cls(ref_shim.vlbert_config(**kw), *args).state_dict()
        b = our_cls(vo.default_config(**kw), *args).state_dict()
        assert list(a.keys()) == list(b.keys())
        assert all(a[k].shape == b[k].shape for k in a)


def test_dropin_install_patches_the_reference_modules():
    ref_shim.install()
    import vlbert_b200
    assert vlbert_b200.dropin.install()
    import common.fast_rcnn
    import common.visual_linguistic_bert as m
    assert m.VisualLinguisticBert is vlbert_b200.VisualLinguisticBert
    assert m.VisualLinguisticBertForPretraining is vlbert_b200.VisualLinguisticBertForPretraining
    from easydict import EasyDict
    frcnn = common.fast_rcnn.FastRCNN(EasyDict({"NETWORK": {"IMAGE_FEAT_PRECOMPUTED": True, "IMAGE_SEMANTIC": False}}), True, 768)
    assert isinstance(frcnn, vlbert_b200.FastRCNN)
    assert list(frcnn.state_dict().keys()) == ["obj_downsample.1.weight", "obj_downsample.1.bias"]
    # the task module builds on the patched classes (construction only: forward needs a GPU)
    import importlib
    import pretrain.modules.resnet_vlbert_for_pretraining as tm
    importlib.reload(tm)
    assert tm.VisualLinguisticBertForPretraining is vlbert_b200.VisualLinguisticBertForPretraining


def test_end_to_end_fastrcnn_has_the_reference_checkpoint_abi(monkeypatch):
    """ResNet-101 C4 + res5 head (common/fast_rcnn.py:36-102): same state_dict keys/shapes and the same trainable set."""
    ref_shim.install()
    import vlbert_b200
    import torch.utils.model_zoo as model_zoo
    monkeypatch.setattr(model_zoo, "load_url", lambda *a, **k: {})       # no network: the zoo download returns no tensors
    import warnings
    import common.backbone.resnet.resnet as ref_resnet
    import common.fast_rcnn
    ref_cls = getattr(common.fast_rcnn.FastRCNN, "__wrapped__", None) or _unpatched_fastrcnn()
    from easydict import EasyDict
    cfg = EasyDict({"NETWORK": dict(IMAGE_FEAT_PRECOMPUTED=False, IMAGE_SEMANTIC=False, IMAGE_STRIDE_IN_1x1=True, IMAGE_C5_DILATED=True,
                                    IMAGE_NUM_LAYERS=101, IMAGE_PRETRAINED="", IMAGE_PRETRAINED_EPOCH=0, OUTPUT_CONV5=False,
                                    IMAGE_FROZEN_BN=True, IMAGE_FROZEN_BACKBONE_STAGES=[1, 2])})
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = ref_cls(cfg, True, 768, True)
    ours = vlbert_b200.FastRCNN(cfg, True, 768, True)
    a, b = ref.state_dict(), ours.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(a[k].shape == b[k].shape for k in a)
    ta = [n for n, p in ref.named_parameters() if p.requires_grad]
    tb = [n for n, p in ours.named_parameters() if p.requires_grad]
    assert ta == tb and len(ta) > 0
    ours.load_state_dict(a, strict=True)


def _unpatched_fastrcnn():
    import importlib
    import common.fast_rcnn as m
    return importlib.reload(m).FastRCNN


def _fake_hf_checkpoint(kind, H=64, L=2, I=128, vocab=120, max_pos=40):
    """a state_dict with HuggingFace BERT (old TF-style gamma/beta names, cls.* heads) or RoBERTa (roberta.*, lm_head.*, one
    token type) key names, random values"""
    g = torch.Generator().manual_seed(7 if kind == "bert" else 8)
    r = lambda *s: torch.randn(*s, generator=g)
    pre = "bert." if kind == "bert" else "roberta."
    ln_w, ln_b = ("gamma", "beta") if kind == "bert" else ("weight", "bias")
    sd = {pre + "embeddings.word_embeddings.weight": r(vocab, H), pre + "embeddings.position_embeddings.weight": r(max_pos, H),
          pre + "embeddings.token_type_embeddings.weight": r(2 if kind == "bert" else 1, H),
          pre + "embeddings.LayerNorm." + ln_w: r(H), pre + "embeddings.LayerNorm." + ln_b: r(H),
          pre + "embeddings.position_ids": torch.arange(max_pos)[None],                       # unexpected key
          pre + "pooler.dense.weight": r(H, H), pre + "pooler.dense.bias": r(H)}
    for l in range(L):
        q = pre + "encoder.layer.%d." % l
        for n, shp in (("attention.self.query", (H, H)), ("attention.self.key", (H, H)), ("attention.self.value", (H, H)),
                       ("attention.output.dense", (H, H)), ("intermediate.dense", (I, H)), ("output.dense", (H, I))):
            sd[q + n + ".weight"], sd[q + n + ".bias"] = r(*shp), r(shp[0])
        for n in ("attention.output.LayerNorm", "output.LayerNorm"):
            sd[q + n + "." + ln_w], sd[q + n + "." + ln_b] = r(H), r(H)
    if kind == "bert":
        sd.update({"cls.seq_relationship.weight": r(2, H), "cls.seq_relationship.bias": r(2),
                   "cls.predictions.bias": r(vocab), "cls.predictions.transform.dense.weight": r(H, H),
                   "cls.predictions.transform.dense.bias": r(H), "cls.predictions.transform.LayerNorm.gamma": r(H),
                   "cls.predictions.transform.LayerNorm.beta": r(H), "cls.predictions.decoder.weight": r(vocab, H)})
    else:
        sd.update({"lm_head.bias": r(vocab), "lm_head.dense.weight": r(H, H), "lm_head.dense.bias": r(H),
                   "lm_head.layer_norm.weight": r(H), "lm_head.layer_norm.bias": r(H), "lm_head.decoder.weight": r(vocab, H)})
    sd["some.other.tensor"] = r(3)
    return sd


@pytest.mark.parametrize("kind", ["bert", "roberta"])
@pytest.mark.parametrize("pretraining", [False, True])
def test_language_checkpoint_loader_equals_the_reference(tmp_path, kind, pretraining, capsys):
    """load_language_pretrained_model (common/visual_linguistic_bert.py:243-309 and :382-470): same tensors end up in the same
    parameters, including the relationship / MLM heads of the pre-training class and the RoBERTa renames."""
    ref_shim.install()
    import vlbert_b200
    import common.visual_linguistic_bert as ref_mod
    import importlib
    ref_mod = importlib.reload(ref_mod)            # undo a drop-in patch made by an earlier test
    path = str(tmp_path / "lm.bin")
    torch.save(_fake_hf_checkpoint(kind), path)
    kw = dict(vocab_size=120, hidden_size=64, num_hidden_layers=2, num_attention_heads=1, intermediate_size=128,
              max_position_embeddings=40, visual_size=64, visual_region_classes=7, pos_embedding_frozen=False)
    if pretraining:
        ref = ref_mod.VisualLinguisticBertForPretraining(ref_shim.vlbert_config(**kw), path)
        ours = vlbert_b200.VisualLinguisticBertForPretraining(vo.default_config(**kw), path)
    else:
        ref = ref_mod.VisualLinguisticBert(ref_shim.vlbert_config(**kw), path)
        ours = vlbert_b200.VisualLinguisticBert(vo.default_config(**kw), path)
    capsys.readouterr()
    a, b = ref.state_dict(), ours.state_dict()
    assert list(a.keys()) == list(b.keys())
    heads = ("relationsip_head.", "mlm_head.") if kind == "bert" else ("mlm_head.",)     # RoBERTa ships no relationship head
    loaded = [k for k in a if k.startswith(("encoder.", "embedding_LayerNorm.", "word_embeddings.", "position_embeddings.", "pooler.")
                                           + heads)] + ["token_type_embeddings.weight"]
    for k in loaded:
        if k == "token_type_embeddings.weight":
            rows = 2 if (kind == "bert" or pretraining) else 3    # rows the loader writes (the rest keep their random init)
            assert torch.equal(a[k][:rows], b[k][:rows]), k
        elif k.startswith("mlm_head.") and "decoder.weight" not in k or not k.startswith("mlm_head."):
            assert torch.equal(a[k], b[k]), k
    assert torch.equal(b["mlm_head.predictions.decoder.weight"], b["word_embeddings.weight"]) if pretraining else True


def test_frontend_oracle_with_instance_masks_equals_the_reference(monkeypatch):
    """FastRCNN end to end with `segms` (the VCR path): oracle/frontend_oracle.py against the live reference module."""
    ref_shim.install()
    import warnings
    import torch.utils.model_zoo as model_zoo
    import frontend_oracle as fo
    from synth import frontend_config, synth_frontend_inputs
    monkeypatch.setattr(model_zoo, "load_url", lambda *a, **k: {})
    ref_cls = _unpatched_fastrcnn()
    import common.lib.roi_pooling.roi_align as ref_roi      # an earlier test may have installed the (CUDA-only) drop-in extension
    f, b = ref_shim.cpu_roi_functions()
    monkeypatch.setattr(ref_roi.C_ROIPooling, "roi_align_forward", f, raising=False)
    monkeypatch.setattr(ref_roi.C_ROIPooling, "roi_align_backward", b, raising=False)
    from easydict import EasyDict
    cfg = EasyDict({"NETWORK": dict(vars(frontend_config(50).NETWORK))})
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = ref_cls(cfg, True, 32, False)
    sd = fo.synth_frontend_state({k: v.shape for k, v in ref.state_dict().items()}, 21)
    ref.load_state_dict(sd, strict=True)
    ref.eval()
    images, boxes, box_mask, im_info, _ = synth_frontend_inputs(12, B=2, R=3, H=80, W=112)
    box_mask[0, 2] = False
    segms = (torch.rand(2, 3, 14, 14, generator=torch.Generator().manual_seed(2)) > 0.5).float()
    torch.set_num_threads(8)
    with torch.no_grad():
        want = ref(images=images, boxes=boxes, box_mask=box_mask, im_info=im_info, segms=segms)
        got_obj, got_raw = fo.fast_rcnn_end2end(sd, images, boxes, box_mask, im_info, layers=fo.LAYERS[50], segms=segms)
    assert (got_obj - want["obj_reps"]).abs().max() <= 1e-4 * want["obj_reps"].abs().max()
    assert (got_raw - want["obj_reps_raw"]).abs().max() <= 1e-4 * want["obj_reps_raw"].abs().max()


def test_reference_pretraining_task_module_runs_unchanged_on_the_dropin(tmp_path, monkeypatch, capsys):
    """The boundary itself: pretrain/modules/resnet_vlbert_for_pretraining.py built from the reference's own
    cfgs/pretrain/base_prec_4x16G_fp32.yaml (2 layers), once untouched and once after vlbert_b200.dropin.install() -- same
    state_dict keys, and on the same weights and inputs the same logits, losses and parameter gradients.  The kernels are
    replaced by fp32 torch stand-ins (tests/cpu_shim.py), so this checks everything the Python layer of the drop-in decides."""
    import importlib
    import warnings
    ref_shim.install()
    import cpu_shim
    import vlbert_b200
    vocab_dir = tmp_path / "bert-base-uncased"
    vocab_dir.mkdir()
    (vocab_dir / "vocab.txt").write_text("\n".join(["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"] + ["tok%d" % i for i in range(30517)]) + "\n")
    import pretrain.function.config as cfgmod
    cfgmod = importlib.reload(cfgmod)               # update_config() is not idempotent on the module-level singleton
    config, update_config = cfgmod.config, cfgmod.update_config
    update_config(os.path.join(ref_shim.REFERENCE_ROOT, "cfgs", "pretrain", "base_prec_4x16G_fp32.yaml"))
    config.NETWORK.VLBERT.num_hidden_layers = 2
    config.NETWORK.VLBERT.hidden_dropout_prob = 0.0
    config.NETWORK.VLBERT.attention_probs_dropout_prob = 0.0
    config.NETWORK.BERT_MODEL_NAME = str(vocab_dir)
    config.NETWORK.BERT_PRETRAINED = ""
    import common.fast_rcnn
    import common.visual_linguistic_bert
    import pretrain.modules.resnet_vlbert_for_pretraining as tm

    def fresh():
        importlib.reload(common.fast_rcnn)
        importlib.reload(common.visual_linguistic_bert)
        return importlib.reload(tm)

    g = torch.Generator().manual_seed(3)
    B, R, T, C = 2, 4, 8, config.NETWORK.VLBERT.visual_region_classes
    x1, y1 = torch.rand(B, R, generator=g) * 300, torch.rand(B, R, generator=g) * 200
    boxes = torch.cat((torch.stack((x1, y1, x1 + 20 + torch.rand(B, R, generator=g) * 250, y1 + 20 + torch.rand(B, R, generator=g) * 150), -1),
                       torch.randn(B, R, 2048, generator=g)), -1)
    boxes[1, 3] = -2.0                                                       # pretrain/data/collate_batch.py:39
    im_info = torch.tensor([[600., 400., 1., 1., 0.], [600., 400., 1., 1., 1.]])
    text = torch.randint(1000, 30522, (B, T), generator=g)
    text[1, 6:] = 0
    rel_label = torch.ones(B, dtype=torch.long)
    mlm_labels = torch.full((B, T), -1, dtype=torch.long)
    mlm_labels[0, 2], mlm_labels[1, 4] = 1234, 4321
    mvrc_ops = torch.tensor([[0, 1, 0, 0], [1, 0, 0, 0]])
    mvrc_labels = torch.zeros(B, R, C)
    mvrc_labels[mvrc_ops == 1] = torch.softmax(torch.randn(2, C, generator=g), -1)

    def run(model):
        model.zero_grad()
        out, loss = model(None, boxes.clone(), im_info, text, rel_label, mlm_labels, mvrc_ops, mvrc_labels.clone())
        loss.backward()
        return out, loss, {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}

    try:
        torch.manual_seed(0)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = fresh().ResNetVLBERTForPretraining(config)
            ref.eval()          # (the reference's train() override returns None)
            sd = {k: v.clone() for k, v in ref.state_dict().items()}
            o1, l1, g1 = run(ref)                 # the untouched reference first: install() rebinds the names its super() calls use
            assert vlbert_b200.dropin.install()
            cpu_shim.install(monkeypatch)
            cpu_shim.install_encoder(monkeypatch)
            ours = importlib.reload(tm).ResNetVLBERTForPretraining(config)
            ours.eval()
            assert isinstance(ours.vlbert, vlbert_b200.VisualLinguisticBertForPretraining)
            assert isinstance(ours.image_feature_extractor, vlbert_b200.FastRCNN)
            assert list(ours.state_dict().keys()) == list(sd.keys())
            ours.load_state_dict(sd, strict=True)
            o2, l2, g2 = run(ours)
        capsys.readouterr()
        assert abs(float(l1.detach()) - float(l2.detach())) <= 1e-5 * abs(float(l1.detach())), (float(l1.detach()), float(l2.detach()))
        for k in ("relationship_logits", "mlm_logits", "mvrc_logits", "relationship_loss", "mlm_loss", "mvrc_loss"):
            assert (o1[k] is None) == (o2[k] is None), k
            if o1[k] is not None:
                assert (o1[k] - o2[k]).abs().max() <= 2e-4 * o1[k].abs().max().clamp_min(1e-6), k
        assert set(g1.keys()) == set(g2.keys())
        for k in g1:
            assert (g1[k] - g2[k]).abs().max() <= 5e-4 * g1[k].abs().max().clamp_min(1e-9), k
    finally:
        fresh()


def _task_module_parity(monkeypatch, capsys, task, yaml_name, cls_name, tweak, build_inputs, call, vocab_dir, tol_out=2e-4, tol_grad=5e-4, zoo=None, after_build=None):
    """Build `<task>.modules.<cls_name>` from the reference's own yaml twice -- untouched, then with vlbert_b200.dropin installed
    and the kernels replaced by the fp32 stand-ins of tests/cpu_shim.py -- load the same weights, run the same inputs, and compare
    every tensor output, the loss and every parameter gradient."""
    import importlib
    import warnings
    import cpu_shim
    import vlbert_b200
    import torch.utils.model_zoo as model_zoo
    monkeypatch.setattr(model_zoo, "load_url", lambda *a, **k: (zoo if zoo is not None else {}))   # no network: the "model zoo"
    cfgmod = importlib.reload(importlib.import_module(task + ".function.config"))   # update_config() is not idempotent on the singleton
    config = cfgmod.config
    cfgmod.update_config(os.path.join(ref_shim.REFERENCE_ROOT, "cfgs", task, yaml_name))
    config.NETWORK.VLBERT.hidden_dropout_prob = 0.0
    config.NETWORK.VLBERT.attention_probs_dropout_prob = 0.0
    config.NETWORK.BERT_MODEL_NAME = str(vocab_dir)
    config.NETWORK.BERT_PRETRAINED = ""
    tweak(config)
    import common.fast_rcnn
    import common.visual_linguistic_bert
    import common.lib.roi_pooling.roi_align as ref_roi
    f, b = ref_shim.cpu_roi_functions()
    monkeypatch.setattr(ref_roi.C_ROIPooling, "roi_align_forward", f, raising=False)
    monkeypatch.setattr(ref_roi.C_ROIPooling, "roi_align_backward", b, raising=False)
    tm = importlib.import_module(task + ".modules")

    def fresh():
        importlib.reload(common.fast_rcnn)
        importlib.reload(common.visual_linguistic_bert)
        for name in sorted(k for k in sys.modules if k.startswith(task + ".modules.")):
            importlib.reload(sys.modules[name])
        return importlib.reload(tm)

    inputs = build_inputs(config)

    def run(model):
        model.zero_grad()
        out, loss = call(model, [t.clone() if torch.is_tensor(t) else t for t in inputs])
        loss.backward()
        return out, loss.detach(), {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}

    try:
        torch.manual_seed(0)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = getattr(fresh(), cls_name)(config)
            ref.eval()
            sd = {k: v.clone() for k, v in ref.state_dict().items()}
            o1, l1, g1 = run(ref)                 # the untouched reference first: install() rebinds the names its super() calls use
            assert vlbert_b200.dropin.install()
            cpu_shim.install(monkeypatch)
            cpu_shim.install_encoder(monkeypatch)
            for name in sorted(k for k in sys.modules if k.startswith(task + ".modules.")):
                importlib.reload(sys.modules[name])
            ours = getattr(importlib.reload(tm), cls_name)(config)
            ours.eval()
            assert isinstance(ours.image_feature_extractor, vlbert_b200.FastRCNN)
            assert list(ours.state_dict().keys()) == list(sd.keys())
            if after_build is not None:
                after_build(sd, ours)
            ours.load_state_dict(sd, strict=True)
            o2, l2, g2 = run(ours)
        capsys.readouterr()
        assert abs(float(l1) - float(l2)) <= 1e-5 * max(1e-6, abs(float(l1))), (float(l1), float(l2))
        for k in o1:
            if torch.is_tensor(o1[k]) and o1[k].is_floating_point():
                assert (o1[k] - o2[k]).abs().max() <= tol_out * o1[k].abs().max().clamp_min(1e-6), k
        assert set(g1.keys()) == set(g2.keys())
        for k in g1:
            assert (g1[k] - g2[k]).abs().max() <= tol_grad * g1[k].abs().max().clamp_min(1e-9), k
        return ours
    finally:
        fresh()


def _vocab_dir(tmp_path, n, with_checkpoint=None):
    d = tmp_path / "bert-base-uncased"
    d.mkdir()
    (d / "vocab.txt").write_text("\n".join(["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"] + ["tok%d" % i for i in range(n - 5)]) + "\n")
    if with_checkpoint is not None:
        torch.save(with_checkpoint, str(d / "pytorch_model.bin"))
    return d


def test_reference_vqa_task_module_runs_unchanged_on_the_dropin(tmp_path, monkeypatch, capsys):
    """vqa/modules/resnet_vlbert_for_vqa.py from cfgs/vqa/base_4x16G_fp32.yaml (BASELINE config 3's module; 2 layers, vocabulary cut to
    1000): `mlm` classifier initialised from a BERT checkpoint through the module's own loader, uint8 text masks from
    prepare_text_from_qa, precomputed region features."""
    ref_shim.install()
    vocab = 1000
    d = _vocab_dir(tmp_path, vocab, _fake_hf_checkpoint("bert", H=768, L=2, I=3072, vocab=vocab, max_pos=512))

    def tweak(config):
        config.NETWORK.VLBERT.num_hidden_layers = 2
        config.NETWORK.VLBERT.vocab_size = vocab
        config.NETWORK.CLASSIFIER_DROPOUT = 0.0
        config.DATASET.ANSWER_VOCAB_SIZE = 37

    def build_inputs(config):
        g = torch.Generator().manual_seed(5)
        B, R, T = 3, 5, 7
        x1, y1 = torch.rand(B, R, generator=g) * 300, torch.rand(B, R, generator=g) * 200
        boxes = torch.cat((torch.stack((x1, y1, x1 + 20 + torch.rand(B, R, generator=g) * 250, y1 + 20 + torch.rand(B, R, generator=g) * 150), -1),
                           torch.randn(B, R, 2048, generator=g)), -1)
        boxes[1, 3:] = -2.0
        boxes[2, 4:] = -2.0
        im_info = torch.tensor([[600., 400., 1., 1.]] * B)
        question = torch.randint(5, vocab, (B, T), generator=g)
        question[0, 5:] = 0
        question[2, 3:] = 0
        label = torch.rand(B, 37, generator=g)
        return [None, boxes, im_info, question, label]

    _task_module_parity(monkeypatch, capsys, "vqa", "base_4x16G_fp32.yaml", "ResNetVLBERT", tweak, build_inputs,
                        lambda m, ins: m.train_forward(*ins), d)


def test_reference_vcr_task_module_runs_unchanged_on_the_dropin(tmp_path, monkeypatch, capsys):
    """vcr/modules/resnet_vlbert_for_vcr.py from cfgs/vcr/base_q2a_4x16G_fp32.yaml (BASELINE config 4's module at base width; 2 layers):
    images through the END-TO-END FastRCNN (ResNet-101 C4 + RoIAlign + dilated res5) with instance masks (`segms`) and object classes,
    TimeDistributed VL-BERT over the four answer choices, text/object outputs returned separately, the CNN regularisation head on
    the object outputs.  Exercises the drop-in's end-to-end front end under its real caller."""
    ref_shim.install()
    import torch.utils.model_zoo as model_zoo
    from common.backbone.resnet.resnet import Bottleneck, ResNet
    vocab = 600
    d = _vocab_dir(tmp_path, vocab)
    torch.manual_seed(1)
    zoo = ResNet(Bottleneck, [3, 4, 23, 3], num_classes=None, expose_stages=[5]).state_dict()      # what the model zoo would return
    for k, v in zoo.items():                                                                      # non-trivial frozen statistics
        if k.endswith("running_var"):
            v.uniform_(0.5, 1.5)
        elif k.endswith("running_mean"):
            v.normal_(0, 0.1)
        elif ".bn" in k and k.endswith("weight") or "downsample.1.weight" in k:
            v.uniform_(0.3, 0.6)

    torch.save(zoo, str(tmp_path / "resnet101-0000.model"))                                       # NETWORK.IMAGE_PRETRAINED checkpoint

    def tweak(config):
        config.NETWORK.VLBERT.num_hidden_layers = 2
        config.NETWORK.VLBERT.vocab_size = vocab
        config.NETWORK.CLASSIFIER_DROPOUT = 0.0
        config.NETWORK.IMAGE_PRETRAINED = str(tmp_path / "resnet101")
        config.NETWORK.IMAGE_PRETRAINED_EPOCH = 0
        assert config.NETWORK.IMAGE_FEAT_PRECOMPUTED is False and config.NETWORK.ENABLE_CNN_REG_LOSS and config.NETWORK.CNN_LOSS_TOP

    def after_build(ref_sd, ours):
        # both sides initialise the backbone and the res5 head (layer4.*) from the same IMAGE_PRETRAINED file (common/fast_rcnn.py:41-63,111-120)
        mine = ours.state_dict()
        for k in ref_sd:
            if k.startswith(("image_feature_extractor.backbone.", "image_feature_extractor.roi_head_feature_extractor.", "image_feature_extractor.head.")):
                assert torch.equal(mine[k], ref_sd[k]), k

    def build_inputs(config):
        g = torch.Generator().manual_seed(9)
        B, R, NC, Lq, La, H, W = 2, 3, 4, 5, 4, 64, 96
        image = torch.randn(B, 3, H, W, generator=g)
        x1, y1 = torch.rand(B, R, generator=g) * 40, torch.rand(B, R, generator=g) * 25
        boxes = torch.stack((x1, y1, x1 + 16 + torch.rand(B, R, generator=g) * 38, y1 + 16 + torch.rand(B, R, generator=g) * 20,
                             torch.randint(1, 81, (B, R), generator=g).float()), -1)
        boxes[:, 0, :4] = torch.tensor([0.0, 0.0, W - 1.0, H - 1.0])          # the whole image as the first box (ADD_IMAGE_AS_A_BOX)
        boxes[:, 0, 4] = 0
        boxes[1, 2] = -1.0                                                     # padded box
        masks = (torch.rand(B, R, 14, 14, generator=g) > 0.35).float()
        masks[:, 0] = 1.0
        question = torch.stack((torch.randint(5, vocab, (B, Lq), generator=g), torch.randint(-1, 2, (B, Lq), generator=g)), -1)
        question[1, 3:] = 0
        answers = torch.stack((torch.randint(5, vocab, (B, NC, La), generator=g), torch.randint(-1, 2, (B, NC, La), generator=g)), -1)
        answers[0, 1, 2:] = 0
        answers[1, 3, 3:] = 0
        answer_label = torch.tensor([2, 0])
        im_info = torch.tensor([[float(W), float(H), 1.0, 1.0]] * B)
        return [image, boxes, masks, question, None, answers, None, answer_label, im_info]

    torch.set_num_threads(8)
    _task_module_parity(monkeypatch, capsys, "vcr", "base_q2a_4x16G_fp32.yaml", "ResNetVLBERT", tweak, build_inputs,
                        lambda m, ins: m.train_forward(*ins), d, zoo=zoo, after_build=after_build)


def test_reference_refcoco_task_module_runs_unchanged_on_the_dropin(tmp_path, monkeypatch, capsys):
    """refcoco/modules/resnet_vlbert_for_refcoco.py from cfgs/refcoco/base_detected_regions_4x16G.yaml (2 layers): end-to-end FastRCNN
    without masks, `with_pooler: false`, per-region classifier on the object outputs."""
    ref_shim.install()
    from common.backbone.resnet.resnet import Bottleneck, ResNet
    vocab = 500
    d = _vocab_dir(tmp_path, vocab)
    torch.manual_seed(2)
    zoo = ResNet(Bottleneck, [3, 4, 23, 3], num_classes=None, expose_stages=[5]).state_dict()
    for k, v in zoo.items():
        if k.endswith("running_var"):
            v.uniform_(0.5, 1.5)
        elif ".bn" in k and k.endswith("weight") or "downsample.1.weight" in k:
            v.uniform_(0.3, 0.6)
    torch.save(zoo, str(tmp_path / "resnet101-0000.model"))

    def tweak(config):
        config.NETWORK.VLBERT.num_hidden_layers = 2
        config.NETWORK.VLBERT.vocab_size = vocab
        config.NETWORK.IMAGE_PRETRAINED = str(tmp_path / "resnet101")
        config.NETWORK.IMAGE_PRETRAINED_EPOCH = 0

    def build_inputs(config):
        g = torch.Generator().manual_seed(13)
        B, R, T, H, W = 2, 4, 6, 64, 80
        image = torch.randn(B, 3, H, W, generator=g)
        x1, y1 = torch.rand(B, R, generator=g) * 30, torch.rand(B, R, generator=g) * 25
        boxes = torch.stack((x1, y1, x1 + 16 + torch.rand(B, R, generator=g) * 30, y1 + 16 + torch.rand(B, R, generator=g) * 20), -1)
        boxes[0, 3] = -2.0
        im_info = torch.tensor([[float(W), float(H), 1.0, 1.0]] * B)
        expression = torch.randint(5, vocab, (B, T), generator=g)
        expression[1, 4:] = 0
        label = (torch.rand(B, R, generator=g) > 0.5).float()
        return [image, boxes, im_info, expression, label]

    torch.set_num_threads(8)
    _task_module_parity(monkeypatch, capsys, "refcoco", "base_detected_regions_4x16G.yaml", "ResNetVLBERT", tweak, build_inputs,
                        lambda m, ins: m.train_forward(*ins), d, zoo=zoo)


@pytest.mark.parametrize("yaml_name", ["base_e2e_16x16G_fp16.yaml", "base_prec_4x16G_fp32.yaml"])
def test_reference_multitask_pretraining_module_runs_unchanged_on_the_dropin(tmp_path, monkeypatch, capsys, yaml_name):
    """pretrain/modules/resnet_vlbert_for_pretraining_multitask.py -- the MODULE both pre-training yamls name, i.e. the real caller of
    BASELINE config 2 (precomputed features) and config 5 (end to end).  It appends text-only samples (no boxes at all: box_mask
    all False) to the batch, masks region features with `mask_visual_embed` on the end-to-end path, and uses all three heads."""
    ref_shim.install()
    from common.backbone.resnet.resnet import Bottleneck, ResNet
    e2e = "e2e" in yaml_name
    vocab = 700
    d = _vocab_dir(tmp_path, vocab)
    zoo = None
    if e2e:
        torch.manual_seed(4)
        zoo = ResNet(Bottleneck, [3, 4, 23, 3], num_classes=None, expose_stages=[5]).state_dict()
        for k, v in zoo.items():
            if k.endswith("running_var"):
                v.uniform_(0.5, 1.5)
            elif ".bn" in k and k.endswith("weight") or "downsample.1.weight" in k:
                v.uniform_(0.3, 0.6)
        torch.save(zoo, str(tmp_path / "resnet101-0000.model"))

    def tweak(config):
        config.NETWORK.VLBERT.num_hidden_layers = 2
        config.NETWORK.VLBERT.vocab_size = vocab
        config.NETWORK.VLBERT.visual_region_classes = 23
        if e2e:
            config.NETWORK.IMAGE_PRETRAINED = str(tmp_path / "resnet101")
            config.NETWORK.IMAGE_PRETRAINED_EPOCH = 0

    def build_inputs(config):
        g = torch.Generator().manual_seed(17)
        B, R, T, C, H, W = 2, 4, 7, 23, 64, 96
        x1, y1 = torch.rand(B, R, generator=g) * 40, torch.rand(B, R, generator=g) * 25
        box4 = torch.stack((x1, y1, x1 + 16 + torch.rand(B, R, generator=g) * 38, y1 + 16 + torch.rand(B, R, generator=g) * 20), -1)
        if e2e:
            image, boxes = torch.randn(B, 3, H, W, generator=g), box4
        else:
            image, boxes = None, torch.cat((box4, torch.randn(B, R, 2048, generator=g)), -1)
        boxes[1, 3] = -2.0
        im_info = torch.tensor([[float(W), float(H), 1.0, 1.0, 0.0], [float(W), float(H), 1.0, 1.0, 1.0]])
        text = torch.randint(5, vocab, (B, T), generator=g)
        text[1, 5:] = 0
        rel = torch.ones(B, dtype=torch.long)
        mlm = torch.full((B, T), -1, dtype=torch.long)
        mlm[0, 2], mlm[1, 3] = 77, 88
        mvrc_ops = torch.tensor([[0, 1, 0, 0], [1, 0, 0, 0]])
        mvrc_labels = torch.zeros(B, R, C)
        mvrc_labels[mvrc_ops == 1] = torch.softmax(torch.randn(2, C, generator=g), -1)
        aux_text = torch.randint(5, vocab, (3, 9), generator=g)            # three text-only samples, longer than the captions
        aux_text[0, 6:] = 0
        aux_mlm = torch.full((3, 9), -1, dtype=torch.long)
        aux_mlm[0, 1], aux_mlm[2, 8] = 99, 111
        return [image, boxes, im_info, text, rel, mlm, mvrc_ops, mvrc_labels, aux_text, aux_mlm]

    torch.set_num_threads(8)
    _task_module_parity(monkeypatch, capsys, "pretrain", yaml_name, "ResNetVLBERTForPretrainingMultitask", tweak, build_inputs,
                        lambda m, ins: m(*ins), d, zoo=zoo)
