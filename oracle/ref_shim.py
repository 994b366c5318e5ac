"""TEST INFRASTRUCTURE ONLY -- import shim for the unmodified reference at /root/reference.

Only usable in the build container (the GPU box has no /root/reference).  It is used by
`oracle/make_golden.py` to produce the committed fixtures in tests/golden/ and by the CPU tests that
pin `oracle/vlbert_oracle.py` (the portable restatement) against the real reference modules.

Stubs follow SURVEY.md section 8(c): boto3/botocore (external/pytorch_pretrained_bert/file_utils.py:18),
jsonlines, tensorboardX, easydict, the native C_ROIPooling extension (common/lib/roi_pooling/__init__.py)
and a PyYAML-6 `yaml.load` default Loader (pretrain/function/config.py:181).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("VLBERT_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "common"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def cpu_roi_functions():
    """CPU stand-ins for the reference's native C_ROIPooling.roi_align_forward/backward (torchvision's op of the same
    maskrcnn-benchmark lineage; bit-identical to the reference CPU kernel for aligned=False)."""
    import torch
    import torchvision.ops as tvo

    def roi_align_forward(input, rois, spatial_scale, pooled_h, pooled_w, sampling_ratio):
        return tvo.roi_align(input, rois, (pooled_h, pooled_w), spatial_scale, sampling_ratio, aligned=False)

    def roi_align_backward(grad, rois, spatial_scale, ph, pw, bs, ch, h, w, sampling_ratio):
        return torch.ops.torchvision._roi_align_backward(grad, rois, spatial_scale, ph, pw, bs, ch, h, w,
                                                         sampling_ratio, False)
    return roi_align_forward, roi_align_backward


def install():
    """Make `import common.visual_linguistic_bert` etc. work.  Idempotent."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    _stub("boto3")
    bc = _stub("botocore")
    bce = _stub("botocore.exceptions", ClientError=type("ClientError", (Exception,), {}))
    bc.exceptions = bce
    _stub("jsonlines")

    class _SummaryWriter(object):
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, name):
            return lambda *a, **k: None

    _stub("tensorboardX", SummaryWriter=_SummaryWriter)

    class EasyDict(dict):
        def __init__(self, d=None, **kw):
            super().__init__()
            d = dict(d or {})
            d.update(kw)
            for k, v in d.items():
                self[k] = v

        def __setitem__(self, k, v):
            if isinstance(v, dict) and not isinstance(v, EasyDict):
                v = EasyDict(v)
            elif isinstance(v, (list, tuple)):
                v = type(v)(EasyDict(x) if isinstance(x, dict) and not isinstance(x, EasyDict) else x for x in v)
            super().__setitem__(k, v)

        __setattr__ = __setitem__

        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

    _stub("easydict", EasyDict=EasyDict)

    # the native extension: forward/backward via torchvision's maskrcnn-benchmark-lineage op
    # (bit-identical to the reference CPU kernel for aligned=False).
    roi_align_forward, roi_align_backward = cpu_roi_functions()

    import importlib
    import importlib.machinery
    pkg_name = "common.lib.roi_pooling"
    _stub(pkg_name + ".C_ROIPooling", roi_align_forward=roi_align_forward, roi_align_backward=roi_align_backward,
          roi_pool_forward=None, roi_pool_backward=None)

    import yaml
    if not getattr(yaml, "_vlb_patched", False):
        _orig = yaml.load

        def _load(stream, Loader=None):
            return _orig(stream, Loader=Loader or yaml.FullLoader)

        yaml.load = _load
        yaml._vlb_patched = True
    return REFERENCE_ROOT


def vlbert_config(**over):
    """NETWORK.VLBERT config with the reference defaults (pretrain/function/config.py:87-115) for BERT-base."""
    from easydict import EasyDict
    cfg = dict(
        vocab_size=30522, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
        intermediate_size=3072, hidden_act="gelu", hidden_dropout_prob=0.0,
        attention_probs_dropout_prob=0.0, max_position_embeddings=512, type_vocab_size=3,
        initializer_range=0.02, visual_size=768, visual_scale_text_init=1.0, visual_scale_object_init=1.0,
        visual_ln=True, word_embedding_frozen=False, with_pooler=True, position_padding_idx=-1,
        obj_pos_id_relative=True,
    )
    cfg.update(over)
    return EasyDict(cfg)
