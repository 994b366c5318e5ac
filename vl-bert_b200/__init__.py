"""vlbert_b200 -- B200-native (sm_100a) implementation of the VL-BERT hot path.

The product is `libvlbert_b200.so` (hand-written CUDA behind the C ABI in include/vlbert_b200.h); this package
is the Python host side that mirrors the reference's module interface:

    from vlbert_b200 import VisualLinguisticBert, FastRCNN, ROIAlign, C_ROIPooling
    vlbert_b200.dropin.install()     # monkey-patches the reference's common.* modules

Import with `import vlbert_b200` (root-level alias of the directory `vl-bert_b200/`).
"""
from . import _lib  # noqa: F401
from . import functional  # noqa: F401
from .modules import (C_ROIPooling, FastRCNN, ROIAlign, VisualLinguisticBert,  # noqa: F401
                      VisualLinguisticBertForPretraining, VisualLinguisticBertMVRCHeadTransform, default_config)
from . import dropin  # noqa: F401
from . import ddp  # noqa: F401
from . import optim  # noqa: F401
from . import glue  # noqa: F401
from . import region_shards  # noqa: F401
from .graphs import GraphedStep  # noqa: F401

__all__ = ["VisualLinguisticBert", "VisualLinguisticBertForPretraining", "VisualLinguisticBertMVRCHeadTransform",
           "FastRCNN", "ROIAlign", "C_ROIPooling", "functional", "dropin"]
