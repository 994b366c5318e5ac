"""Install the library under the reference's own module names, so that the reference's task modules
(`pretrain/modules/resnet_vlbert_for_pretraining.py`, `vqa/...`, `vcr/...`), `train_end2end.py` and
`common/trainer.py` run unchanged on top of it.  Call BEFORE the task modules are imported:

    import vlbert_b200; vlbert_b200.dropin.install()

What gets replaced (reference file:line):
  common.visual_linguistic_bert.VisualLinguisticBert / ...ForPretraining / ...MVRCHeadTransform  (:31, :312, :473)
  common.fast_rcnn.FastRCNN (precomputed features, or a Bottleneck ResNet-C4 end to end)          (common/fast_rcnn.py:17)
  common.lib.roi_pooling.C_ROIPooling, common.lib.roi_pooling.roi_align.ROIAlign                  (vision.cpp:6-11)
"""
import sys
import types


def install(reference_root=None):
    from . import modules as M
    if reference_root is not None and reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    # the native extension module (never built for modern torch; the reference imports it at module import time)
    ext = types.ModuleType("common.lib.roi_pooling.C_ROIPooling")
    ext.roi_align_forward = M.C_ROIPooling.roi_align_forward
    ext.roi_align_backward = M.C_ROIPooling.roi_align_backward
    ext.roi_pool_forward = M.C_ROIPooling.roi_pool_forward
    ext.roi_pool_backward = M.C_ROIPooling.roi_pool_backward
    sys.modules["common.lib.roi_pooling.C_ROIPooling"] = ext
    try:
        import common.visual_linguistic_bert as ref_vlb
        import common.fast_rcnn as ref_frcnn
    except ImportError as e:  # pragma: no cover
        raise RuntimeError("vlbert_b200.dropin.install: the reference tree is not importable (%s)" % e)
    ref_vlb.VisualLinguisticBert = M.VisualLinguisticBert
    ref_vlb.VisualLinguisticBertForPretraining = M.VisualLinguisticBertForPretraining
    ref_vlb.VisualLinguisticBertMVRCHeadTransform = M.VisualLinguisticBertMVRCHeadTransform
    ref_original = ref_frcnn.FastRCNN

    def fast_rcnn_factory(config, *a, **k):
        if config.NETWORK.IMAGE_FEAT_PRECOMPUTED or config.NETWORK.IMAGE_NUM_LAYERS in (50, 101, 152):
            return M.FastRCNN(config, *a, **k)
        return ref_original(config, *a, **k)  # BasicBlock backbones: reference convs + library RoIAlign (patched above)

    ref_frcnn.FastRCNN = fast_rcnn_factory
    return True
