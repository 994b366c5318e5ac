"""Per-stage divergence of the library's ResNet-C4 forward from the CPU oracle (fp32 and bf16-storage graphs)."""
import os, sys
import torch
import torch.nn.functional as F
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (R, os.path.join(R, "tests"), os.path.join(R, "oracle")):
    sys.path.insert(0, p)
import frontend_oracle as fo
from synth import frontend_config, synth_frontend_inputs
import vlbert_b200
from vlbert_b200 import functional as VF
from vlbert_b200.resnet import _cba

torch.set_num_threads(16)
m = vlbert_b200.FastRCNN(frontend_config(101), True, 64, False)
sd = fo.synth_frontend_state({k: v.shape for k, v in m.state_dict().items()}, 77)
m.load_state_dict(sd); m = m.cuda().eval()
images = synth_frontend_inputs(78)[0]
rel = lambda a, b: ((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm()).item()
bb = m.backbone
for storage in (None, "bf16"):
    q = fo._q(storage)
    x = q(torch.relu(fo._bn(sd, "backbone.bn1", F.conv2d(q(images), q(sd["backbone.conv1.weight"]), stride=2, padding=3))))
    with torch.no_grad():
        y = _cba(VF.nchw_to_nhwc_bf16(images.cuda()), bb.conv1, bb.bn1, relu_mode=1)
        print(storage, "stem", "%.2e" % rel(VF.nhwc_to_nchw_f32(y), x))
        x = F.max_pool2d(x, 3, 2, 1); y = VF.maxpool3x3s2(y)
        print(storage, "pool", "%.2e" % rel(VF.nhwc_to_nchw_f32(y), x))
        for li, (blocks, stride) in enumerate(zip((3, 4, 23), (1, 2, 2))):
            layer = getattr(bb, "layer%d" % (li + 1))
            for bi in range(blocks):
                x = fo.bottleneck(sd, "backbone.layer%d.%d" % (li + 1, bi), x, stride if bi == 0 else 1, 1, bi == 0, storage)
                y = layer[bi](y)
                if bi in (0, 1, blocks - 1):
                    print(storage, "layer%d.%d" % (li + 1, bi), "%.2e" % rel(VF.nhwc_to_nchw_f32(y), x))
