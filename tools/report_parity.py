"""Prints measured parity (relative L2 vs the fp32 CPU oracle) of the drop-in encoder for DESIGN.md: per-layer outputs
and the worst / median parameter-gradient error, for 1 / 2 / 12 layers at the config-2 token shape (B = 8)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import vlbert_b200
import vlbert_oracle as vo
from synth import seeded_state_dict, synth_vlbert_inputs, vlbert_loss

torch.set_num_threads(min(32, os.cpu_count() or 8))


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def run(m, ins, dev, seed):
    ins = [t.to(dev) for t in ins]
    tv = ins[2].clone().requires_grad_(True); ov = ins[4].clone().requires_grad_(True)
    layers, pooled = m(ins[0], ins[1], tv, ins[3], ov, ins[5], output_all_encoded_layers=True)
    m.zero_grad(); vlbert_loss(layers, pooled, seed).backward()
    return layers, pooled, {k: p.grad for k, p in m.named_parameters() if p.grad is not None}, tv.grad, ov.grad


for L in (1, 2, 12):
    cfg = vo.default_config(num_hidden_layers=L)
    ora = vo.VisualLinguisticBertOracle(cfg)
    sd = seeded_state_dict(ora, 100 + L)
    ora.load_state_dict(sd)
    mod = vlbert_b200.VisualLinguisticBert(vlbert_b200.default_config(num_hidden_layers=L)).cuda()
    mod.load_state_dict(sd)
    ins = synth_vlbert_inputs(B=8, T=64, R=36, H=768, vocab=30522, seed=7, ragged=True)
    a = run(mod, ins, "cuda", 3)
    b = run(ora, ins, "cpu", 3)
    outs = [rel(x, y) for x, y in zip(a[0], b[0])]
    gs = sorted((rel(a[2][k], b[2][k]), k) for k in b[2] if not k.endswith("key.bias"))
    print("L=%2d  out rel-L2: first %.2e last %.2e max %.2e | pooled %.2e | d_text_visual %.2e d_object_vl %.2e | param grads: median %.2e max %.2e (%s)"
          % (L, outs[0], outs[-1], max(outs), rel(a[1], b[1]), rel(a[3], b[3]), rel(a[4], b[4]), gs[len(gs) // 2][0], gs[-1][0], gs[-1][1]))
