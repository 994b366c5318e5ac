"""How does the CPU arm (oracle port) scale with threads on this host?  (bounded: 2 layers, batch 8)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import vlbert_oracle as vo
import bench
print("cpu_count", os.cpu_count())
cfg = vo.default_config(num_hidden_layers=2)
model = vo.VisualLinguisticBertOracle(cfg)
ins = bench.make_inputs(8, 1, "cpu")
for th in (8, 16, 32, 64, 128):
    if th > (os.cpu_count() or 1):
        break
    torch.set_num_threads(th)
    ts = []
    for i in range(3):
        t0 = time.perf_counter()
        model.zero_grad()
        out, _ = model(*ins, output_all_encoded_layers=False)
        (out ** 2).mean().backward()
        ts.append(time.perf_counter() - t0)
    print("threads %3d  2-layer B=8 step: %.3f s (min of 3: %.3f)" % (th, ts[-1], min(ts)))
