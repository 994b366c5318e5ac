"""bench.py -- forward+backward samples/s of the VL-BERT hot path (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3|4|5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" = one forward + backward pass of the hot path over one synthetic batch per GPU, in TRAINING mode with the
reference cfgs' dropout (hidden_dropout_prob = attention_probs_dropout_prob = 0.1, obj_downsample Dropout(0.1)), bf16
tensor-core GEMMs with fp32 accumulation, fp32 master weights re-cast every step.  The headline workload is BASELINE
config 2 (VL-BERT-base 12L/768, 64 text + 36 precomputed region tokens -> S = 101, batch 64 per GPU, loss = mean of squares
of the last layer).  For N > 1 the batch is sharded (weak scaling) and every step ends with the NCCL gradient all-reduce
(overlapped layer by layer with the backward pass) -- the DDP step of common/trainer.py.

One JSON line on stdout (rank 0).  `value` = device-resident inputs; `e2e` = the same step through the public module API
with inputs copied from pinned host memory and the loss read back every step; `other_configs` (N = 1) = BASELINE configs
3 (VQA shape + head), 4 (VL-BERT-large, VCR shape) and 5 (end-to-end ResNet-101 front end) measured the same way, shorter;
`gpu_eager_baseline` = the reference's module graph (oracle port, same shapes) run by eager PyTorch on the same B200 in
fp32 and under bf16 autocast; `cpu_baseline` = the same graph on the host cores.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

VOCAB = 30522
WORKLOADS = {
    2: dict(name="BASELINE config 2", hidden=768, layers=12, heads=12, inter=3072, text=64, regions=36, batch=64, head=None,
            what="VL-BERT-base 12L/768/12 heads, 64 text + 36 precomputed region tokens"),
    3: dict(name="BASELINE config 3", hidden=768, layers=12, heads=12, inter=3072, text=20, regions=100, batch=64, head="vqa",
            what="VL-BERT-base 12L/768 + VQA answer head (transform + Linear(768, 3129), BCE), 20 text + 100 region tokens"),
    4: dict(name="BASELINE config 4", hidden=1024, layers=24, heads=16, inter=4096, text=128, regions=36, batch=64, head=None,
            what="VL-BERT-large 24L/1024/16 heads, VCR Q->A shape: 128 text + 36 region tokens, 16 samples x 4 answer choices = 64 sequences"),
    6: dict(name="BASELINE config 2 + pre-training loss", hidden=768, layers=12, heads=12, inter=3072, text=64, regions=36, batch=64, head="mlm",
            what="VL-BERT-base 12L/768 + masked-LM head (transform + tied decoder [30522, 768] + cross-entropy on 10 labelled positions per sample, fused loss head), 64 text + 36 region tokens"),
    5: dict(name="BASELINE config 5", hidden=768, layers=12, heads=12, inter=3072, text=64, regions=36, batch=8, head="frontend",
            what="end to end: ResNet-101-C4 + RoIAlign + res5 head on 8 synthetic 600x1000 images (36 boxes each) feeding VL-BERT-base 12L/768"),
}
P_DROP = 0.1


def seq_len(w):
    return w["text"] + w["regions"] + 1


def encoder_flop_per_sample(w):
    """fwd+bwd = 3 x L x (24 S H^2 + 4 S^2 H)"""
    S, H = seq_len(w), w["hidden"]
    return 3 * w["layers"] * (24 * S * H * H + 4 * S * S * H)


def frontend_flop_per_image(regions):
    """backbone C4 166.1 GFLOP fwd (conv1 + layer1 frozen: forward only = 18.8; layer2/3 x3 = fwd + dgrad +
    wgrad) + res5 head 5.857 GFLOP per RoI x3"""
    frozen, trainable = 2 * (1.41 + 7.99) * 1e9, 2 * (11.37 + 62.29) * 1e9
    return frozen + 3 * trainable + 3 * regions * 5.857e9


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(WORKLOADS), help="BASELINE config of the headline line (default 2: the metric's)")
    ap.add_argument("--batch", type=int, default=0, help="samples per GPU (0 = the config's)")
    ap.add_argument("--dropout", type=float, default=P_DROP, help="hidden / attention dropout probability (reference cfgs: 0.1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only: skip the host-input arm")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--with-optimizer", action="store_true", help="also run vlbert_b200.optim.FusedAdamW (+ global-norm clip) in every step")
    return ap.parse_args()


def make_inputs(w, B, seed, device, pin=False):
    """synthetic batch of workload `w` on the host (pinned) or on `device`"""
    g = torch.Generator().manual_seed(seed)
    T, R, H = w["text"], w["regions"], w["hidden"]
    ids = torch.randint(1000, VOCAB, (B, T), generator=g)
    types = torch.zeros(B, T, dtype=torch.long)
    tvis = torch.randn(B, T, H, generator=g)
    tmask = torch.ones(B, T, dtype=torch.bool)
    omask = torch.ones(B, R, dtype=torch.bool)
    if w["head"] == "frontend":
        Hh, Ww = 600, 1000
        images = torch.randn(B, 3, Hh, Ww, generator=g)
        x1 = torch.rand(B, R, generator=g) * Ww * 0.6
        y1 = torch.rand(B, R, generator=g) * Hh * 0.6
        boxes = torch.stack((x1, y1, x1 + 32 + torch.rand(B, R, generator=g) * Ww * 0.36, y1 + 32 + torch.rand(B, R, generator=g) * Hh * 0.36), -1)
        im_info = torch.tensor([[float(Ww), float(Hh), 1.0, 1.0]] * B)
        ts = [images, boxes, omask, im_info, ids, types, tvis, tmask]
    else:
        ovl = torch.randn(B, R, 2 * H, generator=g)
        ts = [ids, types, tvis, tmask, ovl, omask]
        if w["head"] == "vqa":
            ts.append(torch.rand(B, 3129, generator=g).pow(8))      # soft answer scores in [0, 1], mostly ~0
        if w["head"] == "mlm":      # 10 masked positions per sample (~15 % of 64 text tokens), label = the original token id
            labels = torch.full((B, T), -1, dtype=torch.long)
            for bi in range(B):
                pos = torch.randperm(T, generator=g)[:10]
                labels[bi, pos] = ids[bi, pos]
            ts.append(labels)
    if pin:
        return [t.pin_memory() for t in ts]
    return [t.to(device) for t in ts]


def build_workload(w, device, p_drop, args_batch=0):
    """-> (module, loss_fn(module, *inputs)) with the library's modules; module.vlbert is the encoder"""
    import vlbert_b200
    import torch.nn as nn
    from vlbert_b200.modules import BertPredictionHeadTransform
    cfg = vlbert_b200.default_config(hidden_size=w["hidden"], num_hidden_layers=w["layers"], num_attention_heads=w["heads"],
                                     intermediate_size=w["inter"], visual_size=w["hidden"], hidden_dropout_prob=p_drop,
                                     attention_probs_dropout_prob=p_drop)

    class Pipeline(nn.Module):
        def __init__(self):
            super().__init__()
            if w["head"] == "mlm":   # pretrain/modules/resnet_vlbert_for_pretraining.py:57-65 (MLM task only)
                self.vlbert = vlbert_b200.VisualLinguisticBertForPretraining(cfg, with_rel_head=False, with_mvrc_head=False)
            else:
                self.vlbert = vlbert_b200.VisualLinguisticBert(cfg)
            self.vlbert.visual_ln_text.weight.data.fill_(1.0)
            self.vlbert.max_length_hint = seq_len(w)      # all synthetic samples are full length; avoids the per-forward host sync
            if w["head"] == "vqa":       # vqa/modules/resnet_vlbert_for_vqa.py:62-70 (CLASSIFIER_TYPE "mlm"), :246-249
                self.final_mlp = nn.Sequential(BertPredictionHeadTransform(cfg), nn.Dropout(p_drop), nn.Linear(w["hidden"], 3129))
            if w["head"] == "frontend":
                from synth import frontend_config
                self.image_feature_extractor = vlbert_b200.FastRCNN(frontend_config(101), True, w["hidden"], False)
                self.image_feature_extractor.compact_rois = False
                self.image_feature_extractor.obj_downsample[0].p = p_drop
                with torch.no_grad():   # BN statistics of a trained backbone keep activations O(1) through 33 blocks
                    for mod in self.image_feature_extractor.modules():
                        if isinstance(mod, nn.BatchNorm2d):
                            mod.weight.uniform_(0.3, 0.6)
                            mod.running_var.uniform_(0.5, 1.5)
                self.object_linguistic_embeddings = nn.Embedding(1, w["hidden"])

        def train(self, mode=True):
            super().train(mode)
            if w["head"] == "frontend":
                self.image_feature_extractor.bn_eval()          # IMAGE_FROZEN_BN: BatchNorm stays in eval mode (fast_rcnn.py:122-126)
            return self

    torch.manual_seed(12345)
    model = Pipeline().to(device)
    model.train()

    if w["head"] == "vqa":
        ans_pos = w["text"] - 2

        def loss_fn(m, ids, types, tvis, tmask, ovl, omask, label):
            hidden, _ = m.vlbert(ids, types, tvis, tmask, ovl, omask, output_all_encoded_layers=False)
            logits = m.final_mlp(hidden[:, ans_pos])
            return torch.nn.functional.binary_cross_entropy_with_logits(logits, label) * label.size(1)
    elif w["head"] == "mlm":
        n_lab = (args_batch or w["batch"]) * 10

        def loss_fn(m, ids, types, tvis, tmask, ovl, omask, labels):
            text_out, _, _ = vlbert_b200.VisualLinguisticBert.forward(m.vlbert, ids, types, tvis, tmask, ovl, omask,
                                                                       output_all_encoded_layers=False, output_text_and_object_separately=True)
            loss, _, _ = m.vlbert.mlm_loss(text_out, labels, max_labelled=n_lab)
            return loss
    elif w["head"] == "frontend":
        def loss_fn(m, images, boxes, omask, im_info, ids, types, tvis, tmask):
            obj = m.image_feature_extractor(images=images, boxes=boxes, box_mask=omask, im_info=im_info)["obj_reps"]
            ling = m.object_linguistic_embeddings.weight[0].view(1, 1, -1).expand_as(obj)
            out, _ = m.vlbert(ids, types, tvis, tmask, torch.cat((obj, ling), -1), omask, output_all_encoded_layers=False)
            return (out.float() ** 2).mean()
    else:
        def loss_fn(m, ids, types, tvis, tmask, ovl, omask):
            out, _ = m.vlbert(ids, types, tvis, tmask, ovl, omask, output_all_encoded_layers=False)
            return (out.float() ** 2).mean()
    return model, loss_fn


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons DURING the timed region (NVML, ~10 ms period; nvidia-smi as a fallback)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.rows, self.stop_flag, self.index = [], False, index
        self.max_mhz = None
        self.active = False      # samples are kept only while a timed region is running (the thread itself starts earlier: NVML init takes ~0.1 s)
        self.ready = threading.Event()

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
            self.ready.set()
            while not self.stop_flag:
                if not self.active:
                    time.sleep(0.002)
                    continue
                mhz = int(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                try:
                    mask = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    mask = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                if self.active:
                    self.rows.append((mhz, [n for b, n in names.items() if mask & b]))
                time.sleep(0.005)
            return
        except Exception:
            pass
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        self.ready.set()
        while not self.stop_flag:
            if not self.active:
                time.sleep(0.002)
                continue
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                c = [x.strip() for x in out.split(",")]
                if c and c[0].isdigit():
                    self.max_mhz = int(c[1]) if c[1].isdigit() else self.max_mhz
                    self.rows.append((int(c[0]), [n for i, n in enumerate(names) if c[2 + i].lower().startswith("active")]))
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unsampled"], "samples": 0}
        sm = sorted(r[0] for r in self.rows)
        reasons = sorted(set(n for r in self.rows for n in r[1]))
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(self.rows)}


# ----------------------------------------------------------------------------------------------------------------------
# the reference's own module graph (oracle port: the fp32 torch ops of common/visual_linguistic_bert.py +
# external/pytorch_pretrained_bert/modeling.py, dropout through torch's generator like the reference) -- on the host
# cores (cpu_baseline, --impl reference) and in eager PyTorch on the GPU (gpu_eager_baseline).  Checker code: never on
# the product path.
# ----------------------------------------------------------------------------------------------------------------------
def oracle_model(w, p_drop, device):
    import vlbert_oracle as vo
    cfg = vo.default_config(hidden_size=w["hidden"], num_hidden_layers=w["layers"], num_attention_heads=w["heads"],
                            intermediate_size=w["inter"], visual_size=w["hidden"], hidden_dropout_prob=p_drop,
                            attention_probs_dropout_prob=p_drop)
    torch.manual_seed(12345)
    model = vo.VisualLinguisticBertOracle(cfg).to(device)
    model.train()
    model.dropout_state = ("torch",) if p_drop > 0 else None
    return model


def encoder_only(w):
    """the encoder workload the CPU / eager-GPU arms run: same token shape, batch and depth; the VQA head (a [B, 768] MLP) and the
    convolutional front end are not part of those arms"""
    e = dict(w)
    e["head"] = None
    return e


def cpu_baseline(w, p_drop, steps=1, warmup=1, batch=None, threads=None):
    cores = threads or min(os.cpu_count() or 1, 16)  # tools/cpu_threads_probe.py on the 128-core GPU host: 16 threads is the fastest
    torch.set_num_threads(cores)
    w = encoder_only(w)
    batch = batch or w["batch"]
    model = oracle_model(w, p_drop, "cpu")
    ins = make_inputs(w, batch, 12345, "cpu")
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        model.zero_grad()
        out, _ = model(*ins, output_all_encoded_layers=False)
        (out.float() ** 2).mean().backward()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    return {"value": batch / dt, "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": "%d timed step(s) after %d warm-up of the FULL workload (batch %d, %d layers, S=%d, dropout %.2g through torch's generator), "
                      "fp32, torch CPU ops on %d threads" % (steps, warmup, batch, w["layers"], seq_len(w), p_drop, cores),
            "ms_per_step": dt * 1e3}


def gpu_eager_baseline(w, p_drop, device, steps=3, warmup=2):
    """BASELINE.md 3.4 / SURVEY 8(d): the same-box GPU comparator -- the reference's module graph executed by eager PyTorch
    (cuBLAS / ATen kernels) on this B200, fp32 and under torch.autocast(bfloat16), CUDA-event timed."""
    w = encoder_only(w)
    out = {"what": "oracle port of the reference modules (same graph, shapes, dropout p=%.2g via torch's generator) in eager PyTorch on this GPU" % p_drop}
    model = oracle_model(w, p_drop, device)
    ins = make_inputs(w, w["batch"], 12345, device)
    tf32_prev = torch.backends.cuda.matmul.allow_tf32
    for name, ctx in (("fp32", None), ("bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            def one():
                model.zero_grad(set_to_none=True)
                if ctx is not None:
                    with ctx:
                        o, _ = model(*ins, output_all_encoded_layers=False)
                else:
                    o, _ = model(*ins, output_all_encoded_layers=False)
                (o.float() ** 2).mean().backward()
            for _ in range(warmup):
                one()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                one()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"ms_per_step": ms, "samples_per_s": w["batch"] / ms * 1e3,
                         "model_flops_tflops": w["batch"] * encoder_flop_per_sample(w) / ms / 1e9}
        except Exception as e:  # noqa
            out[name] = {"error": str(e)[:200]}
    torch.backends.cuda.matmul.allow_tf32 = tf32_prev
    del model
    torch.cuda.empty_cache()
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.config]
    if args.batch:
        w = dict(w, batch=args.batch)
    cb = cpu_baseline(w, args.dropout, steps=max(1, min(args.steps, 2)), warmup=1)
    line = {"impl": "reference", "metric": "samples/sec VL-BERT-base fwd+bwd", "value": cb["value"], "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s: %s (S=%d), batch %d, fwd+bwd, dropout %.2g, CPU fp32 (the reference's module graph, oracle port)"
                                   % (w["name"], w["what"], seq_len(w), w["batch"], args.dropout)},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


_WATCHDOG = None


def _arm_watchdog(seconds):
    """A hang inside a CUDA/NCCL call holds the GIL, so the guard is a separate process: it polls this PID once a second, kills it
    after `seconds`, and goes away by itself as soon as this process is gone (no stale PID is ever signalled)."""
    global _WATCHDOG
    script = "n=0; while kill -0 %d 2>/dev/null; do sleep 1; n=$((n+1)); if [ $n -ge %d ]; then kill -9 %d; exit 0; fi; done" % (
        os.getpid(), int(seconds), os.getpid())
    try:
        _WATCHDOG = subprocess.Popen(["sh", "-c", script], stdin=subprocess.DEVNULL, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL,
                                     start_new_session=True)
    except Exception:  # noqa
        _WATCHDOG = None


def _disarm_watchdog():
    global _WATCHDOG
    if _WATCHDOG is not None:
        try:
            _WATCHDOG.kill()
        except Exception:  # noqa
            pass
        _WATCHDOG = None


def _ncu_traffic():
    """DRAM bytes per GEMM launch from the committed `ncu --set full` capture of one encoder layer's GEMMs
    (dram__bytes_read.sum + dram__bytes_write.sum, mean over the captured launches); None if the capture is absent."""
    for name in ("r02_ncu_gemm_launches_v2_pair192.csv", "r02_ncu_full_gemm_layer.csv", "r01_ncu_full_gemm_v4_backward_layer.csv"):
        path = os.path.join(ROOT, "profiles", name)
        try:
            import csv
            rows = list(csv.reader(open(path)))
            hdr, units = rows[0], rows[1]
            ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tot = sum(float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]] for r in rows[2:])
            return {"bytes_per_launch": tot / max(1, len(rows) - 2), "launches": len(rows) - 2, "source": "profiles/" + name}
        except Exception:  # noqa
            continue
    return None


PROF_NAMES = ["gemm_nt", "gemm_nn", "gemm_tn", "mhsa_fwd", "mhsa_bwd", "ln_fwd", "ln_bwd", "other", "im2col", "col2im", "conv_elt", "roi_nhwc"]


class Runner(object):
    """One workload on this rank: model, device-resident inputs, eager and CUDA-graph step, per-kernel profile."""

    def __init__(self, w, args, dev, rank, world, dist):
        import vlbert_b200
        self.vb, self.w, self.args, self.dev, self.rank, self.world, self.dist = vlbert_b200, w, args, dev, rank, world, dist
        self.B = args.batch or w["batch"]
        self.model, self.loss_fn = build_workload(w, dev, args.dropout, args.batch)
        enc_ids = set(id(p) for l in self.model.vlbert.encoder.layer for p in l.flat_params())
        self.other_params = [p for p in self.model.parameters() if id(p) not in enc_ids and p.requires_grad]
        self.reducer = None
        if world > 1:
            self.reducer = vlbert_b200.ddp.attach(self.model.vlbert)
            self.model._grad_reducer = self.reducer
        self.opt = None
        if args.with_optimizer:
            self.opt = vlbert_b200.optim.FusedAdamW([p for p in self.model.parameters() if p.requires_grad], lr=1e-5, weight_decay=1e-4,
                                                    max_grad_norm=1.0)
        self.dev_inputs = make_inputs(w, self.B, 12345 + rank, dev)
        self.graphed = None
        use_graph = (not args.no_graph) and (world == 1 or os.environ.get("VLB_GRAPH_DDP", "1") == "1") and self.opt is None
        if use_graph:
            try:
                self.graphed = vlbert_b200.GraphedStep(self.model, self.loss_fn, self.dev_inputs, warmup=3,
                                                       reducer_params=self.other_params if self.reducer else None)
            except Exception as e:  # noqa
                print("[bench] CUDA graph capture failed (%s); running eagerly" % str(e)[:200], file=sys.stderr)
                self.graphed = None
                torch.cuda.synchronize()

    def eager_step(self, ins):
        self.model.zero_grad(set_to_none=self.opt is None)
        loss = self.loss_fn(self.model, *ins)
        loss.backward()
        if self.reducer is not None:
            self.reducer.reduce_params(self.other_params)
        if self.opt is not None:
            self.opt.step()
        return loss

    def step(self, ins):
        if self.graphed is not None:
            return self.graphed(*ins)
        return self.eager_step(ins)

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize()

    def profile(self, steps):
        """per-kernel device time of eager steps (CUDA events around every library launch, on the launch stream)"""
        lib = self.vb._lib.lib()
        for _ in range(3):
            self.eager_step(self.dev_inputs)
        self.barrier()
        lib.vlb_profile_enable(1)
        for _ in range(steps):
            self.eager_step(self.dev_inputs)
        self.barrier()
        lib.vlb_profile_enable(0)
        pms, pwork, pcnt = (ctypes.c_double * 12)(), (ctypes.c_double * 12)(), (ctypes.c_int64 * 12)()
        self.vb._lib.check(lib.vlb_profile_collect(pms, pwork, pcnt))
        return [pms[i] / steps for i in range(12)], [pwork[i] / steps for i in range(12)], [pcnt[i] / steps for i in range(12)]

    def timed(self, steps, warmup, sampler=None):
        for _ in range(max(3, warmup)):
            self.step(self.dev_inputs)
        if sampler:
            sampler.ready.wait(5.0)
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        if sampler:
            sampler.active = True
        e0.record()
        for _ in range(steps):
            self.step(self.dev_inputs)
        e1.record()
        self.barrier()
        if sampler:
            sampler.active = False
        ms = e0.elapsed_time(e1) / steps
        t = torch.tensor([ms], device=self.dev)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.item()

    def launches_per_step(self):
        n1 = self.vb._lib.launch_count()
        self.eager_step(self.dev_inputs)
        torch.cuda.synchronize()
        return self.vb._lib.launch_count() - n1

    def algorithmic_flop_per_step(self):
        f = self.B * encoder_flop_per_sample(self.w)
        if self.w["head"] == "frontend":
            f += self.B * frontend_flop_per_image(self.w["regions"])
        if self.w["head"] == "mlm":      # transform + tied decoder on the 10 labelled rows per sample, fwd + dgrad + wgrad
            f += 3 * 2 * self.B * 10 * (self.w["hidden"] * self.w["hidden"] + self.w["hidden"] * VOCAB)
        return f

    def roofline(self, pms, pwork, ms, pk, pk_kind, prof_steps):
        gemm_ms = pms[0] + pms[1] + pms[2]
        gemm_flops = pwork[0] + pwork[1] + pwork[2]
        achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
        peak_tf = pk.get("bf16_tflops_sustained", pk.get("bf16_tflops"))
        tr = _ncu_traffic() or {}
        return {"bound": "tensor", "kernel": "gemm_kernel<BN,A_MN,B_MN,EPI> / gemm_grouped_tn_kernel (all tcgen05 GEMM launches of the step)",
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": (achieved / peak_tf) if achieved else None,
                "peak_kind": pk_kind + " bf16_tflops_sustained (kernel timed inside a long step)",
                "traffic": tr.get("bytes_per_launch"), "traffic_source": tr.get("source"),
                "share_of_step": gemm_ms / ms if ms > 0 else None,
                "measured": "CUDA events around every GEMM launch on the launch stream, %d eager steps of the same workload inside this run" % prof_steps}

    def close(self):
        self.graphed = None
        self.model = None
        self.opt = None
        torch.cuda.empty_cache()


def optimizer_roofline(run, pk, iters=20):
    """SURVEY 8(f) rank 2, measured on its own: vlbert_b200.optim.FusedAdamW (AdamW of common/nlp/bert/optimization.py:129-187 +
    the trainer's global-norm clip, common/trainer.py:139-147) over every parameter of the workload, two launches per step.
    HBM roofline: per parameter 4 B gradient for the norm + 16 B read (param, grad, two moments) + 12 B written."""
    params = [p for p in run.model.parameters() if p.requires_grad and p.grad is not None]
    if not params:
        run.eager_step(run.dev_inputs)
        params = [p for p in run.model.parameters() if p.requires_grad and p.grad is not None]
    n = sum(p.numel() for p in params)
    opt = run.vb.optim.FusedAdamW(params, lr=1e-6, weight_decay=1e-4, max_grad_norm=1.0)
    for _ in range(3):
        opt.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        opt.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    gbs = n * 32.0 / (ms * 1e-3) / 1e9
    return {"what": "FusedAdamW.step() incl. global-norm clip, all %d parameter tensors (%.1f M parameters), timed alone" % (len(params), n / 1e6),
            "ms_per_step": ms, "bytes_per_param": 32, "achieved_gbs": gbs, "peak_gbs": pk.get("hbm_gbs"),
            "frac": gbs / pk["hbm_gbs"] if pk.get("hbm_gbs") else None, "launches_per_step": 2}


def workload_text(w, B, p_drop, with_opt):
    return "%s: %s (S=%d), batch %d per GPU, fwd+bwd in training mode (dropout p=%.2g at every reference site), %s%s" % (
        w["name"], w["what"], seq_len(w), B, p_drop,
        {None: "loss=mean(out^2)", "vqa": "BCE answer loss", "mlm": "masked-LM cross-entropy (fused loss head)",
         "frontend": "loss=mean(out^2) of the encoder fed by the front end"}[w["head"]],
        ", + FusedAdamW step with global-norm clip" if with_opt else "")


def _bounded_teardown(dist):
    """End a multi-rank run.  destroy_process_group() after graph-captured NCCL collectives needs the graphs (and their
    captured communicator work) released first and the device idle; a watchdog bounds the call so that a hang cannot eat the
    GPU lease, and only then falls back to a plain exit."""
    sys.stdout.flush()
    sys.stderr.flush()
    done = threading.Event()

    def guard():      # c10d's bindings release the GIL, so this thread still runs if the teardown blocks inside NCCL
        if not done.wait(float(os.environ.get("VLB_TEARDOWN_TIMEOUT_S", "30"))):
            print("[bench] destroy_process_group did not return; exiting", file=sys.stderr)
            sys.stderr.flush()
            _disarm_watchdog()
            os._exit(0)

    threading.Thread(target=guard, daemon=True).start()
    try:
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()
    except Exception as e:  # noqa
        print("[bench] destroy_process_group failed: %s" % str(e)[:200], file=sys.stderr)
    done.set()
    _disarm_watchdog()


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        _arm_watchdog(int(os.environ.get("VLB_BENCH_WATCHDOG_S", "420")))   # a multi-rank hang must not outlive the GPU lease
        dist.init_process_group("nccl", device_id=dev)
    import vlbert_b200
    w = WORKLOADS[args.config]
    run = Runner(w, args, dev, rank, world, dist)
    B = run.B
    pk, pk_kind = peaks()

    prof_steps = max(2, min(args.steps, 5))
    pms, pwork, pcnt = run.profile(prof_steps)

    # ---------------- device-resident arm ----------------
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms = run.timed(args.steps, args.warmup, sampler)
    launches = run.launches_per_step()

    # ---------------- end-to-end arm: pinned host inputs -> H2D -> step -> loss read back ----------------
    host_inputs = [make_inputs(w, B, 777 + rank + i, dev, pin=True) for i in range(2)]
    h2d = sum(t_.numel() * t_.element_size() for t_ in host_inputs[0])
    copy_stream = torch.cuda.Stream()
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            ts = [t_.to(dev, non_blocking=True) for t_ in host_inputs[i & 1]]
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ts, ev

    def e2e_loop(n):
        # every step: H2D of that step's inputs from pinned memory (prefetched on a copy stream one step ahead), the step
        # through the public API, and a D2H read of the loss.
        nxt = prefetch(0)
        for i in range(n):
            ts, ev = nxt
            torch.cuda.current_stream().wait_event(ev)
            for t_ in ts:
                t_.record_stream(torch.cuda.current_stream())
            if i + 1 < n:
                nxt = prefetch(i + 1)
            loss = run.step(ts)
            loss_host.copy_(loss.detach(), non_blocking=True)
        torch.cuda.synchronize()
        return float(loss_host)

    e2e_runs = []
    if args.skip_e2e:
        e2e_ms, last_loss = float("nan"), float("nan")
    else:
        e2e_loop(max(3, args.warmup))
        for _rep in range(2):  # two timed repeats (host-side jitter on shared boxes); both are reported, the better one is `value`
            run.barrier()
            t0 = torch.cuda.Event(enable_timing=True)
            t1 = torch.cuda.Event(enable_timing=True)
            if sampler:
                sampler.active = True       # the end-to-end repeats are timed regions too: more clock samples
            t0.record()
            wall0 = time.perf_counter()
            last_loss = e2e_loop(args.steps)
            t1.record()
            run.barrier()
            if sampler:
                sampler.active = False
            e2e_runs.append(max(t0.elapsed_time(t1), (time.perf_counter() - wall0) * 1e3) / args.steps)
        e2e_ms = min(e2e_runs)
    if sampler:
        sampler.stop_flag = True
    t = torch.tensor([e2e_ms], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = t.item()
    try:
        run.model.vlbert.check_errors()     # deferred index check of the graph-friendly path (one sync, outside the timed regions)
    except Exception as e:  # noqa
        raise SystemExit("bench.py: the embedding kernels flagged bad indices: %s" % e)
    used_graph = run.graphed is not None
    flop_step = run.algorithmic_flop_per_step()
    optimizer_rec = None
    if world == 1:
        try:
            optimizer_rec = optimizer_roofline(run, pk)
        except Exception as e:  # noqa
            optimizer_rec = {"error": str(e)[:200]}
    roof = run.roofline(pms, pwork, ms, pk, pk_kind, prof_steps)
    run.close()

    if rank != 0:
        if dist is not None:
            _bounded_teardown(dist)
        return
    prof = {n: {"ms_per_step": pms[i], "launches_per_step": pcnt[i]} for i, n in enumerate(PROF_NAMES) if pcnt[i] > 0}
    line = {
        "metric": "samples/sec VL-BERT-base fwd+bwd", "value": world * B / (ms * 1e-3), "unit": "samples/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload_text(w, B, args.dropout, args.with_optimizer),
                   "global_batch": world * B, "seq_len": seq_len(w), "parallelism": "dp%d" % world,
                   "dropout": {"hidden_dropout_prob": args.dropout, "attention_probs_dropout_prob": args.dropout,
                               "mode": "training: Philox masks regenerated in backward, none stored"},
                   "l2": "no flush needed: per-step working set (saved activations > 2 GB + weights/grads) >> 126 MB L2",
                   "algorithmic_tflop_per_step_per_gpu": flop_step / 1e12},
        "model_flops_tflops": world * flop_step / (ms * 1e-3) / 1e12,
        "e2e": {"value": world * B / (e2e_ms * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_ms, "ms_per_step_runs": e2e_runs, "last_loss": last_loss},
        "gpu_launches": int(launches), "cuda_graph": used_graph,
        "roofline": roof,
        "kernel_profile": prof,
        "clocks": sampler.summary() if sampler else None,
    }
    if optimizer_rec is not None:
        line["optimizer"] = optimizer_rec
    if world == 1 and not args.no_other_configs:
        others = {}
        for c in sorted(WORKLOADS):
            if c == args.config:
                continue
            try:
                wo = WORKLOADS[c]
                a2 = argparse.Namespace(**vars(args))
                a2.batch = 0
                r2 = Runner(wo, a2, dev, rank, world, dist)
                ps = 2
                pm2, pw2, pc2 = r2.profile(ps)
                n_steps = 10 if c != 5 else 5
                ms2 = r2.timed(n_steps, 3)
                roof2 = r2.roofline(pm2, pw2, ms2, pk, pk_kind, ps)
                unit = "images/s" if wo["head"] == "frontend" else ("sequences/s" if c == 4 else "samples/s")
                others["config%d" % c] = {
                    "workload": workload_text(wo, r2.B, args.dropout, args.with_optimizer), "value": r2.B / ms2 * 1e3, "unit": unit,
                    "ms_per_step": ms2, "steps": n_steps, "cuda_graph": r2.graphed is not None,
                    "algorithmic_tflop_per_step": r2.algorithmic_flop_per_step() / 1e12,
                    "model_flops_tflops": r2.algorithmic_flop_per_step() / ms2 / 1e9,
                    "roofline": {k: roof2[k] for k in ("achieved", "peak", "unit", "frac", "share_of_step")},
                    "kernel_ms_per_step": {n: round(pm2[i], 4) for i, n in enumerate(PROF_NAMES) if pc2[i] > 0}}
                r2.close()
                del r2
            except Exception as e:  # noqa
                others["config%d" % c] = {"error": str(e)[:300]}
                torch.cuda.empty_cache()
        line["other_configs"] = others
    if world == 1 and not args.no_gpu_eager:
        line["gpu_eager_baseline"] = gpu_eager_baseline(w, args.dropout, dev)
    if not args.no_cpu_baseline and world == 1:
        cb = cpu_baseline(w, args.dropout)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))
    sys.stdout.flush()
    if dist is not None:
        _bounded_teardown(dist)


if __name__ == "__main__":
    main()
