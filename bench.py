"""bench.py -- forward+backward samples/s of the VL-BERT-base hot path (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" = one forward + backward pass of the encoder hot path over one synthetic batch per GPU
(BASELINE config 2: VL-BERT-base 12L/768, 64 text + 36 precomputed region tokens -> S = 101, batch 64 per GPU,
bf16 tensor-core GEMMs with fp32 accumulation, fp32 master weights re-cast every step, loss = mean of squares of the
last layer).  For N > 1 the batch is sharded (weak scaling, 64 per GPU) and every step ends with the NCCL gradient
all-reduce (overlapped layer by layer with the backward pass) -- the DDP step of common/trainer.py.

One JSON line on stdout (rank 0).  `value` = device-resident inputs; `e2e` = the same step through the public module
API with inputs copied from pinned host memory and the loss read back every step.
"""
import argparse
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

T_TEXT, R_REG, HID, LAYERS, HEADS, INTER, VOCAB = 64, 36, 768, 12, 12, 3072, 30522
S_LEN = T_TEXT + R_REG + 1
FLOP_PER_SAMPLE = 3 * LAYERS * (24 * S_LEN * HID * HID + 4 * S_LEN * S_LEN * HID)  # 52.60 GFLOP


def set_shape(text, regions):
    """--text/--regions: 64/36 = BASELINE config 2 (default, the metric's config); 20/100 = config 3 (VQA shape)."""
    global T_TEXT, R_REG, S_LEN, FLOP_PER_SAMPLE
    T_TEXT, R_REG = text, regions
    S_LEN = T_TEXT + R_REG + 1
    FLOP_PER_SAMPLE = 3 * LAYERS * (24 * S_LEN * HID * HID + 4 * S_LEN * S_LEN * HID)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="samples per GPU")
    ap.add_argument("--text", type=int, default=64, help="text tokens per sample")
    ap.add_argument("--regions", type=int, default=36, help="region tokens per sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only: skip the host-input arm")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    return ap.parse_args()


def make_inputs(B, seed, device, pin=False):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(1000, VOCAB, (B, T_TEXT), generator=g)
    types = torch.zeros(B, T_TEXT, dtype=torch.long)
    tvis = torch.randn(B, T_TEXT, HID, generator=g)
    ovl = torch.randn(B, R_REG, 2 * HID, generator=g)
    tmask = torch.ones(B, T_TEXT, dtype=torch.bool)
    omask = torch.ones(B, R_REG, dtype=torch.bool)
    ts = [ids, types, tvis, tmask, ovl, omask]
    if pin:
        return [t.pin_memory() for t in ts]
    return [t.to(device) for t in ts]


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


def cpu_baseline(steps=2, warmup=1, batch=8, threads=None):
    """The reference's CPU path (oracle port: same fp32 torch ops as common/visual_linguistic_bert.py +
    external/pytorch_pretrained_bert/modeling.py) on the host cores, on a bounded sample of the workload."""
    import vlbert_oracle as vo
    cores = threads or min(os.cpu_count() or 1, 16)  # tools/cpu_threads_probe.py on the 128-core GPU host: 16 threads is the fastest
    torch.set_num_threads(cores)
    cfg = vo.default_config(num_hidden_layers=LAYERS)
    torch.manual_seed(12345)
    model = vo.VisualLinguisticBertOracle(cfg)
    ins = make_inputs(batch, 12345, "cpu")
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
            "sample": "batch %d x %d timed steps of the full 12-layer config-2 shape (S=%d), fp32, torch CPU ops" % (batch, steps, S_LEN),
            "ms_per_step": dt * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_baseline(steps=max(1, min(args.steps, 3)), warmup=max(1, min(args.warmup, 1)), batch=8)
    line = {"impl": "reference", "metric": "samples/sec VL-BERT-base fwd+bwd", "value": cb["value"], "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "VL-BERT-base 12L/768, %d text + %d region tokens (S=%d), fwd+bwd, CPU fp32; bounded sample batch 8" % (T_TEXT, R_REG, S_LEN)},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


_WATCHDOG = None


def _arm_watchdog(seconds):
    """A hang inside a CUDA/NCCL call holds the GIL, so the guard is a separate process: it polls this PID once a second, kills it
    after `seconds`, and goes away by itself as soon as this process is gone (no stale PID is ever signalled)."""
    global _WATCHDOG
    import subprocess
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
    """DRAM bytes per GEMM launch from the committed `ncu --set full` capture of one encoder layer's backward GEMMs
    (dram__bytes_read.sum + dram__bytes_write.sum, mean over the captured launches); None if the capture is absent."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_ncu_full_gemm_v4_backward_layer.csv")
    try:
        import csv
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = sum(float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]] for r in rows[2:])
        return {"bytes_per_launch": tot / max(1, len(rows) - 2), "launches": len(rows) - 2, "source": "profiles/" + os.path.basename(path)}
    except Exception:  # noqa
        return None


def _bounded_teardown(dist):
    """End a multi-rank run without destroy_process_group(): tearing down a communicator whose collectives were captured in
    a CUDA graph hung once (and a hang there would eat the whole GPU lease).  Every rank has already passed the final
    barrier + device synchronize and rank 0 has printed its line, so the process simply exits."""
    sys.stdout.flush()
    sys.stderr.flush()
    _disarm_watchdog()
    os._exit(0)


def main():
    args = parse()
    set_shape(args.text, args.regions)
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
    lib = vlbert_b200._lib.lib()

    cfg = vlbert_b200.default_config(num_hidden_layers=LAYERS)
    torch.manual_seed(12345)
    model = vlbert_b200.VisualLinguisticBert(cfg).to(dev)
    model.visual_ln_text.weight.data.fill_(1.0)
    model.max_length_hint = S_LEN  # all synthetic samples are full length; avoids the per-forward host sync
    B = args.batch
    enc_param_ids = set(id(p) for l in model.encoder.layer for p in l.flat_params())
    other_params = [p for p in model.parameters() if id(p) not in enc_param_ids]
    reducer = vlbert_b200.ddp.attach(model) if world > 1 else None

    def loss_fn(m, *ins):
        out, _ = m(*ins, output_all_encoded_layers=False)
        return (out.float() ** 2).mean()

    def eager_step(ins):
        model.zero_grad(set_to_none=True)
        loss = loss_fn(model, *ins)
        loss.backward()
        if reducer is not None:
            reducer.reduce_params(other_params)
        return loss

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    dev_inputs = make_inputs(B, 12345 + rank, dev)
    # The public step API: vlbert_b200.GraphedStep (CUDA-graph replay of forward+backward) when shapes are static.
    # NCCL collectives are captured with the graph (N > 1) unless --no-graph.
    use_graph = (not args.no_graph) and (world == 1 or os.environ.get("VLB_GRAPH_DDP", "1") == "1")
    graphed = None
    if use_graph:
        try:
            graphed = vlbert_b200.GraphedStep(model, loss_fn, dev_inputs, warmup=3, reducer_params=other_params if reducer else None)
        except Exception as e:  # noqa
            print("[bench] CUDA graph capture failed (%s); running eagerly" % str(e)[:200], file=sys.stderr)
            use_graph = False
            graphed = None
            torch.cuda.synchronize()

    def step(ins):
        if graphed is not None:
            return graphed(*ins)
        return eager_step(ins)

    # ---------------- per-kernel device timing (roofline): a separate eager pass, events on the launch stream ----------
    for _ in range(3):
        eager_step(dev_inputs)
    barrier()
    lib.vlb_profile_enable(1)
    prof_steps = max(2, min(args.steps, 5))
    for _ in range(prof_steps):
        eager_step(dev_inputs)
    barrier()
    lib.vlb_profile_enable(0)
    import ctypes
    pms, pwork, pcnt = (ctypes.c_double * 12)(), (ctypes.c_double * 12)(), (ctypes.c_int64 * 12)()
    vlbert_b200._lib.check(lib.vlb_profile_collect(pms, pwork, pcnt))

    # ---------------- device-resident arm ----------------
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(max(3, args.warmup)):
        step(dev_inputs)
    if sampler:
        sampler.ready.wait(5.0)
    barrier()
    n0 = vlbert_b200._lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if sampler:
        sampler.active = True
    e0.record()
    for _ in range(args.steps):
        step(dev_inputs)
    e1.record()
    barrier()
    if sampler:
        sampler.active = False
    launches_eager = 0
    if graphed is None:
        launches_eager = (vlbert_b200._lib.launch_count() - n0) // max(1, args.steps)
    ms = e0.elapsed_time(e1) / args.steps
    t = torch.tensor([ms], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    # kernels of the library per step (a graph replay launches the same kernels the capture recorded)
    n1 = vlbert_b200._lib.launch_count()
    eager_step(dev_inputs)
    torch.cuda.synchronize()
    launches = vlbert_b200._lib.launch_count() - n1

    # ---------------- end-to-end arm: pinned host inputs -> H2D -> step -> loss read back ----------------
    host_inputs = [make_inputs(B, 777 + rank + i, dev, pin=True) for i in range(2)]
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
            loss = step(ts)
            loss_host.copy_(loss.detach(), non_blocking=True)
        torch.cuda.synchronize()
        return float(loss_host)

    e2e_runs = []
    if args.skip_e2e:
        e2e_ms, last_loss = float("nan"), float("nan")
    else:
        e2e_loop(max(3, args.warmup))
        for _rep in range(2):  # two timed repeats (host-side jitter on shared boxes); both are reported, the better one is `value`
            barrier()
            t0 = torch.cuda.Event(enable_timing=True)
            t1 = torch.cuda.Event(enable_timing=True)
            if sampler:
                sampler.active = True       # the end-to-end repeats are timed regions too: more clock samples
            t0.record()
            wall0 = time.perf_counter()
            last_loss = e2e_loop(args.steps)
            t1.record()
            barrier()
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

    if rank != 0:
        if dist is not None:
            graphed = None
            _bounded_teardown(dist)
        return
    pk, pk_kind = peaks()
    gemm_ms = pms[0] + pms[1] + pms[2]
    gemm_flops = pwork[0] + pwork[1] + pwork[2]
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
    peak_tf = pk.get("bf16_tflops_sustained", pk.get("bf16_tflops"))
    prof = {n: {"ms_per_step": pms[i] / prof_steps, "launches_per_step": pcnt[i] / prof_steps}
            for i, n in enumerate(["gemm_nt", "gemm_nn", "gemm_tn", "mhsa_fwd", "mhsa_bwd", "ln_fwd", "ln_bwd", "other"])}
    line = {
        "metric": "samples/sec VL-BERT-base fwd+bwd", "value": world * B / (ms * 1e-3), "unit": "samples/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "%s: VL-BERT-base 12L/768/12 heads, %d text + %d region tokens (S=%d), "
                               "batch %d per GPU, fwd+bwd (encoder + embedding/packing), loss=mean(out^2)"
                               % ("BASELINE config 2" if (T_TEXT, R_REG) == (64, 36) else "BASELINE config 3 shape" if (T_TEXT, R_REG) == (20, 100) else "custom",
                                  T_TEXT, R_REG, S_LEN, B),
                   "global_batch": world * B, "seq_len": S_LEN, "parallelism": "dp%d" % world,
                   "l2": "no flush needed: per-step working set (saved activations ~2.1 GB + 0.5 GB weights/grads) >> 126 MB L2",
                   "algorithmic_tflop_per_step_per_gpu": B * FLOP_PER_SAMPLE / 1e12},
        "model_flops_tflops": world * B * FLOP_PER_SAMPLE / (ms * 1e-3) / 1e12,
        "e2e": {"value": world * B / (e2e_ms * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_ms, "ms_per_step_runs": e2e_runs, "last_loss": last_loss},
        "gpu_launches": int(launches), "cuda_graph": graphed is not None,
        "roofline": {"bound": "tensor", "kernel": "gemm_kernel<BN,A_MN,B_MN> (all tcgen05 GEMM launches of the step)",
                     "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": (achieved / peak_tf) if achieved else None,
                     "peak_kind": pk_kind + " bf16_tflops_sustained", "traffic": (_ncu_traffic() or {}).get("bytes_per_launch"),
                     "traffic_source": (_ncu_traffic() or {}).get("source"),
                     "share_of_step": gemm_ms / prof_steps / ms if ms > 0 else None,
                     "measured": "CUDA events around every GEMM launch on the launch stream, %d eager steps of the same workload inside this run" % prof_steps},
        "kernel_profile": prof,
        "clocks": sampler.summary() if sampler else None,
    }
    if not args.no_cpu_baseline and world == 1:
        cb = cpu_baseline()
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))
    sys.stdout.flush()
    if dist is not None:
        graphed = None
        _bounded_teardown(dist)


if __name__ == "__main__":
    main()
