"""Times every GEMM launch of one BertLayer (forward + backward) in isolation at a BASELINE config's token count, with the
epilogues the layer uses, through the C ABI.  CUDA events around `reps` launches that rotate over `sets` operand sets
(so no launch finds its operands in L2), after warm-up.  A/B switches are environment variables read by the library at
load (VLB_TAIL_SPLIT, VLB_EPI_PREFETCH, VLB_FORCE_BN, ...): run the script once per setting.

    python tools/layer_gemm_bench.py [--config 2|3|4] [--reps 24] [--sets 4] [--drop 0.1]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CONFIGS = {2: dict(B=64, S=101, H=768, I=3072), 3: dict(B=64, S=121, H=768, I=3072), 4: dict(B=64, S=165, H=1024, I=4096)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--reps", type=int, default=24)
    ap.add_argument("--sets", type=int, default=4)
    ap.add_argument("--drop", type=float, default=0.1)
    a = ap.parse_args()
    import vlbert_b200
    VF = vlbert_b200.functional
    c = CONFIGS[a.config]
    M, H, I = c["B"] * c["S"], c["H"], c["I"]
    dev = "cuda"
    bf, f32 = torch.bfloat16, torch.float32
    g = torch.Generator(device=dev).manual_seed(0)

    def rnd(*shape, dtype=bf, scale=1.0):
        return (torch.randn(*shape, device=dev, generator=g) * scale).to(dtype)

    rng = torch.tensor([1234, 1], dtype=torch.int64, device=dev)
    drop = VF.DropSite(a.drop, 2, rng).with_bits(M, H) if a.drop > 0 else None
    sets = []
    for _ in range(a.sets):
        s = dict(x=rnd(M, H), wqkv=rnd(3 * H, H, scale=0.03), bqkv=rnd(3 * H, dtype=f32), qkv=torch.empty(M, 3 * H, device=dev, dtype=bf),
                 ctx=rnd(M, H), wo=rnd(H, H, scale=0.03), bo=rnd(H, dtype=f32), a32=torch.empty(M, H, device=dev, dtype=f32),
                 h=rnd(M, H), w1=rnd(I, H, scale=0.03), b1=rnd(I, dtype=f32), u=torch.empty(M, I, device=dev, dtype=bf),
                 z=torch.empty(M, I, device=dev, dtype=bf), w2=rnd(H, I, scale=0.03), b2=rnd(H, dtype=f32),
                 y0=torch.empty(M, H, device=dev, dtype=f32), dy0=rnd(M, H), dz=torch.empty(M, I, device=dev, dtype=bf),
                 db1=torch.zeros(I, device=dev), dh=torch.empty(M, H, device=dev, dtype=bf), da=rnd(M, H),
                 dctx=torch.empty(M, H, device=dev, dtype=bf), dqkv=rnd(M, 3 * H), dx=torch.empty(M, H, device=dev, dtype=bf),
                 gp=rnd(M, I), r32=rnd(M, H, dtype=f32), mean=rnd(M, dtype=f32, scale=0.1), rstd=rnd(M, dtype=f32).abs() + 0.5,
                 gam=rnd(H, dtype=f32), bet=rnd(H, dtype=f32))
        sets.append(s)
    lib = vlbert_b200._lib.lib()

    def colsum_gemm(s):   # dz = (d_y0 W2) o gelu' with the fused bias-gradient column sum: only reachable through the layer call
        VF.gemm(1, s["dy0"], s["w2"], s["dz"], act=3, aux=s["gp"])

    launches = [
        ("fwd qkv     NT N=3H K=H  bias->bf16", 2.0 * M * 3 * H * H, lambda s: VF.gemm(0, s["x"], s["wqkv"], s["qkv"], bias=s["bqkv"])),
        ("fwd attnout NT N=H  K=H  bias+drop+LN-resid(f32)->f32", 2.0 * M * H * H,
         lambda s: VF.gemm_bias_residual_f32(s["ctx"], s["wo"], s["bo"], s["r32"], ln=(s["mean"], s["rstd"], s["gam"], s["bet"]), drop=drop)),
        ("fwd ffn-up  NT N=I  K=H  bias+gelu+aux", 2.0 * M * I * H, lambda s: VF.gemm(0, s["h"], s["w1"], s["u"], bias=s["b1"], act=1, aux=s["z"])),
        ("fwd ffn-dn  NT N=H  K=I  bias+drop+LN-resid(f32)->f32", 2.0 * M * H * I,
         lambda s: VF.gemm_bias_residual_f32(s["u"], s["w2"], s["b2"], s["r32"], ln=(s["mean"], s["rstd"], s["gam"], s["bet"]), drop=drop)),
        ("bwd dz      NN N=I  K=H  x gelu'", 2.0 * M * I * H, colsum_gemm),
        ("bwd dh      NN N=H  K=I  plain->bf16", 2.0 * M * H * I, lambda s: VF.gemm(1, s["dz"], s["w1"], s["dh"])),
        ("bwd dctx    NN N=H  K=H  plain", 2.0 * M * H * H, lambda s: VF.gemm(1, s["da"], s["wo"], s["dctx"])),
        ("bwd dx      NN N=H  K=3H plain->bf16", 2.0 * M * H * 3 * H, lambda s: VF.gemm(1, s["dqkv"], s["wqkv"], s["dx"])),
    ]
    res = {"config": a.config, "M": M, "H": H, "I": I, "env": {k: v for k, v in os.environ.items() if k.startswith("VLB_")}, "gemms": {}}
    tot_us, tot_fl = 0.0, 0.0
    side = torch.cuda.Stream()
    for name, flops, fn in launches:
        # `reps` launches captured in one CUDA graph (no host launch gaps: what the step graph sees), replayed 3 times
        with torch.cuda.stream(side):
            for s in sets:
                fn(s)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                for r in range(a.reps):
                    fn(sets[r % a.sets])
            graph.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(side)
            for _ in range(3):
                graph.replay()
            e1.record(side)
            torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (3 * a.reps)
        res["gemms"][name] = {"us": round(us, 2), "tflops": round(flops / us / 1e6, 1)}
        tot_us += us
        tot_fl += flops
        print("%-50s %8.2f us  %7.1f TFLOP/s" % (name, us, flops / us / 1e6))
    # yardstick: the same shapes as plain bf16 GEMMs through torch.matmul (cuBLASLt heuristics; no epilogue work at all)
    if os.environ.get("VLB_BENCH_CUBLAS", "1") == "1":
        shapes = [("qkv / dx^T", M, 3 * H, H), ("attnout / dctx", M, H, H), ("ffn-up / dz", M, I, H), ("ffn-dn / dh", M, H, I)]
        for name, m, n, k in shapes:
            As = [rnd(m, k) for _ in range(a.sets)]
            Bs = [rnd(k, n, scale=0.03) for _ in range(a.sets)]
            Cs = [torch.empty(m, n, device=dev, dtype=bf) for _ in range(a.sets)]
            with torch.cuda.stream(side):
                for i in range(a.sets):
                    torch.matmul(As[i], Bs[i], out=Cs[i])
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side):
                    for r in range(a.reps):
                        torch.matmul(As[r % a.sets], Bs[r % a.sets], out=Cs[r % a.sets])
                graph.replay()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(side)
                for _ in range(3):
                    graph.replay()
                e1.record(side)
                torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / (3 * a.reps)
            res.setdefault("cublas", {})[name] = {"us": round(us, 2), "tflops": round(2.0 * m * n * k / us / 1e6, 1)}
            print("cuBLAS %-44s %8.2f us  %7.1f TFLOP/s" % ("%s  %dx%dx%d" % (name, m, n, k), us, 2.0 * m * n * k / us / 1e6))
    res["sum_us"] = round(tot_us, 2)
    res["tflops"] = round(tot_fl / tot_us / 1e6, 1)
    print("sum %.1f us  %.1f TFLOP/s  (graph-replayed back-to-back launches of one kind)" % (tot_us, tot_fl / tot_us / 1e6))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
