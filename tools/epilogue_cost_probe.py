"""What the fused epilogues cost on top of a plain bf16 store, per variant, on one GEMM shape of the encoder (default: the
attention-output shape 6464 x 768 x 768, which is a single wave of the pair kernel, i.e. the epilogue is fully exposed).
Graph-replayed back-to-back launches rotating over operand sets, CUDA events; run once per library setting (VLB_* env).

    python tools/epilogue_cost_probe.py [M N K] [--reps 24]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import vlbert_b200
    VF = vlbert_b200.functional
    args = [int(a) for a in sys.argv[1:4]] if len(sys.argv) >= 4 else [6464, 768, 768]
    M, N, K = args
    reps, nsets = 24, 4
    dev, bf, f32 = "cuda", torch.bfloat16, torch.float32
    g = torch.Generator(device=dev).manual_seed(0)

    def rnd(*shape, dtype=bf, scale=1.0):
        return (torch.randn(*shape, device=dev, generator=g) * scale).to(dtype)

    rng = torch.tensor([1234, 1], dtype=torch.int64, device=dev)
    drop = VF.DropSite(0.1, 2, rng).with_bits(M, N)
    sets = [dict(a=rnd(M, K), w=rnd(N, K, scale=0.03), wt=rnd(K, N, scale=0.03), b=rnd(N, dtype=f32), o16=torch.empty(M, N, device=dev, dtype=bf),
                 o32=torch.empty(M, N, device=dev, dtype=f32), r32=rnd(M, N, dtype=f32), r16=rnd(M, N), aux=rnd(M, N),
                 z=torch.empty(M, N, device=dev, dtype=bf), mean=rnd(M, dtype=f32, scale=0.1), rstd=rnd(M, dtype=f32).abs() + 0.5,
                 gam=rnd(N, dtype=f32), bet=rnd(N, dtype=f32)) for _ in range(nsets)]
    variants = [
        ("NT plain -> bf16", lambda s: VF.gemm(0, s["a"], s["w"], s["o16"])),
        ("NT bias -> bf16", lambda s: VF.gemm(0, s["a"], s["w"], s["o16"], bias=s["b"])),
        ("NT plain -> f32 (generic epilogue)", lambda s: VF.gemm(0, s["a"], s["w"], s["o32"])),
        ("NT bias + f32 residual -> f32", lambda s: VF.gemm_bias_residual_f32(s["a"], s["w"], s["b"], s["r32"])),
        ("NT bias + LN-recomputed f32 residual -> f32", lambda s: VF.gemm_bias_residual_f32(s["a"], s["w"], s["b"], s["r32"], ln=(s["mean"], s["rstd"], s["gam"], s["bet"]))),
        ("NT bias + dropout + LN-recomputed f32 residual -> f32", lambda s: VF.gemm_bias_residual_f32(s["a"], s["w"], s["b"], s["r32"], ln=(s["mean"], s["rstd"], s["gam"], s["bet"]), drop=drop)),
        ("NT bias + GELU (+ GELU' saved) -> bf16", lambda s: VF.gemm(0, s["a"], s["w"], s["o16"], bias=s["b"], act=1, aux=s["z"])),
        ("NN plain -> bf16", lambda s: VF.gemm(1, s["a"], s["wt"], s["o16"])),
        ("NN x GELU' (aux read) -> bf16", lambda s: VF.gemm(1, s["a"], s["wt"], s["o16"], act=3, aux=s["aux"])),
    ]
    side = torch.cuda.Stream()
    print("M=%d N=%d K=%d  env=%s" % (M, N, K, {k: v for k, v in os.environ.items() if k.startswith("VLB_")}))
    for name, fn in variants:
        with torch.cuda.stream(side):
            for s in sets:
                fn(s)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                for r in range(reps):
                    fn(sets[r % nsets])
            graph.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(side)
            for _ in range(3):
                graph.replay()
            e1.record(side)
            torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (3 * reps)
        print("%-58s %8.2f us  %7.1f TFLOP/s" % (name, us, 2.0 * M * N * K / us / 1e6))


if __name__ == "__main__":
    main()
