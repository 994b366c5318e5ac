"""Timeline of one launch of the CTA-pair GEMM kernel (csrc/gemm_pair192.cuh): every CTA records %globaltimer at
  0 kernel entry | 1 after griddepcontrol.wait | 2 first operand stage landed (leader) | 3 last MMA issued (leader) |
  4 accumulators complete (epilogue warp 2) | 5 rows 128..191 stored | 6 rows 0..127 stored | 7 CTA exit
The launch is the last of a back-to-back series of the same GEMM (as in the step graph), so entry times include the overlap with
the previous launch's tail.  Prints, per stamp, min / median / max over CTAs relative to the earliest kernel entry.

    python tools/gemm_trace.py [M N K] [variant]      variant: plain | resid | gelu | dgelu
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import vlbert_b200
    VF = vlbert_b200.functional
    lib = vlbert_b200._lib.lib()
    M, N, K = [int(a) for a in sys.argv[1:4]] if len(sys.argv) >= 4 else (6464, 768, 768)
    variant = sys.argv[4] if len(sys.argv) > 4 and not sys.argv[4].startswith("--") else "plain"
    dev, bf, f32 = "cuda", torch.bfloat16, torch.float32
    g = torch.Generator(device=dev).manual_seed(0)

    def rnd(*shape, dtype=bf, scale=1.0):
        return (torch.randn(*shape, device=dev, generator=g) * scale).to(dtype)

    rng = torch.tensor([1234, 1], dtype=torch.int64, device=dev)
    drop = VF.DropSite(0.1, 2, rng).with_bits(M, N)
    sets = [dict(a=rnd(M, K), w=rnd(N, K, scale=0.03), wt=rnd(K, N, scale=0.03), b=rnd(N, dtype=f32), o16=torch.empty(M, N, device=dev, dtype=bf),
                 r32=rnd(M, N, dtype=f32), aux=rnd(M, N), z=torch.empty(M, N, device=dev, dtype=bf), mean=rnd(M, dtype=f32, scale=0.1),
                 rstd=rnd(M, dtype=f32).abs() + 0.5, gam=rnd(N, dtype=f32), bet=rnd(N, dtype=f32)) for _ in range(4)]
    fns = {
        "plain": lambda s: VF.gemm(1, s["a"], s["wt"], s["o16"]),
        "bias": lambda s: VF.gemm(0, s["a"], s["w"], s["o16"], bias=s["b"]),
        "resid": lambda s: VF.gemm_bias_residual_f32(s["a"], s["w"], s["b"], s["r32"], ln=(s["mean"], s["rstd"], s["gam"], s["bet"]), drop=drop),
        "gelu": lambda s: VF.gemm(0, s["a"], s["w"], s["o16"], bias=s["b"], act=1, aux=s["z"]),
        "dgelu": lambda s: VF.gemm(1, s["a"], s["wt"], s["o16"], act=3, aux=s["aux"]),
    }
    fn = fns[variant]
    persistent = "--persistent" in sys.argv     # the 148-CTA persistent kernel: per CTA 8 items x (MMA start, MMA end, epilogue start, epilogue end)
    buf = torch.zeros(296 * 32, dtype=torch.int64, device=dev)
    for i in range(8):
        fn(sets[i % 4])
    torch.cuda.synchronize()
    for i in range(6):
        fn(sets[i % 4])
    lib.vlb_debug_gemm_trace(buf.data_ptr())
    fn(sets[2])
    lib.vlb_debug_gemm_trace(None)
    fn(sets[3])
    torch.cuda.synchronize()
    if persistent:
        t = buf.view(-1, 8, 4).cpu().double()
        t = t[t[:, 0, 0] > 0]
        t0 = t[:, 0, 0].min()
        print("M=%d N=%d K=%d variant=%s CTAs=%d (persistent kernel)  env=%s" % (M, N, K, variant, t.shape[0], {k: v for k, v in os.environ.items() if k.startswith("VLB_")}))
        for it in range(8):
            col = t[:, it, :]
            ok = col[:, 0] > 0
            if ok.sum() == 0:
                break
            c = (col[ok] - t0) / 1e3
            print("  item %d (%3d CTAs): MMA start %7.2f  MMA end %7.2f | epilogue start %7.2f  end %7.2f   (medians, us; epilogue stamps of warp 2)" %
                  (it, int(ok.sum()), c[:, 0].median(), c[:, 1].median(), c[:, 2][c[:, 2] > -1e6].median(), c[:, 3].median()))
        return
    t = buf.view(-1, 8)[:296].cpu()
    used = t[:, 0] > 0
    t = t[used].double()
    t0 = t[:, 0].min()
    names = ["entry", "after pdl wait", "first stage landed", "last MMA issued", "accumulators complete", "rows 128.. stored", "rows 0..127 stored", "exit"]
    print("M=%d N=%d K=%d variant=%s CTAs=%d  env=%s" % (M, N, K, variant, t.shape[0], {k: v for k, v in os.environ.items() if k.startswith("VLB_")}))
    for i, n in enumerate(names):
        col = t[:, i]
        col = col[col > 0] - t0
        if col.numel() == 0:
            continue
        print("  %-24s min %8.2f  median %8.2f  max %8.2f us   (%d CTAs)" % (n, col.min() / 1e3, col.median() / 1e3, col.max() / 1e3, col.numel()))


if __name__ == "__main__":
    main()
