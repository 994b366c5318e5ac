"""Bring-up probe for the tcgen05 GEMM: run on a B200 (gpurun), compares against torch fp32 matmul
of the same bf16-rounded operands and (optionally) sweeps the MN-major descriptor geometry."""
import ctypes
import importlib.util
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("vlb_lib", os.path.join(ROOT, "vl-bert_b200", "_lib.py"))
L = importlib.util.module_from_spec(spec)
spec.loader.exec_module(L)
lib = L.lib()


def run(mode, M, N, K, bn=0, split_k=1, act=0, with_bias=False, resid_kind=0, out_kind=0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    dev = "cuda"
    if mode == 0:
        A = torch.randn(M, K, device=dev, generator=g).bfloat16()
        B = torch.randn(N, K, device=dev, generator=g).bfloat16()
        ref = A.float() @ B.float().t()
    elif mode == 1:
        A = torch.randn(M, K, device=dev, generator=g).bfloat16()
        B = torch.randn(K, N, device=dev, generator=g).bfloat16()
        ref = A.float() @ B.float()
    else:
        A = torch.randn(K, M, device=dev, generator=g).bfloat16()
        B = torch.randn(K, N, device=dev, generator=g).bfloat16()
        ref = A.float().t() @ B.float()
    bias = torch.randn(N, device=dev, generator=g) if with_bias else None
    resid = None
    if resid_kind == 1:
        resid = torch.randn(M, N, device=dev, generator=g).bfloat16()
    elif resid_kind == 2:
        resid = torch.randn(M, N, device=dev, generator=g)
    aux = None
    if bias is not None:
        ref = ref + bias
    if act == 1:
        aux = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        zref = ref.clone()
        ref = torch.nn.functional.gelu(ref)
    if resid is not None:
        ref = ref + resid.float()
    if out_kind == 0:
        out = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16)
    elif out_kind == 1:
        out = torch.full((M, N), float("nan"), device=dev, dtype=torch.float32)
    else:
        out = torch.zeros(M, N, device=dev, dtype=torch.float32)
    stream = torch.cuda.current_stream().cuda_stream
    rc = lib.vlb_gemm_bf16(mode, M, N, K, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0),
                           out.data_ptr(), out.stride(0), out_kind,
                           bias.data_ptr() if bias is not None else None,
                           resid.data_ptr() if resid is not None else None,
                           resid.stride(0) if resid is not None else 0, resid_kind, act,
                           aux.data_ptr() if aux is not None else None,
                           aux.stride(0) if aux is not None else 0, 1.0, split_k, bn, stream)
    if rc != 0:
        return "rc=%d %s" % (rc, L.last_error())
    try:
        torch.cuda.synchronize()
    except Exception as e:  # noqa
        return "EXC %s" % (str(e)[:200])
    o = out.float()
    err = (o - ref).norm() / ref.norm()
    mx = (o - ref).abs().max()
    nan = int(torch.isnan(o).sum())
    s = "relL2=%.3e maxabs=%.3e nan=%d" % (err.item(), mx.item(), nan)
    if act == 1:
        zerr = (aux.float() - zref).norm() / zref.norm()
        s += " zrel=%.3e" % zerr.item()
    return s


def sec_nt():
    print("== NT (K-major x K-major) ==")
    for (M, N, K) in [(128, 128, 64), (128, 128, 256), (256, 256, 768), (6464, 768, 768), (6464, 2304, 768),
                      (26, 768, 768), (6464, 3072, 768), (6464, 768, 3072), (1000, 1608, 200)]:
        for bn in (128, 256, 64):
            print("NT", M, N, K, "bn", bn, run(0, M, N, K, bn=bn))
    print("NT epilogues:")
    print(" bias+gelu+aux   ", run(0, 6464, 3072, 768, act=1, with_bias=True))
    print(" bias+resid->f32 ", run(0, 6464, 768, 3072, with_bias=True, resid_kind=1, out_kind=1))
    print(" bias+resid32    ", run(0, 300, 768, 768, with_bias=True, resid_kind=2, out_kind=1))


def sec_geom(geom):
    lib.vlb_debug_gemm_desc(*geom)
    print("NN geom(lbo,sbo,kadv)=", geom, run(1, 256, 256, 256, bn=128))
    print("TN geom(lbo,sbo,kadv)=", geom, run(2, 256, 256, 256, bn=128, out_kind=1))


def sec_mn(geom):
    lib.vlb_debug_gemm_desc(*geom)
    for (M, N, K) in [(6464, 3072, 768), (6464, 768, 3072), (6464, 768, 2304), (26, 768, 768)]:
        for bn in (128, 256, 64):
            print("NN", M, N, K, "bn", bn, run(1, M, N, K, bn=bn))
    print("== TN (both MN-major) ==")
    for (M, N, K) in [(128, 128, 64), (256, 256, 256), (768, 768, 6464), (3072, 768, 6464), (768, 3072, 6464),
                      (2304, 768, 6464), (768, 768, 26)]:
        for bn in (128, 256, 64):
            print("TN", M, N, K, "bn", bn, run(2, M, N, K, bn=bn, out_kind=1))
    print("TN split-K atomic:", run(2, 768, 768, 6464, split_k=4, out_kind=2))
    print("TN split-K atomic:", run(2, 3072, 768, 6464, split_k=2, out_kind=2))


def sec_time(geom, modes):
    print("== timing ==")
    lib.vlb_debug_gemm_desc(*geom)
    for (mode, M, N, K, name) in [(0, 6464, 2304, 768, "qkv"), (0, 6464, 768, 768, "oproj"), (0, 6464, 3072, 768, "ffn1"),
                                  (0, 6464, 768, 3072, "ffn2"), (1, 6464, 768, 3072, "dgrad_ffn1"),
                                  (1, 6464, 3072, 768, "dgrad_ffn2"), (2, 3072, 768, 6464, "wgrad_ffn1"),
                                  (2, 768, 768, 6464, "wgrad_o")]:
        if mode not in modes:
            continue
        for bn in (64, 128, 256):
            dev = "cuda"
            if mode == 0:
                A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
            elif mode == 1:
                A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(K, N, device=dev).bfloat16()
            else:
                A = torch.randn(K, M, device=dev).bfloat16(); B = torch.randn(K, N, device=dev).bfloat16()
            out_kind = 1 if mode == 2 else 0
            out = torch.empty(M, N, device=dev, dtype=torch.float32 if out_kind else torch.bfloat16)
            st = torch.cuda.current_stream().cuda_stream

            def call():
                return lib.vlb_gemm_bf16(mode, M, N, K, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0),
                                         out.data_ptr(), out.stride(0), out_kind, None, None, 0, 0, 0, None, 0,
                                         1.0, 1, bn, st)
            for _ in range(3):
                rc = call()
            if rc != 0:
                print(name, bn, "rc", rc, L.last_error()); continue
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                call()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print("%-11s mode%d %5dx%5dx%5d bn%3d  %.3f ms  %.1f TFLOP/s" % (name, mode, M, N, K, bn, ms, 2.0 * M * N * K / ms / 1e9))
        if mode == 0:
            for _ in range(3):
                torch.matmul(A, B.t())
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                torch.matmul(A, B.t())
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print("%-11s cuBLAS                       %.3f ms  %.1f TFLOP/s" % (name, ms, 2.0 * M * N * K / ms / 1e9))


def main():
    print("device:", torch.cuda.get_device_name(0), "abi", lib.vlb_abi_version(), "argv", sys.argv[1:])
    sec = sys.argv[1]
    geom = tuple(int(x) for x in sys.argv[2:5]) if len(sys.argv) >= 5 else (0, 0, 0)
    if sec == "nt":
        sec_nt()
    elif sec == "geom":
        sec_geom(geom)
    elif sec == "mn":
        sec_mn(geom)
    elif sec == "time_nt":
        sec_time(geom, (0,))
    elif sec == "time_mn":
        sec_time(geom, (1, 2))
    sys.stdout.flush()


if __name__ == "__main__":
    main()
