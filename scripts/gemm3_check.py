"""Accuracy and timing of the bf16x3 tensor-core GEMM (csrc/hs_gemm3_tc.cu) at the linear-layer shapes of BASELINE
configs[1], against an fp64 product, torch fp32, and the cuBLAS TF32 library GEMM (torch, allow_tf32) it replaces.
Run on the GPU box:  python scripts/gemm3_check.py [--quick] [--time]"""
import ctypes as C
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heal_swin_b200._lib import check, current_stream, lib, ptr  # noqa: E402

HBM = 6550.7  # GB/s, MEASURED_PEAKS.json


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def split(w, transposed=False, prec=0):
    w = w.contiguous()
    n, k = w.shape
    rows, cols = (k, n) if transposed else (n, k)
    out = torch.empty((rows, 2 * ((cols + 31) // 32 * 32)), device=w.device, dtype=torch.bfloat16)
    check(lib.hs_weight_split(ptr(w), rows, cols, k, 1 if transposed else 0, 1 if prec == 1 else 0, ptr(out), current_stream()))
    return out


def gemm3(a, ws, bias=None, aux=None, mode=0, drop=0.0, seed=0, d=None, d2=None, colsum=None, prec=0):
    T, K = a.shape
    N = ws.shape[0]
    if d is None:
        d = torch.empty((T, N), device=a.device, dtype=torch.float32)
    if mode == 2 and d2 is None:
        d2 = torch.empty_like(d)
    check(lib.hs_gemm3(ptr(a), ptr(ws), ptr(bias), ptr(aux), ptr(d), ptr(d2), ptr(colsum), T, N, K, mode, prec, C.c_float(drop),
                       C.c_uint64(seed), current_stream()))
    return (d, d2) if mode == 2 else d


def rel(x, ref):
    return ((x.double() - ref).norm() / ref.norm()).item()


def main():
    quick = "--quick" in sys.argv
    do_time = "--time" in sys.argv
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    # (label, T, N, K)
    B = 8
    T0, T1, T2, T3 = B * 196608, B * 49152, B * 12288, B * 3072
    shapes = [("tiny", 300, 96, 96), ("tiny-ragged", 1000, 288, 64)]
    if not quick:
        shapes += [
            ("s0 qkv", T0, 288, 96), ("s0 proj", T0, 96, 96), ("s0 fc1", T0, 384, 96), ("s0 fc2", T0, 96, 384),
            ("s0 qkv dgrad", T0, 96, 288),
            ("s1 qkv", T1, 576, 192), ("s1 proj", T1, 192, 192), ("s1 fc1", T1, 768, 192), ("s1 fc2", T1, 192, 768),
            ("s2 qkv", T2, 1152, 384), ("s2 fc1", T2, 1536, 384), ("s2 fc2", T2, 384, 1536),
            ("s3 qkv", T3, 2304, 768), ("s3 fc1", T3, 3072, 768), ("s3 fc2", T3, 768, 3072),
            ("merge0", T1, 192, 384), ("expand3", T3, 1536, 768), ("final expand", T0, 384, 96), ("concat1", T1, 192, 384),
        ]
    only = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--only=")]
    if only:
        shapes = [sh for sh in shapes if any(o in sh[0] for o in only)]
    dprec = 1 if "--tf32-dgrad" in sys.argv else 0  # time the dgrad + aux column in the single-MMA TF32 mode
    ok = True
    for label, T, N, K in shapes:
        a = torch.randn(T, K, device=dev)
        w = torch.randn(N, K, device=dev) / math.sqrt(K)
        bias = torch.randn(N, device=dev)
        ws = split(w)
        d = gemm3(a, ws, bias)
        torch.cuda.synchronize()
        rows = min(T, 65536)  # fp64 reference on a row sample (first and last rows)
        idx = torch.cat([torch.arange(0, rows // 2, device=dev), torch.arange(T - (rows - rows // 2), T, device=dev)])
        ref = a[idx].double() @ w.double().t() + bias.double()
        e3 = rel(d[idx], ref)
        torch.backends.cuda.matmul.allow_tf32 = False
        e32 = rel(torch.nn.functional.linear(a[idx], w, bias), ref)
        torch.backends.cuda.matmul.allow_tf32 = True
        y_lt = torch.nn.functional.linear(a, w, bias)
        etf = rel(y_lt[idx], ref)
        line = f"{label:14s} T={T:8d} N={N:5d} K={K:5d}: rel err bf16x3 {e3:.2e} | torch fp32 {e32:.2e} | cuBLAS tf32 {etf:.2e}"
        good = e3 < 2e-5
        # transposed split: dgrad operand   dx = dy @ w
        wst = split(w, transposed=True)
        dy = torch.randn(T, N, device=dev)
        dx = gemm3(dy, wst)
        refx = dy[idx].double() @ w.double()
        ex = rel(dx[idx], refx)
        good = good and ex < 2e-5
        line += f" | dgrad {ex:.2e}"
        # mode 1: + aux
        aux = torch.randn(T, K, device=dev)
        dx1 = gemm3(dy, wst, None, aux, mode=1)
        e1 = rel(dx1[idx], refx + aux[idx].double())
        good = good and e1 < 2e-5
        # mode 2: z, h = gelu(z + b)
        z, h = gemm3(a, ws, bias, mode=2)
        refz = ref - bias.double()
        ez = rel(z[idx], refz)
        eh = rel(h[idx], torch.nn.functional.gelu(ref))
        good = good and ez < 2e-5 and eh < 2e-5
        # mode 3: acc * gelu'(aux + bias)
        zz = torch.randn(T, N, device=dev)
        g3 = gemm3(a, ws, bias, zz, mode=3)
        u = (zz[idx].double() + bias.double()).requires_grad_(True)
        torch.nn.functional.gelu(u).sum().backward()
        e3g = rel(g3[idx], refz * u.grad)
        good = good and e3g < 2e-5
        line += f" | add {e1:.2e} gelu z {ez:.2e} h {eh:.2e} gelu' {e3g:.2e}  {'OK' if good else 'FAIL'}"
        ok = ok and good
        if do_time and T >= 1024:
            y32 = torch.empty_like(d)
            t3 = timeit(lambda: gemm3(a, ws, bias, d=d))
            tlt = timeit(lambda: torch.nn.functional.linear(a, w, bias, out=None))
            z2, h2 = torch.empty_like(d), torch.empty_like(d)
            tg = timeit(lambda: gemm3(a, ws, bias, mode=2, d=z2, d2=h2))
            wst_t = split(w, transposed=True, prec=dprec) if dprec else wst
            ta = timeit(lambda: gemm3(dy, wst_t, None, aux, mode=1, d=dx1, prec=dprec))
            tgg = timeit(lambda: gemm3(a, ws, bias, zz, mode=3, d=g3))
            gb = T * (N + K) * 4 / 1e9
            fl = 2.0 * T * N * K / 1e12
            line += (f"\n{'':14s} plain {t3:.3f} ms ({gb / t3 * 1e3:.0f} GB/s = {gb / t3 * 1e3 / HBM:.2f} of HBM, {fl / t3 * 1e3:.0f} "
                     f"TFLOP/s) | cuBLAS tf32 {tlt:.3f} ms | gelu(z,h) {tg:.3f} ms ({T * (2 * N + K) * 4 / 1e6 / tg / HBM:.2f}) | "
                     f"dgrad+aux {ta:.3f} ms ({T * (N + 2 * K) * 4 / 1e6 / ta / HBM:.2f}) | gelu' {tgg:.3f} ms "
                     f"({T * (2 * N + K) * 4 / 1e6 / tgg / HBM:.2f})")
            del y32, z2, h2
        print(line, flush=True)
        del a, w, d, dy, dx, aux, dx1, z, h, zz, g3, y_lt
        torch.cuda.empty_cache()
    print("gemm3_check:", "ALL OK" if ok else "FAILURES")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
