"""Timing of the fused MLP backward kernel against the pair it replaces (library dgrad GEMM + hs_bias_gelu_bwd) at the
bench stages.  Run on the GPU box: python scripts/mlp_check.py"""
import ctypes as C
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heal_swin_b200._lib import check, current_stream, lib, ptr  # noqa: E402


def timeit(fn, n=20):
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


def main():
    torch.backends.cuda.matmul.allow_tf32 = True
    dev = torch.device("cuda:0")
    for T, Cc in [(8 * 196608, 96), (8 * 49152, 192)]:
        J = 4 * Cc
        dy = torch.randn(T, Cc, device=dev)
        w2 = torch.randn(Cc, J, device=dev) / math.sqrt(J)
        z = torch.randn(T, J, device=dev)
        b1 = torch.randn(J, device=dev)
        dz = torch.empty_like(z)
        dz2 = torch.empty_like(z)
        db = torch.zeros(J, device=dev)

        def fused():
            check(lib.hs_mlp_dgrad_gelu(ptr(dy), ptr(w2), ptr(z), ptr(b1), C.c_float(0.0), C.c_uint64(0), ptr(dz), T, Cc, J, 0,
                                        current_stream()))

        def pair():
            dh = dy @ w2
            check(lib.hs_bias_gelu_bwd(ptr(dh), ptr(z), ptr(b1), C.c_float(0.0), C.c_uint64(0), ptr(dz2), ptr(db), T, J,
                                       current_stream()))

        def gelu_fwd():
            check(lib.hs_bias_gelu_fwd(ptr(z), ptr(b1), C.c_float(0.0), C.c_uint64(0), ptr(dz2), T, J, current_stream()))

        def gelu_bwd():
            check(lib.hs_bias_gelu_bwd(ptr(dz), ptr(z), ptr(b1), C.c_float(0.0), C.c_uint64(0), ptr(dz2), ptr(db), T, J,
                                       current_stream()))

        print(f"T={T} J={J}: bias_gelu_fwd {timeit(gelu_fwd):.3f} ms, bias_gelu_bwd {timeit(gelu_bwd):.3f} ms", flush=True)
        fused()
        pair()
        torch.cuda.synchronize()
        err = ((dz - dz2).norm() / dz2.norm()).item()
        tf, tp = timeit(fused), timeit(pair)
        gb = (T * Cc + 2 * T * J) * 4 / 1e9
        print(f"T={T} C={Cc} J={J}: fused {tf:.3f} ms ({gb / tf * 1e3:.0f} GB/s algorithmic), library dgrad + gelu_bwd {tp:.3f} ms, "
              f"rel diff {err:.2e}", flush=True)


if __name__ == "__main__":
    main()
