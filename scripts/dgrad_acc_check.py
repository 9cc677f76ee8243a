"""Timing of the residual-folding input-gradient GEMM (hs_linear_dgrad_acc, cuBLASLt out-of-place C/D) against the two
alternatives: torch.addmm (copies c, then an in-place GEMM) and mm followed by add.  python scripts/dgrad_acc_check.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heal_swin_b200._lib import check, current_stream, lib, ptr  # noqa: E402
from scripts.mlp_check import timeit  # noqa: E402


def main():
    torch.backends.cuda.matmul.allow_tf32 = True
    dev = torch.device("cuda:0")
    ws = torch.empty(32 << 20, dtype=torch.uint8, device=dev)
    for T, N, K in [(8 * 196608, 288, 96), (8 * 196608, 384, 96), (8 * 49152, 576, 192), (8 * 49152, 768, 192),
                    (8 * 12288, 1152, 384), (8 * 12288, 1536, 384)]:
        dy = torch.randn(T, N, device=dev)
        w = torch.randn(N, K, device=dev) / N ** 0.5
        c = torch.randn(T, K, device=dev)
        dx = torch.empty_like(c)

        def lt():
            check(lib.hs_linear_dgrad_acc(ptr(dy), ptr(w), ptr(c), ptr(dx), T, N, K, ptr(ws), ws.numel(), current_stream()))

        lt()
        ref = torch.addmm(c, dy, w)
        err = ((dx - ref).norm() / ref.norm()).item()
        print(f"T={T} N={N} K={K}: dgrad_acc {timeit(lt):.3f} ms | addmm {timeit(lambda: torch.addmm(c, dy, w)):.3f} ms | "
              f"mm + add {timeit(lambda: (dy @ w) + c):.3f} ms | mm {timeit(lambda: dy @ w):.3f} ms | diff {err:.1e}", flush=True)


if __name__ == "__main__":
    main()
