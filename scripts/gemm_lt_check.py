"""Library GEMM selection: torch's default (cuBLAS) against the cuBLASLt heuristic pick with a 32 MB workspace
(hs_linear_fwd / hs_linear_dgrad_acc) for every dense-linear shape of the bench network.  python scripts/gemm_lt_check.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heal_swin_b200._lib import check, current_stream, lib, ptr  # noqa: E402
from scripts.mlp_check import timeit  # noqa: E402


def shapes():
    out = []
    T0 = 8 * 196608
    for s in range(4):
        T, Cc = T0 // 4 ** s, 96 * 2 ** s
        out += [("qkv", T, 3 * Cc, Cc, True), ("proj", T, Cc, Cc, False), ("fc1", T, 4 * Cc, Cc, False),
                ("fc2", T, Cc, 4 * Cc, False)]
        if s < 3:
            out += [("merge", T // 4, 2 * Cc, 4 * Cc, False), ("expand", T // 4, 4 * Cc, 2 * Cc, False),
                    ("concat_back", T, Cc, 2 * Cc, True)]
    out += [("final_expand", T0, 384, 96, False), ("head", 4 * T0, 10, 96, False), ("embed", T0, 96, 12, True)]
    return out


def main():
    torch.backends.cuda.matmul.allow_tf32 = True
    dev = torch.device("cuda:0")
    ws = torch.empty(32 << 20, dtype=torch.uint8, device=dev)
    tot = {"fwd_torch": 0.0, "fwd_lt": 0.0, "dgrad_torch": 0.0, "dgrad_lt": 0.0}
    for name, T, N, K, has_b in shapes():
        x = torch.randn(T, K, device=dev)
        w = torch.randn(N, K, device=dev) / K ** 0.5
        b = torch.randn(N, device=dev) if has_b else None
        y = torch.empty(T, N, device=dev)
        dx = torch.empty(T, K, device=dev)

        def fwd_lt():
            check(lib.hs_linear_fwd(ptr(x), ptr(w), ptr(b), ptr(y), T, N, K, ptr(ws), ws.numel(), current_stream()))

        def dgrad_lt():
            check(lib.hs_linear_dgrad_acc(ptr(y), ptr(w), None, ptr(dx), T, N, K, ptr(ws), ws.numel(), current_stream()))

        try:
            fwd_lt()
            ref = torch.nn.functional.linear(x, w, b)
            e1 = ((y - ref).norm() / ref.norm()).item()
            t = [timeit(lambda: torch.nn.functional.linear(x, w, b), 10), timeit(fwd_lt, 10)]
            y.copy_(ref)
            dgrad_lt()
            ref2 = y @ w
            e2 = ((dx - ref2).norm() / ref2.norm()).item()
            t += [timeit(lambda: y @ w, 10), timeit(dgrad_lt, 10)]
        except Exception as e:  # noqa: BLE001
            print(f"{name:13s} T={T} N={N} K={K}: FAILED {e}", flush=True)
            continue
        for k, v in zip(tot, t):
            tot[k] += v
        print(f"{name:13s} T={T:8d} N={N:5d} K={K:5d} | fwd torch {t[0]:.3f} lt {t[1]:.3f} ({e1:.0e}) | dgrad torch {t[2]:.3f} "
              f"lt {t[3]:.3f} ({e2:.0e})", flush=True)
        del x, w, y, dx
    print("sum over shapes (ms):", {k: round(v, 3) for k, v in tot.items()})


if __name__ == "__main__":
    main()
