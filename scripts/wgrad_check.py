#!/usr/bin/env python
"""Library-GEMM timing at the BASELINE shapes: forward (y = x W^T), dgrad (dx = dy W), wgrad (dW = dy^T x) in the two
transposition forms torch offers, TF32.  Prints ms and the HBM-roofline ms of each (bytes / 6.55 TB/s)."""
import torch

torch.backends.cuda.matmul.allow_tf32 = True
dev = torch.device("cuda:0")


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for stage, (M, C) in enumerate([(8 * 196608, 96), (8 * 49152, 192), (8 * 12288, 384), (8 * 3072, 768)]):
    for name, N, K in (("qkv", 3 * C, C), ("proj", C, C), ("fc1", 4 * C, C), ("fc2", C, 4 * C)):
        x = torch.randn(M, K, device=dev)
        dy = torch.randn(M, N, device=dev)
        w = torch.randn(N, K, device=dev)
        roof = lambda nbytes: nbytes / 6.55e12 * 1e3  # noqa: E731
        t_f = timeit(lambda: torch.nn.functional.linear(x, w))
        t_d = timeit(lambda: dy @ w)
        t_w1 = timeit(lambda: dy.t() @ x)
        t_w2 = timeit(lambda: (x.t() @ dy))
        b_f = (M * K + M * N) * 4
        print(f"stage {stage} {name:4s} M={M} N={N} K={K}: fwd {t_f:.3f} ms (roof {roof(b_f):.3f}) | dgrad {t_d:.3f} (roof {roof(b_f):.3f}) | "
              f"wgrad dy^T@x {t_w1:.3f}, x^T@dy {t_w2:.3f} (roof {roof(b_f):.3f})", flush=True)
        del x, dy, w

print("--- hand-written wgrad kernel (hs_linear_wgrad)")
import os, sys  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heal_swin_b200._lib import check, current_stream, lib, ptr  # noqa: E402

for stage, (M, C) in enumerate([(8 * 196608, 96), (8 * 49152, 192), (8 * 12288, 384), (8 * 3072, 768)]):
    for name, N, K in (("qkv", 3 * C, C), ("proj", C, C), ("fc1", 4 * C, C), ("fc2", C, 4 * C)):
        if not lib.hs_linear_wgrad_supported(M, N, K):
            print(f"stage {stage} {name}: not covered")
            continue
        x = torch.randn(M, K, device=dev)
        dy = torch.randn(M, N, device=dev)
        dw = torch.zeros(N, K, device=dev)
        t = timeit(lambda: check(lib.hs_linear_wgrad(ptr(dy), ptr(x), ptr(dw), None, M, N, K, 0, current_stream())))
        print(f"stage {stage} {name:4s} M={M} N={N} K={K}: custom wgrad {t:.3f} ms (roof {(M * (N + K)) * 4 / 6.55e12 * 1e3:.3f})", flush=True)
        del x, dy, dw
