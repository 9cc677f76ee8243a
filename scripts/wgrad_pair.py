"""Weight-gradient kernel at the tensor-bound shapes: error vs fp64 (row sample) and time; run with HEALSWIN_WGRAD_PAIR=0 / 1."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heal_swin_b200._lib import check, current_stream, lib, ptr  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    ok = True
    for label, T, N, K in (("s2 qkv", 98304, 1152, 384), ("s2 fc1", 98304, 1536, 384), ("s2 fc2", 98304, 384, 1536),
                           ("s2 proj", 98304, 384, 384), ("s3 fc1", 24576, 3072, 768), ("s3 qkv", 24576, 2304, 768),
                           ("s3 fc2", 24576, 768, 3072), ("s1 fc1", 393216, 768, 192), ("s1 fc2", 393216, 192, 768),
                           ("ragged", 5000, 512, 256)):
        dy = torch.randn(T, N, device=dev)
        x = torch.randn(T, K, device=dev)
        dw = torch.zeros(N, K, device=dev)
        check(lib.hs_linear_wgrad(ptr(dy), ptr(x), ptr(dw), None, T, N, K, 0, current_stream()))
        Ts = min(T, 8192)
        ref = (dy[:Ts].double().t() @ x[:Ts].double())
        dws = torch.zeros(N, K, device=dev)
        check(lib.hs_linear_wgrad(ptr(dy), ptr(x), ptr(dws), None, Ts, N, K, 0, current_stream()))
        e_small = float((dws.double() - ref).norm() / ref.norm())
        full = dy.t() @ x  # TF32 / fp32 library reference for the full contraction
        e_full = float((dw - full).norm() / full.norm())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            check(lib.hs_linear_wgrad(ptr(dy), ptr(x), ptr(dw), None, T, N, K, 0, current_stream()))
        e0.record()
        for _ in range(10):
            check(lib.hs_linear_wgrad(ptr(dy), ptr(x), ptr(dw), None, T, N, K, 0, current_stream()))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        good = e_small < 1e-3 and e_full < 2e-3
        ok = ok and good
        print(f"{label:8s} T={T:7d} N={N:5d} K={K:5d}: err(8192 rows vs fp64) {e_small:.2e}  err(full vs torch) {e_full:.2e}  "
              f"{ms:.3f} ms  {2.0 * T * N * K / ms / 1e9:.0f} TFLOP/s  {'OK' if good else 'FAIL'}", flush=True)
    print("wgrad_pair:", "ALL OK" if ok else "FAILURES")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
