#!/usr/bin/env python
"""GPU timing of the LayerNorm-in-the-epilogue GEMM (hs_gemm3_ln) against the two launches it replaces (hs_gemm3 +
hs_layernorm_fwd with the residual add) at the BASELINE shapes: block tails of stages 0-1 (proj and fc2) and PatchExpand."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heal_swin_b200 import ops  # noqa: E402

HBM = 6550.7  # GB/s, MEASURED_PEAKS.json


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


def main():
    dev = torch.device("cuda:0")
    cases = [("proj s0", 8 * 196608, 96, 96, 96, True), ("fc2 s0", 8 * 196608, 96, 384, 96, True),
             ("proj s1", 8 * 49152, 192, 192, 192, True), ("fc2 s1", 8 * 49152, 192, 768, 192, True),
             ("expand 1->0", 8 * 49152, 384, 192, 96, False), ("expand 2->1", 8 * 12288, 768, 384, 192, False)]
    for name, T, N, K, G, res in cases:
        x = torch.randn(T, K, device=dev)
        sc = torch.randn(T, N, device=dev) if res else None
        lin = torch.nn.Linear(K, N).to(dev)
        norm = torch.nn.LayerNorm(G).to(dev)
        x.requires_grad_(True)  # training form: the pre-norm tensor and the statistics are written as well
        assert ops.linear_ln_supported(x, lin.weight, norm)

        def unfused():
            y = ops.linear(x, lin.weight)
            return ops.layer_norm(y.view(T * (N // G), G), norm, residual=None if sc is None else sc.view(-1, G),
                                  pre_bias=lin.bias if G == N else None)

        t_f = timeit(lambda: ops.linear_ln(x, lin.weight, lin.bias, norm, residual=sc))
        t_u = timeit(unfused)
        with torch.no_grad():
            t_e = timeit(lambda: ops.linear_ln(x, lin.weight, lin.bias, norm, residual=sc))
        gb = T * (K + 2 * N + (N if res else 0)) * 4 / 1e9
        gbe = T * (K + N + (N if res else 0)) * 4 / 1e9
        print(f"{name:12s} T={T} N={N} K={K} G={G}: fused {t_f:.3f} ms ({gb / t_f * 1e3:.0f} GB/s, {gb / t_f * 1e3 / HBM:.2f}) "
              f"| gemm + LN {t_u:.3f} ms | no-grad fused {t_e:.3f} ms ({gbe / t_e * 1e3 / HBM:.2f})", flush=True)
        del x, sc


if __name__ == "__main__":
    main()
