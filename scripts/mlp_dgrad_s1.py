#!/usr/bin/env python
"""Stage-1 MLP input gradient through the activation, dz = (dy W2) * GELU'(z + b1): the TF32 kernel hs_mlp_dgrad_gelu
against hs_gemm3 mode 3 in its three precisions (which one should ops._MlpFn.backward pick at C = 192?)."""
import ctypes as C
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heal_swin_b200 import _lib, ops  # noqa: E402
from heal_swin_b200._lib import check, current_stream, lib, ptr  # noqa: E402
from scripts.mlp_check import timeit  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    for T, Cc in [(8 * 196608, 96), (8 * 49152, 192)]:
        J = 4 * Cc
        dy = torch.randn(T, Cc, device=dev)
        w2 = torch.nn.Parameter(torch.randn(Cc, J, device=dev) / math.sqrt(J))
        z = torch.randn(T, J, device=dev)
        b1 = torch.randn(J, device=dev)
        dz = torch.empty_like(z)
        t0 = timeit(lambda: check(lib.hs_mlp_dgrad_gelu(ptr(dy), ptr(w2), ptr(z), ptr(b1), C.c_float(0.0), C.c_uint64(0), ptr(dz),
                                                        T, Cc, J, 0, current_stream())))
        gb = (T * Cc + 2 * T * J) * 4 / 1e9
        line = f"T={T} C={Cc}: hs_mlp_dgrad_gelu {t0:.3f} ms ({gb / t0 * 1e3 / 6550.7:.2f})"
        for name, prec in [("bf16x3", _lib.PREC_BF16X3), ("tf32", _lib.PREC_TF32)]:
            ws = ops.split_weight(w2, transposed=True, prec=prec)
            t1 = timeit(lambda: ops._gemm3(dy, ws, J, b1, z, _lib.GEMM_GELU_GRAD, prec=prec))
            line += f" | gemm3 mode 3 {name} {t1:.3f} ms ({gb / t1 * 1e3 / 6550.7:.2f})"
        print(line, flush=True)


if __name__ == "__main__":
    main()
