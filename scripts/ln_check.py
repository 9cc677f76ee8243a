#!/usr/bin/env python
"""GPU timing of the LayerNorm kernels at the BASELINE stage shapes next to torch's own kernels."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heal_swin_b200 import ops  # noqa: E402


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
    for rows, C in [(8 * 196608, 96), (8 * 49152, 192), (8 * 12288, 384), (8 * 3072, 768), (8 * 49152, 384), (8 * 786432, 96)]:
        x = torch.randn(rows, C, device=dev, requires_grad=True)
        res = torch.randn(rows, C, device=dev)
        norm = torch.nn.LayerNorm(C).to(dev)
        dy = torch.randn(rows, C, device=dev)
        gb = rows * C * 4 / 1e9
        t_f = timeit(lambda: ops.layer_norm(x, norm, residual=res))
        t_ft = timeit(lambda: res + F.layer_norm(x, (C,), norm.weight, norm.bias, norm.eps))
        y = ops.layer_norm(x, norm, residual=res)
        t_b = timeit(lambda: torch.autograd.grad(y, [x, norm.weight, norm.bias], dy, retain_graph=True))
        yt = F.layer_norm(x, (C,), norm.weight, norm.bias, norm.eps)
        t_bt = timeit(lambda: torch.autograd.grad(yt, [x, norm.weight, norm.bias], dy, retain_graph=True))
        pb = torch.randn(C, device=dev, requires_grad=True)
        y2 = ops.layer_norm(x, norm, residual=res, pre_bias=pb)
        t_b2 = timeit(lambda: torch.autograd.grad(y2, [x, norm.weight, norm.bias, pb], dy, retain_graph=True))
        y3 = ops.layer_norm(x, norm, residual=res, pre_bias=pb, in_drop=0.1, seed=5)
        t_f3 = timeit(lambda: ops.layer_norm(x, norm, residual=res, pre_bias=pb, in_drop=0.1, seed=5))
        t_b3 = timeit(lambda: torch.autograd.grad(y3, [x, norm.weight, norm.bias, pb], dy, retain_graph=True))
        print(f"   with pre-bias: bwd {t_b2:.3f} ms; with pre-bias + dropout 0.1: fwd {t_f3:.3f} ms bwd {t_b3:.3f} ms")
        del y2, y3
        print(f"rows={rows} C={C}: fused fwd {t_f:.3f} ms ({3 * gb / t_f * 1e3:.0f} GB/s; torch LN+add {t_ft:.3f} ms)  "
              f"bwd {t_b:.3f} ms ({3 * gb / t_b * 1e3:.0f} GB/s; torch {t_bt:.3f} ms)", flush=True)
        del x, res, dy, y, yt


if __name__ == "__main__":
    main()
