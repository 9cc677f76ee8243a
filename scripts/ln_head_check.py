"""Timing of the fused decoder tail (LayerNorm + 1x1 output projection, forward and backward) against the unfused chain
(hs_layernorm + library GEMMs) at the bench size.  python scripts/ln_head_check.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heal_swin_b200 import ops  # noqa: E402
from scripts.mlp_check import timeit  # noqa: E402


def main():
    torch.backends.cuda.matmul.allow_tf32 = True
    dev = torch.device("cuda:0")
    B, P, Cc, K = 8, 786432, 96, 10
    x = torch.randn(B, P, Cc, device=dev).requires_grad_(True)
    norm = torch.nn.LayerNorm(Cc).to(dev)
    w = (torch.randn(K, Cc, device=dev) / 10).requires_grad_(True)
    gy = torch.randn(B, K, P, device=dev)

    def fused_fwd():
        return ops.ln_head(x, norm, w)

    def unfused_fwd():
        return ops.linear(ops.layer_norm(x, norm), w).permute(0, 2, 1).contiguous()

    def fb(f):
        def run():
            y = f()
            y.backward(gy)
            x.grad = w.grad = None
            norm.zero_grad(set_to_none=True)
        return run

    with torch.no_grad():
        tf, tu = timeit(fused_fwd, 10), timeit(unfused_fwd, 10)
    tfb, tub = timeit(fb(fused_fwd), 10), timeit(fb(unfused_fwd), 10)
    gb_f = B * P * (Cc + K) * 4 / 1e9
    gb_b = B * P * (2 * Cc + K) * 4 / 1e9
    print(f"forward fused {tf:.3f} ms ({gb_f / tf * 1e3:.0f} GB/s) vs "
          f"unfused {tu:.3f} ms | fwd+bwd fused {tfb:.3f} ms (bwd {tfb - tf:.3f} ms, {gb_b / (tfb - tf) * 1e3:.0f} GB/s) vs "
          f"unfused {tub:.3f} ms", flush=True)


if __name__ == "__main__":
    main()
