"""One shape of the LayerNorm-epilogue GEMM (hs_gemm3_ln, training form: pre-norm tensor + statistics written, shortcut
added) or, with G = 0, of the LayerNorm-prologue GEMM (hs_gemm3_lnin): the target of an `ncu --set full` capture.
python scripts/gemm3_ln_prof.py T N K G [shortcut 0/1]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heal_swin_b200 import ops  # noqa: E402


def main():
    T, N, K, G = (int(v) for v in sys.argv[1:5])
    res = len(sys.argv) > 5 and sys.argv[5] == "1"
    dev = torch.device("cuda:0")
    x = torch.randn(T, K, device=dev, requires_grad=True)
    lin = torch.nn.Linear(K, N, bias=G > 0).to(dev)
    if G > 0:
        norm = torch.nn.LayerNorm(G).to(dev)
        sc = torch.randn(T, N, device=dev) if res else None
        for _ in range(4):
            ops.linear_ln(x, lin.weight, lin.bias, norm, residual=sc)
    else:
        norm = torch.nn.LayerNorm(K).to(dev)
        for _ in range(4):
            ops.ln_linear(x, norm, lin.weight)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
