"""One shape of the bf16x3 GEMM, a few launches: the target of an `ncu --set full` capture.
python scripts/gemm3_prof.py T N K [mode [precision]]"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gemm3_check import gemm3, split  # noqa: E402


def main():
    T, N, K = (int(v) for v in sys.argv[1:4])
    mode = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    prec = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    dev = torch.device("cuda:0")
    a = torch.randn(T, K, device=dev)
    w = torch.randn(N, K, device=dev) / math.sqrt(K)
    bias = torch.randn(N, device=dev)
    aux = torch.randn(T, N, device=dev) if mode in (1, 3) else None
    ws = split(w, prec=prec)
    d = torch.empty(T, N, device=dev)
    d2 = torch.empty(T, N, device=dev) if mode == 2 else None
    for _ in range(4):
        gemm3(a, ws, bias, aux, mode=mode, d=d, d2=d2, prec=prec)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
