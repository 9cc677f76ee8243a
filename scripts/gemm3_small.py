"""Small / ragged token counts through the tensor-bound (ss, CTA-pair) regime of hs_gemm3, every mode, against fp64."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gemm3_check import gemm3, rel, split  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    ok = True
    for T in (256, 384, 300, 129, 640, 1000, 2048 + 17):
        for N, K in ((1152, 384), (384, 1536), (384, 384), (1536, 768), (200, 512)):
            a = torch.randn(T, K, device=dev)
            w = torch.randn(N, K, device=dev) / math.sqrt(K)
            bias = torch.randn(N, device=dev)
            aux = torch.randn(T, N, device=dev)
            ws = split(w)
            ref = a.double() @ w.double().t()
            errs = []
            for prec in (0, 1):
                wsp = split(w, prec=prec) if prec else ws
                cs = torch.zeros(K, device=dev)
                d0 = gemm3(a, wsp, bias, prec=prec, colsum=cs)
                tol = 2e-5 if prec == 0 else 2e-3
                errs.append(rel(d0, ref + bias.double()))
                errs.append(rel(cs, a.double().sum(0)) * (tol / 1e-4))
                d1 = gemm3(a, wsp, None, aux, mode=1, prec=prec)
                errs.append(rel(d1, ref + aux.double()))
                if prec == 0:
                    z, h = gemm3(a, wsp, bias, mode=2)
                    errs.append(rel(z, ref))
                    errs.append(rel(h, torch.nn.functional.gelu(ref + bias.double())))
                    g = gemm3(a, wsp, bias, aux, mode=3)
                    u = (aux.double() + bias.double()).requires_grad_(True)
                    torch.nn.functional.gelu(u).sum().backward()
                    errs.append(rel(g, ref * u.grad))
                good = all(e < tol for e in errs[-(6 if prec == 0 else 3):])
                ok = ok and good
            print(f"T={T:5d} N={N:5d} K={K:5d}: " + " ".join(f"{e:.1e}" for e in errs) + ("  OK" if good else "  FAIL"), flush=True)
    print("gemm3_small:", "ALL OK" if ok else "FAILURES")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
