"""One shape of the weight-gradient kernel, a few launches: the target of an ncu capture.
python scripts/wgrad_prof.py T N K"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heal_swin_b200._lib import check, current_stream, lib, ptr  # noqa: E402

T, N, K = (int(v) for v in sys.argv[1:4])
dev = torch.device("cuda:0")
dy = torch.randn(T, N, device=dev)
x = torch.randn(T, K, device=dev)
dw = torch.zeros(N, K, device=dev)
for _ in range(4):
    check(lib.hs_linear_wgrad(ptr(dy), ptr(x), ptr(dw), None, T, N, K, 0, current_stream()))
torch.cuda.synchronize()
