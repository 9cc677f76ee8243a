#!/usr/bin/env python
"""How much of the training step is GPU-busy?  Sums the kernel durations of a few steps with torch.profiler and compares
with the CUDA-event step time (a large gap = launch/CPU-bound stretches)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from heal_swin_b200.factory import build_hp_model as build_product_model  # noqa: E402


def main():
    a = bench.parse_args()
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = True
    kw = bench.model_kwargs(a)
    torch.manual_seed(0)
    model = build_product_model(kw, None, dev).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
    x = torch.randn(a.batch, 3, kw["dim_in"], device=dev)
    t = torch.randint(0, kw["f_out"], (a.batch, kw["dim_in"]), device=dev)
    from heal_swin_b200 import ops

    loss_fn = ops.CrossEntropyLoss()  # (as bench.py)

    def step():
        opt.zero_grad(set_to_none=True)
        loss_fn(model(x), t).backward()
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        step()
    e1.record()
    torch.cuda.synchronize()
    print(f"step (CUDA events, no profiler): {e0.elapsed_time(e1) / 3:.1f} ms")
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(2):
            step()
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    busy = sum(e.device_time for e in ev) / 1e3 / 2
    print(f"GPU kernel time per step (sum of durations under the profiler): {busy:.1f} ms over {len(ev) // 2} launches")
    agg = {}
    for e in ev:
        k = e.name.replace("(anonymous namespace)::", "").replace("at::native::", "")[:110]
        t, n = agg.get(k, (0.0, 0))
        agg[k] = (t + e.device_time / 1e3 / 2, n + 1)
    for k, (v, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:32]:
        print(f"  {v:7.2f} ms  x{n // 2:<4d} {k}")


if __name__ == "__main__":
    main()
