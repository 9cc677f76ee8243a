#!/usr/bin/env python
"""GPU development check: the tcgen05 attention kernels against the exact-fp32 CUDA-core kernels on
random inputs, for every shift strategy / flag combination.  Prints one line per case (relative L2
error) and, with --time, CUDA-event timings at the BASELINE stage-0 shape."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from heal_swin_b200 import ops  # noqa: E402
from heal_swin_b200.models_torch import hp_shifting as S  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def tables(strategy, nside, base_pix, ws, dev):
    N = base_pix * nside * nside
    if strategy == "none":
        return None, None
    if strategy == "nest_roll":
        sh = S.NestRollShift(ws // 4, N, ws)
    elif strategy == "nest_grid_shift":
        sh = S.NestGridShift(nside, base_pix, ws)
    else:
        sh = S.RingShift(nside, base_pix, ws, ws // 4)
    src, _, grp = sh.device_tables(dev)
    return src, grp


def run_case(B, nside, base_pix, H, strategy, cos, use_bias, dev, seed=0):
    ws, D = 64, 32
    C = H * D
    N = base_pix * nside * nside
    g = torch.Generator(device="cpu").manual_seed(seed)
    qkv = torch.randn(B, N, 3 * C, generator=g).to(dev)
    table = (torch.randn(225, H, generator=g) * 0.5).to(dev) if use_bias else None
    rel_index = None
    if use_bias:
        from heal_swin_b200 import hp_index
        rel_index = hp_index.rel_pos_index(ws).to(torch.int32).reshape(-1).contiguous().to(dev)
    ls = (torch.log(torch.tensor(10.0)) + 0.3 * torch.randn(H, 1, 1, generator=g)).to(dev) if cos else None
    src, grp = tables(strategy, nside, base_pix, ws, dev)
    outs = {}
    for mode in ("fp32", "tf32"):
        ops.set_attention_precision(mode)
        with torch.no_grad():
            outs[mode] = ops.window_attention_core(qkv, table, ls, src, grp, None, rel_index, D ** -0.5, H, ws, cos)
        torch.cuda.synchronize()
    ops.set_attention_precision("tf32")
    e = rel(outs["tf32"], outs["fp32"])
    mx = float((outs["tf32"] - outs["fp32"]).abs().max())
    return e, mx


def run_case_bwd(B, nside, base_pix, H, strategy, cos, use_bias, dev, seed=0):
    """Backward of the attention core: tcgen05 kernels vs the exact-fp32 CUDA-core kernels.
    Returns {name: relative L2 error} for dqkv, dtable, dlogit_scale."""
    ws, D = 64, 32
    C = H * D
    N = base_pix * nside * nside
    g = torch.Generator(device="cpu").manual_seed(seed + 17)
    qkv0 = torch.randn(B, N, 3 * C, generator=g).to(dev)
    wgt = torch.randn(B, N, C, generator=g).to(dev)
    table0 = (torch.randn(225, H, generator=g) * 0.5).to(dev) if use_bias else None
    rel_index = None
    if use_bias:
        from heal_swin_b200 import hp_index
        rel_index = hp_index.rel_pos_index(ws).to(torch.int32).reshape(-1).contiguous().to(dev)
    ls0 = (torch.log(torch.tensor(10.0)) + 0.3 * torch.randn(H, 1, 1, generator=g)).to(dev) if cos else None
    src, grp = tables(strategy, nside, base_pix, ws, dev)
    grads = {}
    for mode in ("fp32", "tf32"):
        ops.set_attention_precision(mode)
        qkv = qkv0.clone().requires_grad_(True)
        table = table0.clone().requires_grad_(True) if use_bias else None
        ls = ls0.clone().requires_grad_(True) if cos else None
        out = ops.window_attention_core(qkv, table, ls, src, grp, None, rel_index, D ** -0.5, H, ws, cos)
        (out * wgt).sum().backward()
        torch.cuda.synchronize()
        grads[mode] = {"dqkv": qkv.grad, "dtable": None if table is None else table.grad,
                       "dlogit_scale": None if ls is None else ls.grad}
    ops.set_attention_precision("tf32")
    return {k: rel(grads["tf32"][k], v) for k, v in grads["fp32"].items() if v is not None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--time", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    bad = 0
    cases = [
        (1, 4, 8, 1, "none", False, False),
        (1, 4, 8, 3, "none", False, True),
        (2, 8, 8, 3, "none", True, True),
        (1, 4, 12, 3, "nest_roll", True, True),
        (3, 4, 8, 2, "nest_roll", False, True),     # odd number of (batch x window) units
        (1, 8, 8, 3, "nest_grid_shift", True, True),
        (2, 8, 8, 6, "ring_shift", True, True),
        (1, 16, 8, 24, "ring_shift", False, False),
        (2, 32, 12, 3, "nest_roll", True, True),
    ]
    for c in cases:
        e, mx = run_case(*c, dev)
        ok = e < 2e-3
        bad += not ok
        print(f"B={c[0]} nside={c[1]} base_pix={c[2]} H={c[3]} {c[4]:16s} cos={int(c[5])} bias={int(c[6])}: "
              f"rel {e:.3e} max_abs {mx:.3e} {'ok' if ok else 'FAIL'}", flush=True)
    for c in cases:
        errs = run_case_bwd(*c, dev)
        ok = all(e < (1e-2 if k == "dlogit_scale" else 3e-3) for k, e in errs.items())
        bad += not ok
        print(f"BWD B={c[0]} nside={c[1]} base_pix={c[2]} H={c[3]} {c[4]:16s} cos={int(c[5])} bias={int(c[6])}: "
              + " ".join(f"{k} {e:.3e}" for k, e in errs.items()) + (" ok" if ok else " FAIL"), flush=True)
    if a.time:
        B, N, H = 8, 12 * 256 * 256 // 4, 3
        C = 96
        qkv = torch.randn(B, N, 3 * C, device=dev)
        dout = torch.randn(B, N, C, device=dev)
        for mode in ("tf32", "fp32"):
            ops.set_attention_precision(mode)
            q = qkv.clone().requires_grad_(True)
            o = ops.window_attention_core(q, None, None, None, None, None, None, 32 ** -0.5, H, 64, False)
            for _ in range(2):
                torch.autograd.grad(o, q, dout, retain_graph=True)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            for _ in range(5):
                torch.autograd.grad(o, q, dout, retain_graph=True)
            ev[1].record()
            torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[1]) / 5
            gb = B * N * 7 * C * 4 / 1e9
            print(f"stage-0 bwd ({mode}): {ms:.3f} ms  -> {gb / ms * 1e3:.0f} GB/s algorithmic (q,k,v,dO in; dq,dk,dv out)", flush=True)
            del q, o
        for mode in ("tf32", "fp32"):
            ops.set_attention_precision(mode)
            for _ in range(3):
                ops.window_attention_core(qkv, None, None, None, None, None, None, 32 ** -0.5, H, 64, False)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            for _ in range(10):
                ops.window_attention_core(qkv, None, None, None, None, None, None, 32 ** -0.5, H, 64, False)
            ev[1].record()
            torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[1]) / 10
            gb = B * N * 4 * C * 4 / 1e9
            print(f"stage-0 fwd ({mode}): {ms:.3f} ms  -> {gb / ms * 1e3:.0f} GB/s algorithmic", flush=True)
        ops.set_attention_precision("tf32")
    print("tc_check", "FAILED" if bad else "passed")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
