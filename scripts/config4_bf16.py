#!/usr/bin/env python
"""BASELINE configs[3] (depth head, C = 128, heads [4, 8, 16, 32], f_out = 1, batch 8) in the bf16-operand mode: ms per
training step and the per-kernel composition (torch.profiler).  `python scripts/config4_bf16.py [fp32]`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heal_swin_b200 import ops  # noqa: E402
from heal_swin_b200.data_spec import DataSpec  # noqa: E402
from heal_swin_b200.models_torch import swin_hp_transformer as HP  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
    ops.set_gemm_precision(mode)
    n = 12 * 256 * 256
    cfg = HP.SwinHPTransformerConfig(patch_size=4, window_size=64, shift_size=4, shift_strategy="nest_roll", rel_pos_bias="flat",
                                     embed_dim=128, depths=[2, 2, 6, 2], num_heads=[4, 8, 16, 32], use_cos_attn=True,
                                     use_v2_norm_placement=True, drop_path_rate=0.0)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(8, 3, n, generator=g).to(dev)
    d = torch.randn(8, 1, n, generator=g).to(dev)
    model = HP.SwinHPTransformerSys(cfg, DataSpec(n, 3, 1, 12)).to(dev).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)

    def step():
        opt.zero_grad(set_to_none=True)
        (model(x) - d).square().mean().backward()
        opt.step()

    for _ in range(2):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        step()
    e1.record()
    torch.cuda.synchronize()
    print(f"config 4, {mode} GEMMs (ridge {os.environ.get('HEALSWIN_GEMM3_BF16_RIDGE', 'default')}): "
          f"{e0.elapsed_time(e1) / 3:.1f} ms/step", flush=True)
    if "--profile" in sys.argv:
        from torch.profiler import ProfilerActivity, profile

        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step()
            torch.cuda.synchronize()
        agg = {}
        for e in prof.events():
            if e.device_type == torch.autograd.DeviceType.CUDA:
                k = e.name.replace("(anonymous namespace)::", "")[:60]
                t, c = agg.get(k, (0.0, 0))
                agg[k] = (t + e.device_time / 1e3, c + 1)
        for k, (v, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:14]:
            print(f"  {v:7.2f} ms  x{c:<4d} {k}")


if __name__ == "__main__":
    main()
