#!/usr/bin/env python
"""One training step (forward + CE + backward) of every BASELINE.json configuration that fits one GPU, on the B200
kernels: config 1 (tiny, window 16 / head_dim 16: exact-fp32 CUDA-core attention path), config 2 (the benchmark model),
config 4 (depth head, C=128: fp32-class GEMMs and the bf16-operand mode), config 5 (flat SWIN-UNet 640x640,
window 8).  Prints ms per step (CUDA events, 3 steps after 2 warm-ups) and checks that outputs and gradients are finite."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heal_swin_b200.data_spec import DataSpec  # noqa: E402
from heal_swin_b200.models_torch import swin_hp_transformer as HP  # noqa: E402
from heal_swin_b200.models_torch import swin_transformer as FL  # noqa: E402

dev = torch.device("cuda:0")


def run(name, model, x, target_fn):
    model = model.to(dev).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)

    def step():
        opt.zero_grad(set_to_none=True)
        y = model(x)
        loss = target_fn(y)
        loss.backward()
        opt.step()
        return y, loss

    for _ in range(2):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        y, loss = step()
    e1.record()
    torch.cuda.synchronize()
    ok = bool(torch.isfinite(y).all()) and all(bool(torch.isfinite(p.grad).all()) for p in model.parameters() if p.grad is not None)
    npix = x.shape[0] * (x.shape[-1] if x.dim() == 3 else x.shape[-1] * x.shape[-2])
    ms = e0.elapsed_time(e1) / 3
    print(f"{name}: {ms:.1f} ms/step, {npix / ms * 1e3 / 1e6:.1f} Mpix/s, loss {float(loss.detach()):.4f}, finite={ok}, "
          f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    assert ok
    del model, opt
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()


ce = torch.nn.CrossEntropyLoss()
g = torch.Generator().manual_seed(0)

# config 1: HEAL-SWIN-tiny N_side=64, window 16, depths [2,2], C=48, batch 2
cfg = HP.SwinHPTransformerConfig(patch_size=4, window_size=16, shift_size=4, shift_strategy="nest_roll", rel_pos_bias="flat",
                                 embed_dim=48, depths=[2, 2], num_heads=[3, 6], drop_path_rate=0.0)
n = 12 * 64 * 64
x = torch.randn(2, 3, n, generator=g).to(dev)
t = torch.randint(0, 10, (2, n), generator=g).to(dev)
run("config 1 (tiny, ws16/d16, SIMT attention)", HP.SwinHPTransformerSys(cfg, DataSpec(n, 3, 10, 12)), x, lambda y: ce(y, t))

# config 2: the benchmark model, batch 8
n = 12 * 256 * 256
cfg = HP.SwinHPTransformerConfig(patch_size=4, window_size=64, shift_size=4, shift_strategy="nest_roll", rel_pos_bias="flat",
                                 embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], use_cos_attn=True,
                                 use_v2_norm_placement=True, drop_path_rate=0.0)
x = torch.randn(8, 3, n, generator=g).to(dev)
t = torch.randint(0, 10, (8, n), generator=g).to(dev)
run("config 2 (UNet N_side=256, C=96, 10 classes, B=8)", HP.SwinHPTransformerSys(cfg, DataSpec(n, 3, 10, 12)), x, lambda y: ce(y, t))

# config 2b: the paper-faithful variant: base_pix=8, ring_shift, all drop rates 0.1
n8 = 8 * 256 * 256
cfg = HP.SwinHPTransformerConfig(patch_size=4, window_size=64, shift_size=4, shift_strategy="ring_shift", rel_pos_bias="flat",
                                 embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], use_cos_attn=True,
                                 use_v2_norm_placement=True, drop_rate=0.1, attn_drop_rate=0.1, drop_path_rate=0.1)
x8 = torch.randn(8, 3, n8, generator=g).to(dev)
t8 = torch.randint(0, 10, (8, n8), generator=g).to(dev)
run("config 2b (base_pix=8, ring_shift, drop rates 0.1, B=8)", HP.SwinHPTransformerSys(cfg, DataSpec(n8, 3, 10, 8)), x8, lambda y: ce(y, t8))
del x8, t8

# config 4: depth-estimation head, C=128, heads [4,8,16,32], f_out=1 (fp32/TF32 here, not bf16)
cfg = HP.SwinHPTransformerConfig(patch_size=4, window_size=64, shift_size=4, shift_strategy="nest_roll", rel_pos_bias="flat",
                                 embed_dim=128, depths=[2, 2, 6, 2], num_heads=[4, 8, 16, 32], use_cos_attn=True,
                                 use_v2_norm_placement=True, drop_path_rate=0.0)
d = torch.randn(8, 1, n, generator=g).to(dev)
run("config 4 (depth head, C=128, f_out=1, B=8, fp32-class GEMMs)", HP.SwinHPTransformerSys(cfg, DataSpec(n, 3, 1, 12)), x,
    lambda y: (y - d).square().mean())
from heal_swin_b200 import ops  # noqa: E402

ops.set_gemm_precision("bf16")  # "bf16 operands, fp32 accumulate": the arithmetic BASELINE configs[3] names
run("config 4 (depth head, C=128, f_out=1, B=8, bf16 operands)", HP.SwinHPTransformerSys(cfg, DataSpec(n, 3, 1, 12)), x,
    lambda y: (y - d).square().mean())
ops.set_gemm_precision("fp32")
del x, t, d

# config 5: flat SWIN-UNet 640x640, patch 2, window 8, shift 2, C=96, batch 8
cfg = FL.SwinTransformerConfig(patch_size=2, window_size=8, shift_size=2, embed_dim=96, depths=[2, 2, 6, 2],
                               num_heads=[3, 6, 12, 24], drop_path_rate=0.0)
xf = torch.randn(8, 3, 640, 640, generator=g).to(dev)
tf = torch.randint(0, 10, (8, 640, 640), generator=g).to(dev)
run("config 5 (flat SWIN-UNet 640x640, window 8, B=8)", FL.SwinTransformerSys(cfg, DataSpec((640, 640), 3, 10, None)), xf,
    lambda y: ce(y, tf))
