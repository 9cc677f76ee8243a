"""End-to-end sanity of the whole autograd chain on the B200 kernels: a small HEAL-SWIN-UNet overfits one fixed batch
(cross-entropy must drop well below chance) in the configuration bench.py runs (TF32 library GEMMs -> hand-written weight
gradient kernel active) and in the reference's default training configuration (all drop rates 0.1, stochastic depth)."""
import math

import pytest
import torch

from tests.util import build_product_model

pytestmark = pytest.mark.gpu

KW = dict(patch_size=4, window_size=64, shift_size=4, shift_strategy="ring_shift", rel_pos_bias="flat", embed_dim=96,
          depths=[2, 2], num_heads=[3, 6], use_cos_attn=True, use_v2_norm_placement=True, dim_in=8 * 32 * 32, f_in=3,
          f_out=5, base_pix=8)


def _overfit(model, steps, dev, tf32):
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    try:
        g = torch.Generator().manual_seed(0)
        x = torch.randn(2, KW["f_in"], KW["dim_in"], generator=g).to(dev)
        # a learnable per-pixel target: which quadrant the first two input channels fall in (classes 0..3 of 5)
        t = ((x[:, 0] > 0).long() + 2 * (x[:, 1] > 0).long())
        opt = torch.optim.Adam(model.parameters(), lr=2e-3)
        loss_fn = torch.nn.CrossEntropyLoss()
        losses = []
        for _ in range(steps):
            opt.zero_grad(set_to_none=True)
            loss = loss_fn(model(x), t)
            loss.backward()
            opt.step()
            losses.append(float(loss.detach()))
        return losses
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def test_overfits_one_batch_with_tf32_gemms_and_custom_wgrad():
    from heal_swin_b200 import ops

    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = build_product_model(KW, None, dev).train()
    ops.STATS.reset()
    losses = _overfit(model, 60, dev, tf32=True)
    assert abs(losses[0] - math.log(KW["f_out"])) < 0.7, losses[0]
    assert losses[-1] < 0.5 * losses[0], losses[::10]
    assert all(math.isfinite(v) for v in losses)
    assert ops.STATS.launches > 0


def test_trains_with_the_reference_default_drop_rates():
    from heal_swin_b200.data_spec import DataSpec
    from heal_swin_b200.models_torch import swin_hp_transformer as M

    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    cfgkw = {k: v for k, v in KW.items() if k not in ("dim_in", "f_in", "f_out", "base_pix")}
    cfg = M.SwinHPTransformerConfig(**cfgkw, drop_rate=0.1, attn_drop_rate=0.1, drop_path_rate=0.1)
    model = M.SwinHPTransformerSys(cfg, data_spec=DataSpec(KW["dim_in"], KW["f_in"], KW["f_out"], KW["base_pix"])).to(dev).train()
    losses = _overfit(model, 60, dev, tf32=True)
    assert all(math.isfinite(v) for v in losses)
    assert min(losses[-10:]) < 0.8 * losses[0], losses[::10]
