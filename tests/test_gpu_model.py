"""GPU parity: whole HEAL-SWIN-UNet forward/backward through the drop-in modules against the
reference-generated fixtures and the CPU oracle.

Tolerances (relative L2):  forward 1e-3 (BASELINE.json north_star), gradients 5e-3.
"""
import numpy as np
import pytest
import torch

from oracle import hp_oracle as O
from oracle.make_golden import MODEL_CASES, GRAD_KEYS
from tests.util import build_product_model, load_model_case, rel_err

pytestmark = pytest.mark.gpu

FWD_TOL = 1e-3
GRAD_TOL = 5e-3
LOGIT_SCALE_GRAD_TOL = 2e-2


@pytest.mark.parametrize("name", list(MODEL_CASES))
def test_model_forward_backward_vs_reference_fixture(name):
    dev = torch.device("cuda:0")
    kw, cfg, sd, gold = load_model_case(name)
    model = build_product_model(kw, sd, dev)
    model.train()
    x = torch.from_numpy(gold["x"]).to(dev)
    y = model(x)
    assert y.shape == gold["y"].shape
    err = rel_err(y.detach().cpu(), gold["y"])
    assert err < FWD_TOL, err
    (y * torch.from_numpy(gold["wgt"]).to(dev)).sum().backward()
    params = dict(model.named_parameters())
    checked = 0
    for k in GRAD_KEYS:
        if "grad:" + k in gold.files:
            e = rel_err(params[k].grad.cpu(), gold["grad:" + k])
            # logit_scale: one scalar per head, a heavily cancelling sum over all windows -> TF32 noise is amplified
            assert e < (LOGIT_SCALE_GRAD_TOL if k.endswith("logit_scale") else GRAD_TOL), (k, e)
            checked += 1
    assert checked >= 6


def test_model_vs_oracle_on_fresh_inputs():
    dev = torch.device("cuda:0")
    kw, cfg, sd, gold = load_model_case("ring_cos_v2_ws64")
    model = build_product_model(kw, sd, dev).eval()
    x = torch.randn(2, kw["f_in"], kw["dim_in"], generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        want = O.hp_unet_forward(x, sd, cfg)
        got = model(x.to(dev)).cpu()
    assert rel_err(got, want) < FWD_TOL


def test_outputs_are_plain_writable_tensors():
    # the Lightning depth wrapper mutates the model output in place
    dev = torch.device("cuda:0")
    kw, cfg, sd, gold = load_model_case("roll_v1_ws16")
    model = build_product_model(kw, sd, dev).eval()
    with torch.no_grad():
        y = model(torch.from_numpy(gold["x"]).to(dev))
    y[:, 0] = y[:, 0].exp()
    assert torch.isfinite(y).all()


def test_default_mode_is_the_hand_written_gemm_path():
    """There is ONE product configuration: what these tests, smoke() and bench.py run.  Every dense linear of the block
    goes through hs_gemm3 (bf16x3 tensor-core GEMM), the weight gradients through hs_linear_wgrad, independent of torch's
    global allow_tf32 switch."""
    from heal_swin_b200 import ops

    assert ops.get_gemm_mode() == "bf16x3" and ops.get_attention_precision() == "tf32"
    dev = torch.device("cuda:0")
    kw, cfg, sd, gold = load_model_case("ring_cos_v2_ws64")
    model = build_product_model(kw, sd, dev).train()
    x = torch.from_numpy(gold["x"]).to(dev)
    outs = []
    for tf32 in (False, True):
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        try:
            ops.STATS.reset()
            y = model(x)
            y.sum().backward()
            outs.append(y.detach().clone())
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev
        n = ops.STATS.by_name
        blocks = sum(cfg.depths) * 2 - cfg.depths[-1]  # encoder + decoder blocks
        assert n.get("gemm3", 0) >= 7 * blocks, n          # qkv, proj, fc1(+GELU), fc2 forward + 3 input gradients per block
        assert n.get("window_attn_fwd", 0) == blocks and n.get("window_attn_bwd", 0) == blocks, n
    assert torch.equal(outs[0], outs[1])


DEEP_KW = dict(patch_size=4, window_size=64, shift_size=4, shift_strategy="nest_roll", rel_pos_bias="flat", embed_dim=96,
               depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], use_cos_attn=True, use_v2_norm_placement=True,
               dim_in=12 * 64 * 64, f_in=3, f_out=10, base_pix=12)


def test_bench_architecture_at_reduced_nside_vs_oracle():
    """The architecture bench.py runs (BASELINE configs[1]: C=96, depths [2,2,6,2], heads [3,6,12,24], window 64, cos
    attention, v2 norm placement, nest_roll, 10 classes) at N_side=64 instead of 256: all 22 blocks, all four stage widths
    (C = 96 ... 768, i.e. every GEMM shape class of the bench), forward within 1e-3 and gradients within 5e-3 of the CPU
    oracle's autograd."""
    dev = torch.device("cuda:0")
    cfg = O.HPConfig(**DEEP_KW)
    sd = O.synth_state_dict(cfg, seed=7)
    x = torch.randn(2, 3, DEEP_KW["dim_in"], generator=torch.Generator().manual_seed(8))
    sd_g = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    want = O.hp_unet_forward(x, sd_g, cfg)
    wgt = torch.randn(want.shape, generator=torch.Generator().manual_seed(9))
    (want * wgt).sum().backward()
    model = build_product_model(DEEP_KW, sd, dev).train()
    got = model(x.to(dev))
    err = rel_err(got.detach().cpu(), want.detach())
    assert err < FWD_TOL, err
    (got * wgt.to(dev)).sum().backward()
    params = dict(model.named_parameters())
    for k in ("layers.0.blocks.1.attn.qkv.weight", "layers.2.blocks.3.mlp.fc1.weight", "layers.3.blocks.0.attn.proj.weight",
              "layers.1.downsample.reduction.weight", "decoder.layers_up.1.blocks.0.mlp.fc2.weight",
              "decoder.concat_back_dim.2.weight", "decoder.up.expand.weight", "patch_embed.proj.weight",
              "layers.2.blocks.5.norm2.weight", "decoder.layers_up.3.blocks.1.attn.relative_position_bias_table"):
        e = rel_err(params[k].grad.cpu(), sd_g[k].grad)
        assert e < GRAD_TOL, (k, e)


def test_reference_own_test_config_vs_oracle():
    """The model of the reference's own pytest config (heal_swin/testing/swin_hp_test_run_config.py:29-55: N_side=32,
    base_pix=8, window_size=4, patch_size=4, depths=(2, 1), num_heads=(1, 1), embed_dim=2 -> head_dim 2, C=2): every kernel
    takes its generic path (CUDA-core attention, scalar LayerNorm).  Forward and a gradient vs the CPU oracle."""
    dev = torch.device("cuda:0")
    kw = dict(patch_size=4, window_size=4, shift_size=2, shift_strategy="nest_roll", rel_pos_bias=None, embed_dim=2,
              depths=[2, 1], num_heads=[1, 1], dim_in=8 * 32 * 32, f_in=3, f_out=10, base_pix=8)
    cfg = O.HPConfig(**kw)
    sd = O.synth_state_dict(cfg, seed=21)
    x = torch.randn(2, 3, kw["dim_in"], generator=torch.Generator().manual_seed(4))
    sd_g = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    want = O.hp_unet_forward(x, sd_g, cfg)
    wgt = torch.randn(want.shape, generator=torch.Generator().manual_seed(5))
    (want * wgt).sum().backward()
    model = build_product_model(kw, sd, dev).train()
    got = model(x.to(dev))
    assert rel_err(got.detach().cpu(), want.detach()) < FWD_TOL
    (got * wgt.to(dev)).sum().backward()
    params = dict(model.named_parameters())
    # With C = 2 a LayerNorm output is (+1, -1) whatever its input, so every gradient that has to pass BACK through a
    # LayerNorm is pure cancellation (1e-7 of the upstream gradient, see test_gpu_layernorm.py) and not comparable;
    # the parameters after the last normalisation are.
    for k in ("decoder.output.weight", "decoder.up.norm.weight", "decoder.up.norm.bias"):
        assert rel_err(params[k].grad.cpu(), sd_g[k].grad) < GRAD_TOL, k
    assert all(torch.isfinite(p.grad).all() for p in params.values() if p.grad is not None)
