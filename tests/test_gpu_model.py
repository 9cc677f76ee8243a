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


TF32_GEMM_FWD_TOL = 2e-3


def test_forward_error_with_tf32_library_gemms_is_bounded():
    """bench.py runs the library GEMMs (cuBLAS) in TF32 like the reference's pinned torch 1.8 did by default.  TF32 GEMMs
    alone put the network output 1.1-1.4e-3 away from the fp32 oracle (measured, scripts/tf32_model_check.py), slightly
    outside the 1e-3 bound that holds with fp32 GEMMs (tests above); this pins the documented figure at 2e-3."""
    dev = torch.device("cuda:0")
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        for name in ("ring_cos_v2_ws64", "roll_v1_ws64"):
            kw, cfg, sd, gold = load_model_case(name)
            model = build_product_model(kw, sd, dev).eval()
            with torch.no_grad():
                y = model(torch.from_numpy(gold["x"]).to(dev))
            assert rel_err(y.cpu(), gold["y"]) < TF32_GEMM_FWD_TOL, name
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


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
