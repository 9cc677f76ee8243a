"""BASELINE configs[3]: the depth-estimation head configuration (embed_dim 128, heads [4, 8, 16, 32], one output channel,
"bf16 operands / fp32 accumulate") through ``ops.set_gemm_precision("bf16")``: every dense linear runs as ONE
bf16 tensor-core MMA per product on the hi terms of the operands (fp32 accumulation, fp32 activations in HBM; attention,
LayerNorm, softmax, GELU as in the fp32 configuration), the decoder tail through the fused C = 128 kernel.

Stated tolerance of this mode (relative L2 against the fp32 CPU oracle): forward 1.5e-2, gradients 5e-2 -- bf16 operands
carry 8 mantissa bits (2^-9 relative per operand); the fp32 mode holds 1e-3 / 5e-3 on the same model (checked here too)."""
import pytest
import torch

from oracle import hp_oracle as O
from tests.util import build_product_model, rel_err

pytestmark = pytest.mark.gpu

KW = dict(patch_size=4, window_size=64, shift_size=4, shift_strategy="nest_roll", rel_pos_bias="flat", embed_dim=128,
          depths=[2, 2, 2], num_heads=[4, 8, 16], use_cos_attn=True, use_v2_norm_placement=True,
          dim_in=12 * 32 * 32, f_in=3, f_out=1, base_pix=12)
KEYS = ("layers.0.blocks.1.attn.qkv.weight", "layers.1.blocks.0.mlp.fc1.weight", "decoder.up.expand.weight",
        "layers.2.blocks.1.mlp.fc2.weight", "decoder.output.weight")


@pytest.fixture
def bf16_mode():
    from heal_swin_b200 import ops

    ops.set_gemm_precision("bf16")
    yield ops
    ops.set_gemm_precision("fp32")


def _run(dev):
    cfg = O.HPConfig(**KW)
    sd = O.synth_state_dict(cfg, seed=3)
    x = torch.randn(2, 3, KW["dim_in"], generator=torch.Generator().manual_seed(4))
    sd_g = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    want = O.hp_unet_forward(x, sd_g, cfg)
    wgt = torch.randn(want.shape, generator=torch.Generator().manual_seed(5))
    (want * wgt).sum().backward()
    model = build_product_model(KW, sd, dev).train()
    got = model(x.to(dev))
    (got * wgt.to(dev)).sum().backward()
    params = dict(model.named_parameters())
    return (rel_err(got.detach().cpu(), want.detach()),
            {k: rel_err(params[k].grad.cpu(), sd_g[k].grad) for k in KEYS})


def test_depth_head_config_in_bf16_operand_mode(bf16_mode):
    dev = torch.device("cuda:0")
    bf16_mode.STATS.reset()
    fwd, grads = _run(dev)
    assert 1e-3 < fwd < 1.5e-2, fwd  # (above 1e-3: this really is the reduced-precision path)
    for k, e in grads.items():
        assert e < 5e-2, (k, e)
    n = bf16_mode.STATS.by_name
    assert n.get("gemm3", 0) > 0 and n.get("ln_head_fwd", 0) == 1 and n.get("ln_head_bwd", 0) == 1, n  # fused C=128 tail


def test_depth_head_config_in_fp32_mode_holds_the_fp32_tolerances():
    fwd, grads = _run(torch.device("cuda:0"))
    assert fwd < 1e-3, fwd
    for k, e in grads.items():
        assert e < 5e-3, (k, e)
