"""GPU parity of the tcgen05 / TMA windowed-attention kernels (window 64, head_dim 32; TF32 operands,
fp32 accumulation) through the C-ABI.

Tolerance: 2e-3 relative L2 on the attention-core output (TF32 has a 10-bit mantissa: 2^-11 per
operand element; the logits of cosine attention are scaled by up to 100).  The model-level bound of
BASELINE.json (1e-3 on the network output) is checked in test_gpu_model.py with these kernels active.
"""
import numpy as np
import pytest
import torch

from oracle import hp_oracle as O
from scripts.tc_check import run_case, run_case_bwd
from tests.util import rel_err

pytestmark = pytest.mark.gpu

TC_TOL = 2e-3
BWD_TC_TOL = 3e-3  # gradients: TF32 operands in five chained products (S, dP, dV, dQ, dK)

CASES = [
    (1, 4, 8, 1, "none", False, False),
    (2, 8, 8, 3, "none", True, True),
    (1, 4, 12, 3, "nest_roll", True, True),        # the wrap-around window goes through the gather path
    (3, 4, 8, 2, "nest_roll", False, True),        # odd number of units: half-empty last pair
    (1, 8, 8, 3, "nest_grid_shift", True, True),
    (2, 8, 8, 6, "ring_shift", True, True),
    (1, 16, 8, 24, "ring_shift", False, False),
    (2, 32, 12, 3, "nest_roll", True, True),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"B{c[0]}-ns{c[1]}-bp{c[2]}-H{c[3]}-{c[4]}-cos{int(c[5])}-bias{int(c[6])}")
def test_tc_forward_matches_exact_fp32_kernels(case):
    err, _ = run_case(*case, torch.device("cuda:0"))
    assert err < TC_TOL, err


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"B{c[0]}-ns{c[1]}-bp{c[2]}-H{c[3]}-{c[4]}-cos{int(c[5])}-bias{int(c[6])}")
def test_tc_backward_matches_exact_fp32_kernels(case):
    """dqkv, d(relative_position_bias_table), d(logit_scale) of the tcgen05 backward vs the fp32 CUDA-core backward."""
    errs = run_case_bwd(*case, torch.device("cuda:0"))
    for k, e in errs.items():
        # d(logit_scale) is one scalar per head: a sum of dS * cos over all windows whose terms largely cancel, so the
        # TF32 noise of S is amplified relative to the (small) total
        assert e < (1e-2 if k == "dlogit_scale" else BWD_TC_TOL), (k, e)


def test_tc_backward_full_size_linearity_property():
    """BASELINE configs[1] stage-0 size.  The backward is linear in dO: bwd(a*dO1 + dO2) == a*bwd(dO1) + bwd(dO2)
    (size-independent property; exercises every window / head of the full-size launch)."""
    from heal_swin_b200 import ops

    dev = torch.device("cuda:0")
    B, N, C, H, ws = 8, 12 * 128 * 128, 96, 3, 64
    g = torch.Generator().manual_seed(3)
    qkv = torch.randn(2, N, 3 * C, generator=g).to(dev).repeat(B // 2, 1, 1).requires_grad_(True)
    out = ops.window_attention_core(qkv, None, None, None, None, None, None, 32 ** -0.5, H, ws, False)
    d1 = torch.randn(B, N, C, device=dev)
    d2 = torch.randn(B, N, C, device=dev)
    g1, = torch.autograd.grad(out, qkv, d1, retain_graph=True)
    g2, = torch.autograd.grad(out, qkv, d2, retain_graph=True)
    g12, = torch.autograd.grad(out, qkv, 0.5 * d1 + d2, retain_graph=True)
    assert rel_err((0.5 * g1 + g2).cpu(), g12.cpu()) < BWD_TC_TOL
    # and samples b, b+2 of the repeated batch see identical q, k, v: identical dO must give identical gradients
    d3 = d1[:2].repeat(B // 2, 1, 1)
    g3, = torch.autograd.grad(out, qkv, d3)
    assert torch.equal(g3[0], g3[2]) and torch.equal(g3[1], g3[7])


@pytest.mark.parametrize("cos", [False, True])
def test_tc_backward_is_bit_identical_run_to_run(cos):
    """dqkv has no atomics in its path: every element is written once by one thread, so repeated launches on the same
    inputs must agree bit for bit.  A race in the kernel's barrier protocol (warp roles that alternate between units,
    ring slots, staging tiles) shows up here as run-to-run differences -- with a nest_roll shift so that both the TMA
    and the gathered-window paths run, at a size that gives every CTA dozens of units."""
    from heal_swin_b200 import hp_index, ops
    from scripts.tc_check import tables

    dev = torch.device("cuda:0")
    B, nside, base_pix, H, ws, D = 2, 128, 12, 3, 64, 32
    C, N = H * D, base_pix * nside * nside
    g = torch.Generator().manual_seed(11)
    qkv = torch.randn(B, N, 3 * C, generator=g).to(dev)
    table = (torch.randn(225, H, generator=g) * 0.5).to(dev)
    rel_index = hp_index.rel_pos_index(ws).to(torch.int32).reshape(-1).contiguous().to(dev)
    ls = (torch.log(torch.tensor(10.0)) + 0.3 * torch.randn(H, 1, 1, generator=g)).to(dev) if cos else None
    src, grp = tables("nest_roll", nside, base_pix, ws, dev)
    dout = torch.randn(B, N, C, generator=g).to(dev)
    grads = []
    for _ in range(4):
        q = qkv.clone().requires_grad_(True)
        out = ops.window_attention_core(q, table, ls, src, grp, None, rel_index, D ** -0.5, H, ws, cos)
        (gq,) = torch.autograd.grad(out, q, dout)
        grads.append(gq)
    for gq in grads[1:]:
        assert torch.equal(grads[0], gq)


def test_tc_forward_vs_oracle_through_the_block():
    """SwinTransformerBlock attention (ring shift, cos, bias) on the tensor-core path vs the CPU oracle."""
    from heal_swin_b200 import ops
    from heal_swin_b200.models_torch.swin_hp_transformer import SwinTransformerBlock
    from tests.test_gpu_attention import _oracle_attention

    assert ops.get_attention_precision() == "tf32"
    dev = torch.device("cuda:0")
    nside, bp, ws, C, h = 16, 8, 64, 96, 3
    N = bp * nside * nside
    g = torch.Generator().manual_seed(7)
    blk = SwinTransformerBlock(C, N, bp, h, window_size=ws, shift_size=4, shift_strategy="ring_shift",
                               rel_pos_bias="flat", use_cos_attn=True)
    with torch.no_grad():
        for p in blk.attn.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * 0.3)
        blk.attn.logit_scale.copy_(np.log(10.0) + 0.3 * torch.randn(h, 1, 1, generator=g))
    sd = {"a." + k: v.detach().clone() for k, v in blk.attn.state_dict().items()}
    x = torch.randn(2, N, C, generator=g)
    tabs = O.make_shift_tables("ring_shift", 4, N, bp, ws)
    with torch.no_grad():
        want = _oracle_attention(x, sd, h, ws, True, True, tabs.shift_idcs, tabs.groups)
        blk = blk.to(dev)
        got = blk.attn.forward_tokens(x.to(dev), ws, blk._hs_src, blk._hs_groups).cpu()
    assert rel_err(got, want) < TC_TOL


def test_tc_full_size_uniform_softmax_property():
    """BASELINE configs[1] stage-0 size: q = k = 0 makes every softmax row uniform, so each window's output is
    the window mean of v (exercises the addressing of all 8 x 3072 windows x 3 heads through TMA)."""
    from heal_swin_b200 import ops

    dev = torch.device("cuda:0")
    B, N, C, H, ws = 8, 12 * 128 * 128, 96, 3, 64
    qkv = torch.zeros(B, N, 3 * C, device=dev)
    v = torch.randn(B, N, C, device=dev)
    qkv[:, :, 2 * C:] = v
    out = ops.window_attention_core(qkv, None, None, None, None, None, None, 1.0, H, ws, False)
    want = v.view(B, N // ws, ws, C).mean(2, keepdim=True).expand(-1, -1, ws, -1).reshape(B, N, C)
    assert rel_err(out.cpu(), want.cpu()) < TC_TOL
