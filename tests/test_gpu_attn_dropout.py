"""Attention-probability dropout inside the windowed-attention kernels (nn.Dropout(attn_drop) on the softmax output,
swin_hp_transformer.py:167-169).  The RNG stream of torch cannot be matched, so parity is checked the other way
round: the mask the kernels use is a documented pure function of (seed, window, head, i, j) (csrc/hs_common.h); the test
rebuilds it on the host and compares the kernels with plain torch attention using that very mask (forward and
gradients), for the exact-fp32 kernels (1e-4) and the tcgen05 TF32 kernels (2e-3 / 3e-3)."""
import numpy as np
import pytest
import torch

from tests.util import rel_err

pytestmark = pytest.mark.gpu
M32 = 0xFFFFFFFF


def _mix32(x):
    x = x.astype(np.uint64)
    x ^= x >> 16
    x = (x * 0x7FEB352D) & M32
    x ^= x >> 15
    x = (x * 0x846CA68B) & M32
    x ^= x >> 16
    return x


def host_mask(seed, n_units, H, ws, p):
    """(n_units, H, ws, ws) float32 multiplier: 0 where dropped, 1/(1-p) where kept."""
    thresh = min(int(float(np.float32(p)) * 4294967296.0), M32)
    wb = np.arange(n_units, dtype=np.uint64)[:, None]
    h = np.arange(H, dtype=np.uint64)[None, :]
    inner = _mix32(((seed >> 32) + wb * H + h) & M32)
    key = _mix32((np.uint64(seed & M32) ^ inner) & M32)  # (n_units, H)
    e = (np.arange(ws * ws, dtype=np.uint64) * 0x9E3779B9) & M32
    hsh = _mix32((key[:, :, None] ^ e[None, None, :]) & M32)
    keep = hsh >= thresh
    scale = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    return (keep.astype(np.float32) * scale).reshape(n_units, H, ws, ws)


def torch_attention(qkv, mask_mult, H, ws, scale, table_bias):
    B, N, C3 = qkv.shape
    C = C3 // 3
    d = C // H
    q, k, v = qkv.view(B, N // ws, ws, 3, H, d).permute(3, 0, 1, 4, 2, 5)  # (B, nW, H, ws, d)
    attn = (q * scale) @ k.transpose(-2, -1)
    if table_bias is not None:
        attn = attn + table_bias
    attn = attn.softmax(-1) * mask_mult.view(B, N // ws, H, ws, ws)
    return (attn @ v).permute(0, 1, 3, 2, 4).reshape(B, N, C)


@pytest.mark.parametrize("mode,ws,H,d,ftol,gtol", [("fp32", 64, 3, 32, 1e-4, 1e-4), ("tf32", 64, 3, 32, 2e-3, 3e-3),
                                                    ("fp32", 16, 2, 8, 1e-4, 1e-4)])
def test_dropout_matches_torch_with_the_same_mask(mode, ws, H, d, ftol, gtol):
    from heal_swin_b200 import hp_index, ops

    dev = torch.device("cuda:0")
    B, nW, p, seed = 2, 12, 0.25, 0x1234_5678_9ABC_DEF0
    N, C = nW * ws, H * d
    g = torch.Generator().manual_seed(5)
    qkv0 = torch.randn(B, N, 3 * C, generator=g)
    S = int(ws ** 0.5)
    table0 = torch.randn((2 * S - 1) ** 2, H, generator=g) * 0.5
    rel = hp_index.rel_pos_index(ws)
    wgt = torch.randn(B, N, C, generator=g)
    mult = torch.from_numpy(host_mask(seed, B * nW, H, ws, p))

    ops.set_attention_precision(mode)
    try:
        qkv = qkv0.clone().to(dev).requires_grad_(True)
        table = table0.clone().to(dev).requires_grad_(True)
        out = ops.window_attention_core(qkv, table, None, None, None, None,
                                        rel.to(torch.int32).reshape(-1).contiguous().to(dev), d ** -0.5, H, ws, False,
                                        attn_drop=p, seed=seed)
        (out * wgt.to(dev)).sum().backward()
    finally:
        ops.set_attention_precision("tf32")

    q2 = qkv0.clone().requires_grad_(True)
    t2 = table0.clone().requires_grad_(True)
    bias = t2[rel.view(-1)].view(ws, ws, H).permute(2, 0, 1)
    ref = torch_attention(q2, mult, H, ws, d ** -0.5, bias)
    (ref * wgt).sum().backward()
    assert rel_err(out.detach().cpu(), ref.detach()) < ftol
    assert rel_err(qkv.grad.cpu(), q2.grad) < gtol
    assert rel_err(table.grad.cpu(), t2.grad) < gtol


def test_dropout_rate_seed_and_eval_behaviour():
    from heal_swin_b200 import ops

    dev = torch.device("cuda:0")
    B, N, H, ws = 2, 64 * 40, 3, 64
    C = 32 * H
    # v = 1 everywhere, q = k = 0: every row of P is uniform 1/64, so out = (number of kept entries) / 64 / (1 - p)
    qkv = torch.zeros(B, N, 3 * C, device=dev)
    qkv[:, :, 2 * C:] = 1.0
    p = 0.1
    o1 = ops.window_attention_core(qkv, None, None, None, None, None, None, 1.0, H, ws, False, attn_drop=p, seed=11)
    kept = float(o1.mean()) * (1 - p)
    assert abs(kept - (1 - p)) < 3e-3  # 2 * 2560 * 3 * 64 = 983040 Bernoulli draws per channel
    o1b = ops.window_attention_core(qkv, None, None, None, None, None, None, 1.0, H, ws, False, attn_drop=p, seed=11)
    o2 = ops.window_attention_core(qkv, None, None, None, None, None, None, 1.0, H, ws, False, attn_drop=p, seed=12)
    assert torch.equal(o1, o1b) and not torch.equal(o1, o2)
    o0 = ops.window_attention_core(qkv, None, None, None, None, None, None, 1.0, H, ws, False, attn_drop=0.0)
    assert float((o0 - 1).abs().max()) < 1e-3


def test_module_trains_with_attn_drop_and_is_deterministic_in_eval():
    from heal_swin_b200.models_torch.swin_hp_transformer import SwinTransformerBlock

    dev = torch.device("cuda:0")
    blk = SwinTransformerBlock(96, 8 * 16 * 16, 8, 3, window_size=64, shift_size=4, shift_strategy="ring_shift",
                               rel_pos_bias="flat", attn_drop=0.1, drop=0.1, use_cos_attn=True,
                               use_v2_norm_placement=True).to(dev)
    x = torch.randn(2, 8 * 16 * 16, 96, device=dev, requires_grad=True)
    blk.train()
    y1, y2 = blk(x), blk(x)
    assert not torch.equal(y1, y2)  # fresh masks per call
    y1.square().mean().backward()
    assert torch.isfinite(x.grad).all() and float(x.grad.abs().max()) > 0
    blk.eval()
    with torch.no_grad():
        assert torch.equal(blk(x), blk(x))
