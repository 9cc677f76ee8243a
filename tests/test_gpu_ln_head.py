"""GPU parity of the fused decoder tail  logits = Conv1d_1x1(LayerNorm(x))  (csrc/hs_ln_head.cu, exact fp32) and of its
one-pass backward against plain torch autograd of swin_hp_transformer.py:450 + 781-786."""
import pytest
import torch
import torch.nn.functional as F

from tests.util import rel_err

pytestmark = pytest.mark.gpu


def _reference(x, norm, w, b):
    """LayerNorm + 1x1 Conv1d written as an fp32 matmul (F.conv1d itself runs in TF32 under cudnn's default)."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        y = F.layer_norm(x, (x.shape[-1],), norm.weight, norm.bias, norm.eps)
        out = torch.matmul(y, w.t())
        if b is not None:
            out = out + b
        return out.transpose(1, 2)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


@pytest.mark.parametrize("B,P,Cc,K,bias", [(2, 1000, 96, 10, False), (3, 77, 64, 1, True), (1, 5003, 32, 16, True),
                                           (2, 4096, 96, 3, False), (1, 9, 96, 13, True), (4, 20000, 96, 10, False),
                                           (2, 3000, 128, 1, False), (1, 777, 128, 10, True)])  # C = 128: BASELINE configs[3]
def test_ln_head_forward_and_gradients_match_torch(B, P, Cc, K, bias):
    from heal_swin_b200 import ops

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(B * P + K)
    x = (torch.randn(B, P, Cc, generator=g) * 2 + 0.5).to(dev).requires_grad_(True)
    norm = torch.nn.LayerNorm(Cc).to(dev)
    with torch.no_grad():
        norm.weight.copy_(1 + 0.3 * torch.randn(Cc, generator=g))
        norm.bias.copy_(0.2 * torch.randn(Cc, generator=g))
    w = (torch.randn(K, Cc, generator=g) / Cc ** 0.5).to(dev).requires_grad_(True)
    b = torch.randn(K, generator=g).to(dev).requires_grad_(True) if bias else None
    gy = torch.randn(B, K, P, generator=g).to(dev)
    ops.STATS.reset()
    y = ops.ln_head(x, norm, w, b)
    assert ops.STATS.launches == 1 and y.shape == (B, K, P) and y.is_contiguous()
    y.backward(gy)
    assert ops.STATS.launches == 2
    got = [y.detach().clone(), x.grad.clone(), norm.weight.grad.clone(), norm.bias.grad.clone(), w.grad.clone()]
    if bias:
        got.append(b.grad.clone())
    x.grad = w.grad = None
    norm.zero_grad()
    if bias:
        b.grad = None
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False  # fp32 backward GEMMs for the reference
    y2 = _reference(x, norm, w, b)
    y2.backward(gy)
    torch.backends.cuda.matmul.allow_tf32 = prev
    want = [y2.detach(), x.grad, norm.weight.grad, norm.bias.grad, w.grad] + ([b.grad] if bias else [])
    for i, (a, ww) in enumerate(zip(got, want)):
        assert rel_err(a.cpu(), ww.cpu()) < (1e-5 if i == 0 else 2e-4), i


def test_ln_head_falls_back_for_uncovered_shapes():
    from heal_swin_b200 import ops

    dev = torch.device("cuda:0")
    x = torch.randn(2, 100, 48, device=dev)  # C = 48 is not covered by the fused kernel
    norm = torch.nn.LayerNorm(48).to(dev)
    w = torch.randn(5, 48, device=dev)
    y = ops.ln_head(x, norm, w, None)
    assert rel_err(y.cpu(), _reference(x, norm, w, None).cpu()) < 1e-5
    w20 = torch.randn(20, 96, device=dev)  # K = 20 > 16
    x96 = torch.randn(2, 100, 96, device=dev)
    norm96 = torch.nn.LayerNorm(96).to(dev)
    assert rel_err(ops.ln_head(x96, norm96, w20, None).cpu(), _reference(x96, norm96, w20, None).cpu()) < 1e-4


def test_ln_head_full_size_linearity():
    """BASELINE size (8 x 786432 rows): logits are affine in the head weights and dx is linear in d(logits)."""
    from heal_swin_b200 import ops

    dev = torch.device("cuda:0")
    B, P, Cc, K = 8, 786432, 96, 10
    x = torch.randn(B, P, Cc, device=dev).requires_grad_(True)
    norm = torch.nn.LayerNorm(Cc).to(dev)
    w1 = torch.randn(K, Cc, device=dev) / 10
    w2 = torch.randn(K, Cc, device=dev) / 10
    y1, y2, y12 = ops.ln_head(x, norm, w1), ops.ln_head(x, norm, w2), ops.ln_head(x, norm, w1 + 0.5 * w2)
    assert rel_err((y1 + 0.5 * y2).detach().cpu(), y12.detach().cpu()) < 1e-5
    g1 = torch.randn(B, K, P, device=dev)
    (d1,) = torch.autograd.grad(y1, x, g1, retain_graph=True)
    (d2,) = torch.autograd.grad(y1, x, 2 * g1)
    assert rel_err((2 * d1).cpu(), d2.cpu()) < 1e-5
