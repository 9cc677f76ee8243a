"""Host-side logic of the autograd nodes in heal_swin_b200/ops.py, run on CPU with the device library replaced by a
torch emulation of each C-ABI entry point (same argument order as include/healswin_b200.h): which kernels a node calls,
how the shortcut gradient of a forked node is routed, and the parameter-gradient algebra of the fused decoder tail
(dW = gamma * S + beta * G, dgamma = sum_k W * S, dbeta = sum_k W * G).  The kernels themselves are checked on the GPU."""
import math

import pytest
import torch
import torch.nn.functional as F

from heal_swin_b200 import ops


def _val(v):
    return getattr(v, "value", v)


def _gelu_grad(u):
    return 0.5 * (1 + torch.erf(u / math.sqrt(2))) + u * torch.exp(-0.5 * u * u) / math.sqrt(2 * math.pi)


class FakeLib:
    """torch emulation of the entry points the three nodes use; tensors arrive in place of device pointers."""

    def __init__(self):
        self.calls = []

    def hs_linear_wgrad_supported(self, T, N, K):
        return 2

    def hs_mlp_dgrad_gelu_supported(self, T, Cc, J):
        return 1

    def hs_ln_head_supported(self, rows, Cc, K):
        return 1

    def hs_linear_wgrad(self, dy, x, dw, db, T, N, K, flags, stream):
        self.calls.append("wgrad")
        dw += dy.t() @ x
        if db is not None:
            db += dy.sum(0)
        return 0

    def hs_linear_fwd(self, x, w, b, y, T, N, K, ws, ws_bytes, stream):
        self.calls.append("fwd")
        y.copy_(x @ w.t() + (b if b is not None else 0))
        return 0

    def hs_linear_dgrad_acc(self, dy, w, c, dx, T, N, K, ws, ws_bytes, stream):
        self.calls.append("dgrad_acc" if c is not None else "dgrad")
        dx.copy_(dy @ w + (c if c is not None else 0))
        return 0

    def hs_bias_gelu_fwd(self, z, b, drop, seed, h, rows, Cc, stream):
        assert _val(drop) == 0.0
        h.copy_(F.gelu(z + b))
        return 0

    def hs_mlp_dgrad_gelu(self, dy, w2, z, b1, drop, seed, dz, T, Cc, J, flags, stream):
        self.calls.append("mlp_dgrad_gelu")
        dz.copy_((dy @ w2) * _gelu_grad(z + b1))
        return 0

    def hs_ln_head_fwd(self, x, gamma, beta, w, hb, logits, mean, rstd, rows, P, Cc, K, eps, stream):
        mu = x.mean(1)
        rs = torch.rsqrt(x.var(1, unbiased=False) + _val(eps))
        y = (x - mu[:, None]) * rs[:, None] * gamma + beta
        out = y @ w.t() + (hb if hb is not None else 0)
        logits.copy_(out.view(rows // P, P, K).transpose(1, 2))
        mean.copy_(mu)
        rstd.copy_(rs)
        return 0

    def hs_ln_head_bwd(self, dl, x, mean, rstd, gamma, w, dx, s_acc, g_acc, rows, P, Cc, K, stream):
        g = dl.transpose(1, 2).reshape(rows, K)
        xh = (x - mean[:, None]) * rstd[:, None]
        wv = (g @ w) * gamma
        dx.copy_(rstd[:, None] * (wv - wv.mean(1, keepdim=True) - xh * (wv * xh).mean(1, keepdim=True)))
        s_acc += g.t() @ xh
        g_acc += g.sum(0)
        return 0


@pytest.fixture
def fake(monkeypatch):
    lib = FakeLib()
    monkeypatch.setattr(ops, "lib", lib)
    monkeypatch.setattr(ops, "ptr", lambda t: t)
    monkeypatch.setattr(ops, "current_stream", lambda: None)
    monkeypatch.setattr(ops, "require_cuda", lambda *a: None)
    monkeypatch.setattr(ops, "check", lambda rc: None)
    monkeypatch.setattr(ops.STATS, "launch", lambda name, fn, *a, tag=None: fn(*a))
    monkeypatch.setattr(ops, "_LT_GEMM", True)
    return lib


def _close(a, b, tol=1e-5):
    return float((a - b).norm() / b.norm().clamp_min(1e-30)) < tol


def test_forked_linear_routes_the_shortcut_gradient_through_the_dgrad_gemm(fake):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 5, 8, generator=g, requires_grad=True)
    w = torch.randn(6, 8, generator=g, requires_grad=True)
    b = torch.randn(6, generator=g, requires_grad=True)
    wy, wsc = torch.randn(2, 5, 6, generator=g), torch.randn(2, 5, 8, generator=g)
    y, shortcut = ops._LinearFn.apply(x, w, b, True)
    ((y * wy).sum() + (shortcut * wsc).sum()).backward()
    assert fake.calls == ["fwd", "dgrad_acc", "wgrad"]
    got = (x.grad.clone(), w.grad.clone(), b.grad.clone())
    x.grad = w.grad = b.grad = None
    ((F.linear(x, w, b) * wy).sum() + (x * wsc).sum()).backward()
    for a, ref in zip(got, (x.grad, w.grad, b.grad)):
        assert _close(a, ref)


def test_forked_linear_with_only_one_output_used(fake):
    x = torch.randn(3, 8, requires_grad=True)
    w = torch.randn(6, 8, requires_grad=True)
    y, shortcut = ops._LinearFn.apply(x, w, None, True)
    shortcut.sum().backward()  # the branch is unused: no GEMM at all in the backward
    assert fake.calls == ["fwd"] and torch.equal(x.grad, torch.ones_like(x)) and w.grad is None
    x.grad = None
    fake.calls.clear()
    y, shortcut = ops._LinearFn.apply(x, w, None, True)
    y.sum().backward()  # the shortcut is unused: plain dgrad, no C operand
    assert fake.calls == ["fwd", "dgrad", "wgrad"] and _close(x.grad, torch.ones(3, 6) @ w.detach())


def test_plain_linear_node_matches_torch(fake):
    x = torch.randn(7, 8, requires_grad=True)
    w = torch.randn(4, 8, requires_grad=True)
    y = ops._LinearFn.apply(x, w, None, False)
    assert isinstance(y, torch.Tensor)
    y.square().sum().backward()
    got = (x.grad.clone(), w.grad.clone())
    x.grad = w.grad = None
    F.linear(x, w).square().sum().backward()
    assert _close(got[0], x.grad) and _close(got[1], w.grad)


@pytest.mark.parametrize("fork", [False, True])
def test_fused_mlp_node_matches_torch_autograd(fake, fork):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 6, 8, generator=g, requires_grad=True)
    w1 = (torch.randn(32, 8, generator=g) / 3).requires_grad_(True)
    b1 = torch.randn(32, generator=g, requires_grad=True)
    w2 = (torch.randn(8, 32, generator=g) / 6).requires_grad_(True)
    gy = torch.randn(2, 6, 8, generator=g)
    out = ops._MlpFn.apply(x, w1, b1, w2, 0.0, 0, fork)
    y = out[0] + out[1] if fork else out
    y.backward(gy)
    # wgrad(fc2), fused dgrad + GELU', wgrad(fc1) + bias, library dgrad (with the shortcut gradient when forked)
    assert fake.calls == ["fwd", "fwd", "wgrad", "mlp_dgrad_gelu", "wgrad", "dgrad_acc" if fork else "dgrad"]
    got = [t.grad.clone() for t in (x, w1, b1, w2)]
    for t in (x, w1, b1, w2):
        t.grad = None
    ref = F.linear(F.gelu(F.linear(x, w1, b1)), w2)
    (ref + x if fork else ref).backward(gy)
    assert _close(y.detach(), (ref + x if fork else ref).detach())
    for a, t in zip(got, (x, w1, b1, w2)):
        assert _close(a, t.grad, 1e-4)


@pytest.mark.parametrize("bias", [False, True])
def test_decoder_tail_parameter_gradients_follow_from_s_and_g(fake, bias):
    g = torch.Generator().manual_seed(2)
    B, P, Cc, K = 2, 11, 8, 3
    x = torch.randn(B, P, Cc, generator=g, requires_grad=True)
    gamma = (1 + 0.3 * torch.randn(Cc, generator=g)).requires_grad_(True)
    beta = (0.2 * torch.randn(Cc, generator=g)).requires_grad_(True)
    w = torch.randn(K, Cc, generator=g, requires_grad=True)
    hb = torch.randn(K, generator=g, requires_grad=True) if bias else None
    gy = torch.randn(B, K, P, generator=g)
    params = [x, gamma, beta, w] + ([hb] if bias else [])
    y = ops.LnHeadFn.apply(x, gamma, beta, w, hb, 1e-5)
    assert y.shape == (B, K, P)
    y.backward(gy)
    got = [t.grad.clone() for t in params]
    for t in params:
        t.grad = None
    ref = F.linear(F.layer_norm(x, (Cc,), gamma, beta, 1e-5), w, hb).transpose(1, 2)
    ref.backward(gy)
    assert _close(y.detach(), ref.detach())
    for a, t in zip(got, params):
        assert _close(a, t.grad, 1e-4)
