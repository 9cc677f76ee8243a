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
        self.splits = 0
        self.formats = {}

    wgrad_cover = 2

    def hs_linear_wgrad_supported(self, T, N, K):
        return self.wgrad_cover

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

    def hs_gemm3_supported(self, T, N, K):
        return 1 if (N % 4 == 0 and K % 4 == 0) else 0

    def hs_bias_gelu_supported(self, rows, Cc):
        return 1

    def hs_weight_split(self, w, rows, cols, ld, transposed, fmt, out, stream):
        """[hi(32) | lo(32)] bf16 per 32-wide chunk of the contraction axis, zero padded (include/healswin_b200.h)."""
        self.splits += 1
        m = (w.t() if transposed else w).float()
        assert m.shape == (rows, cols) and ld == w.shape[1]
        nk = (cols + 31) // 32
        pad = torch.zeros(rows, nk * 32)
        pad[:, :cols] = m
        if fmt == 1:  # fp32 (rounded to TF32 on the device) viewed through the bf16 buffer
            out.view(torch.float32).copy_(pad)
            self.formats[id(out)] = 1
            return 0
        self.formats[id(out)] = 0
        hi = pad.bfloat16()
        lo = (pad - hi.float()).bfloat16()
        assert out.shape == (rows, 2 * nk * 32) and out.dtype == torch.bfloat16
        o = out.view(rows, nk, 64)
        o[:, :, :32] = hi.view(rows, nk, 32)
        o[:, :, 32:] = lo.view(rows, nk, 32)
        return 0

    def hs_gemm3(self, a, ws, bias, aux, d, d2, colsum, T, N, K, mode, prec, drop, seed, stream):
        self.calls.append(f"gemm3:{mode}" + ("+colsum" if colsum is not None else "") + ("/tf32" if prec == 1 else ""))
        assert self.formats[id(ws)] == (1 if prec == 1 else 0), "weight operand format does not match the precision"
        if colsum is not None:
            colsum += a.sum(0)
        assert _val(drop) == 0.0 and a.shape == (T, K) and ws.shape[0] == N
        if prec == 1:
            w = ws.view(torch.float32)[:, :K]
        else:
            o = ws.view(N, -1, 64).float()
            w = (o[:, :, :32] + o[:, :, 32:]).reshape(N, -1)[:, :K]   # hi + lo: the weight to ~2^-17
        acc = a @ w.t()
        b = bias if bias is not None else 0
        if mode == 0:
            d.copy_(acc + b)
        elif mode == 1:
            d.copy_(acc + b + aux)
        elif mode == 2:
            d.copy_(acc)
            d2.copy_(F.gelu(acc + b))
        elif mode == 6:  # compact: the activation's derivative as FP16 instead of z
            assert d.dtype == torch.float16
            d.copy_(_gelu_grad(acc + b))
            d2.copy_(F.gelu(acc + b))
        elif mode == 7:
            assert aux.dtype == torch.float16 and bias is None
            d.copy_(acc * aux.float())
        else:
            d.copy_(acc * _gelu_grad(aux + b))
        return 0

    def _weight_of(self, ws, N, K):
        o = ws.view(N, -1, 64).float()
        return (o[:, :, :32] + o[:, :, 32:]).reshape(N, -1)[:, :K]

    def hs_gemm3_ln_supported(self, T, N, K, G):
        return 1 if (G % 32 == 0 and G <= 192 and N % G == 0) else 0

    def hs_gemm3_ln(self, a, ws, bias, gamma, beta, aux, pre, y, mean, rstd, T, N, K, G, eps, prec, stream):
        """y = [aux +] LayerNorm over groups of G columns of (a W^T + bias); pre / mean / rstd optional (include/healswin_b200.h)."""
        self.calls.append("gemm3_ln" + ("+aux" if aux is not None else "") + ("+save" if pre is not None else ""))
        p = a @ self._weight_of(ws, N, K).t() + (bias if bias is not None else 0)
        v = p.view(T * (N // G), G)
        mu, rs = v.mean(1), torch.rsqrt(v.var(1, unbiased=False) + _val(eps))
        out = (((v - mu[:, None]) * rs[:, None]) * gamma + beta).view(T, N)
        y.copy_(out + (aux if aux is not None else 0))
        if pre is not None:
            pre.copy_(p)
        if mean is not None:
            mean.copy_(mu)
            rstd.copy_(rs)
        return 0

    def hs_gemm3_lnin_supported(self, T, N, K):
        return 1 if (K % 32 == 0 and K >= 160) else 0

    def hs_gemm3_lnin(self, a, ws, wsum, b0, d, mean, rstd, T, N, K, eps, prec, stream):
        """d = rstd (a W'^T - mean wsum) + b0 with the row statistics of a (include/healswin_b200.h)."""
        self.calls.append("gemm3_lnin")
        mu, rs = a.mean(1), torch.rsqrt(a.var(1, unbiased=False) + _val(eps))
        acc = a @ self._weight_of(ws, N, K).t()
        d.copy_(rs[:, None] * (acc - mu[:, None] * wsum) + (b0 if b0 is not None else 0))
        if mean is not None:
            mean.copy_(mu)
            rstd.copy_(rs)
        return 0

    def hs_layernorm_fwd(self, x, pb, res, gamma, beta, rsc, rps, in_drop, seed, y, mean, rstd, rows, Cc, eps, stream):
        self.calls.append("ln_fwd")
        assert pb is None and res is None and rsc is None and _val(in_drop) == 0.0
        y.copy_(F.layer_norm(x, (Cc,), gamma, beta, _val(eps)))
        return 0

    def hs_layernorm_bwd(self, dy, x, pb, mean, rstd, gamma, rsc, rps, in_drop, seed, dx, dw, db, dpb, rows, Cc, stream):
        """dx of LayerNorm from the saved statistics; dgamma / dbeta / d(pre-bias) ACCUMULATE (the kernel adds atomically)."""
        self.calls.append("ln_bwd")
        assert rsc is None and _val(in_drop) == 0.0
        xv, dyv = x.reshape(rows, Cc) + (pb if pb is not None else 0), dy.reshape(rows, Cc)
        xh = (xv - mean[:, None]) * rstd[:, None]
        wv = dyv * gamma
        out = rstd[:, None] * (wv - wv.mean(1, keepdim=True) - xh * (wv * xh).mean(1, keepdim=True))
        dx.view(rows, Cc).copy_(out)
        if dw is not None:
            dw += (dyv * xh).sum(0)
        if db is not None:
            db += dyv.sum(0)
        if dpb is not None:
            dpb += out.sum(0)
        return 0

    def hs_weight_split_batch(self, table, n, total_tiles, stream):
        self.calls.append(f"split_batch:{n}")
        return 0

    def hs_bias_gelu_fwd(self, z, b, drop, seed, h, rows, Cc, stream):
        assert _val(drop) == 0.0
        h.copy_(F.gelu(z + b))
        return 0

    def hs_mlp_dgrad_gelu(self, dy, w2, z, b1, drop, seed, dz, T, Cc, J, flags, stream):
        self.calls.append("mlp_dgrad_gelu")
        if flags & 256:  # HS_MLP_GRAD16: z is the FP16 derivative tensor
            assert z.dtype == torch.float16
            dz.copy_((dy @ w2) * z.float())
        else:
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
    monkeypatch.setattr(ops, "_GEMM_MODE", "bf16x3")
    monkeypatch.setattr(ops, "_on_device", lambda t: True)
    ops.invalidate_weight_splits()
    return lib


def _close(a, b, tol=1e-5):
    return float((a - b).norm() / b.norm().clamp_min(1e-30)) < tol


def test_forked_linear_routes_the_shortcut_gradient_through_the_dgrad_gemm(fake):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 5, 8, generator=g, requires_grad=True)
    w = torch.randn(12, 8, generator=g, requires_grad=True)
    b = torch.randn(12, generator=g, requires_grad=True)
    wy, wsc = torch.randn(2, 5, 12, generator=g), torch.randn(2, 5, 8, generator=g)
    y, shortcut = ops._LinearFn.apply(x, w, b, True)
    ((y * wy).sum() + (shortcut * wsc).sum()).backward()
    assert fake.calls == ["gemm3:0", "gemm3:1", "wgrad"]  # forward, dgrad with the shortcut gradient as aux, wgrad
    got = (x.grad.clone(), w.grad.clone(), b.grad.clone())
    x.grad = w.grad = b.grad = None
    ((F.linear(x, w, b) * wy).sum() + (x * wsc).sum()).backward()
    for a, ref in zip(got, (x.grad, w.grad, b.grad)):
        assert _close(a, ref, 1e-4)  # the emulated split weight is hi + lo in bf16: ~2^-17 relative


def test_forked_linear_with_only_one_output_used(fake):
    x = torch.randn(3, 8, requires_grad=True)
    w = torch.randn(12, 8, requires_grad=True)
    y, shortcut = ops._LinearFn.apply(x, w, None, True)
    shortcut.sum().backward()  # the branch is unused: no GEMM at all in the backward
    assert fake.calls == ["gemm3:0"] and torch.equal(x.grad, torch.ones_like(x)) and w.grad is None
    x.grad = None
    fake.calls.clear()
    y, shortcut = ops._LinearFn.apply(x, w, None, True)
    y.sum().backward()  # the shortcut is unused: plain dgrad, no C operand
    assert fake.calls == ["gemm3:0", "gemm3:0", "wgrad"] and _close(x.grad, torch.ones(3, 12) @ w.detach(), 1e-4)


def test_plain_linear_node_matches_torch(fake):
    x = torch.randn(7, 8, requires_grad=True)
    w = torch.randn(8, 8, requires_grad=True)
    y = ops._LinearFn.apply(x, w, None, False)
    assert isinstance(y, torch.Tensor)
    y.square().sum().backward()
    got = (x.grad.clone(), w.grad.clone())
    x.grad = w.grad = None
    F.linear(x, w).square().sum().backward()
    assert _close(got[0], x.grad, 1e-4) and _close(got[1], w.grad, 1e-4)


@pytest.mark.parametrize("fork", [False, True])
@pytest.mark.parametrize("compact", [False, True])
def test_fused_mlp_node_matches_torch_autograd(fake, fork, compact, monkeypatch):
    # compact: the forward saves GELU'(z + b1) as FP16 instead of z (hs_gemm3 modes 6 / 7, HS_MLP_GRAD16)
    monkeypatch.setattr(ops, "_MLP_COMPACT", "all" if compact else "0")
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 6, 8, generator=g, requires_grad=True)
    w1 = (torch.randn(32, 8, generator=g) / 3).requires_grad_(True)
    b1 = torch.randn(32, generator=g, requires_grad=True)
    w2 = (torch.randn(8, 32, generator=g) / 6).requires_grad_(True)
    gy = torch.randn(2, 6, 8, generator=g)
    out = ops._MlpFn.apply(x, w1, b1, w2, 0.0, 0, fork)
    y = out[0] + out[1] if fork else out
    y.backward(gy)
    # fc1 + GELU epilogue, fc2 | wgrad(fc2), fused dgrad + GELU', dgrad (+ shortcut gradient when forked), wgrad(fc1) + bias
    assert fake.calls == ["gemm3:6" if compact else "gemm3:2", "gemm3:0", "wgrad", "mlp_dgrad_gelu",
                          "gemm3:1" if fork else "gemm3:0", "wgrad"]
    got = [t.grad.clone() for t in (x, w1, b1, w2)]
    for t in (x, w1, b1, w2):
        t.grad = None
    ref = F.linear(F.gelu(F.linear(x, w1, b1)), w2)
    (ref + x if fork else ref).backward(gy)
    assert _close(y.detach(), (ref + x if fork else ref).detach(), 1e-4)
    for a, t in zip(got, (x, w1, b1, w2)):
        assert _close(a, t.grad, 1e-3 if compact else 1e-4)  # (FP16 derivative: 2^-12 relative)


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


@pytest.mark.parametrize("compact", [False, True])
def test_mlp_node_uses_the_gelu_grad_epilogue_where_the_tf32_kernel_does_not_cover(fake, monkeypatch, compact):
    monkeypatch.setattr(ops, "_MLP_COMPACT", "all" if compact else "0")
    monkeypatch.setattr(fake, "hs_mlp_dgrad_gelu_supported", lambda T, Cc, J: 0)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(5, 8, generator=g, requires_grad=True)
    w1 = (torch.randn(32, 8, generator=g) / 3).requires_grad_(True)
    b1 = torch.randn(32, generator=g, requires_grad=True)
    w2 = (torch.randn(8, 32, generator=g) / 6).requires_grad_(True)
    ops._MlpFn.apply(x, w1, b1, w2, 0.0, 0, False).sum().backward()
    assert fake.calls == (["gemm3:6", "gemm3:0", "wgrad", "gemm3:7", "gemm3:0", "wgrad"] if compact else
                          ["gemm3:2", "gemm3:0", "wgrad", "gemm3:3", "gemm3:0", "wgrad"])
    got = [t.grad.clone() for t in (x, w1, b1, w2)]
    for t in (x, w1, b1, w2):
        t.grad = None
    F.linear(F.gelu(F.linear(x, w1, b1)), w2).sum().backward()
    for a, t in zip(got, (x, w1, b1, w2)):
        assert _close(a, t.grad, 1e-3 if compact else 1e-4)


def test_weight_splits_are_cached_until_the_parameter_changes(fake):
    w = torch.nn.Parameter(torch.randn(8, 12))
    a = ops.split_weight(w)
    assert ops.split_weight(w) is a and fake.splits == 1
    at = ops.split_weight(w, transposed=True)
    assert at.shape == (12, 64) and a.shape == (8, 64) and fake.splits == 2
    with torch.no_grad():
        w.mul_(2.0)  # an optimizer step bumps the version counter
    b = ops.split_weight(w)
    assert fake.splits == 3 and b is a  # same buffer (stable address), new contents
    # fused optimizers do not bump the version counter: the global optimizer post-step hook starts a new cache epoch
    opt = torch.optim.Adam([w], lr=0.1, fused=True)
    w.grad = torch.ones_like(w)
    v0 = w._version
    opt.step()
    assert w._version == v0, "torch changed: fused Adam now bumps versions (the hook is then merely redundant)"
    assert ops.split_weight(w) is a and fake.splits == 4
    assert ops.split_weight(w) is a and fake.splits == 4  # cached again until the next step
    b = ops.split_weight(w)
    hi_lo = b.view(8, 1, 64).float()
    assert _close((hi_lo[:, 0, :32] + hi_lo[:, 0, 32:])[:, :12], w.detach(), 1e-4)


def test_bias_gradient_rides_with_the_dgrad_gemm_where_the_wgrad_kernel_cannot_fuse_it(fake):
    fake.wgrad_cover = 1  # weight gradient covered, bias not (K > 224 in the real kernel)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(9, 8, generator=g, requires_grad=True)
    w = torch.randn(12, 8, generator=g, requires_grad=True)
    b = torch.randn(12, generator=g, requires_grad=True)
    ops._LinearFn.apply(x, w, b, False).square().sum().backward()
    assert fake.calls == ["gemm3:0", "gemm3:0+colsum", "wgrad"]
    got = [t.grad.clone() for t in (x, w, b)]
    for t in (x, w, b):
        t.grad = None
    F.linear(x, w, b).square().sum().backward()
    for a, t in zip(got, (x, w, b)):
        assert _close(a, t.grad, 1e-4)
    # no input gradient wanted (first layer): plain reduction
    fake.calls.clear()
    x2 = torch.randn(9, 8, generator=g)
    w.grad = b.grad = None
    ops._LinearFn.apply(x2, w, b, False).sum().backward()
    assert fake.calls == ["gemm3:0", "wgrad"] and _close(b.grad, torch.full((12,), 9.0))


# ---------------------------------------------------------------------------------------------------------------------
# LayerNorm inside the GEMM: autograd plumbing of ops.linear_ln / ops.mlp_ln / ops.ln_linear (hs_gemm3_ln, hs_gemm3_lnin)

@pytest.mark.parametrize("res", [False, True])
@pytest.mark.parametrize("G", [32, 64])
def test_linear_ln_node_matches_torch_autograd(fake, res, G):
    """y = [residual +] LN_G(linear(x)): one fused forward launch, then LayerNorm backward (which also yields the linear's
    bias gradient when G = N), dgrad, wgrad -- all seven gradients against plain torch."""
    g = torch.Generator().manual_seed(4)
    T, K, N = 10, 8, 64
    x = torch.randn(T, K, generator=g, requires_grad=True)
    w = (torch.randn(N, K, generator=g) / 3).requires_grad_(True)
    b = torch.randn(N, generator=g, requires_grad=True)
    gamma = (1 + 0.3 * torch.randn(G, generator=g)).requires_grad_(True)
    beta = (0.2 * torch.randn(G, generator=g)).requires_grad_(True)
    r = torch.randn(T, N, generator=g, requires_grad=True) if res else None
    gy = torch.randn(T, N, generator=g)
    params = [x, w, b, gamma, beta] + ([r] if res else [])
    y = ops._LinearLnFn.apply(x, w, b, gamma, beta, r, 1e-5)
    y.backward(gy)
    # forward; LN backward (takes the linear's bias gradient too when G = N, else the weight-gradient kernel does); dgrad; wgrad
    assert fake.calls == ["gemm3_ln" + ("+aux" if res else "") + "+save", "ln_bwd", "gemm3:0", "wgrad"]
    got = [t.grad.clone() for t in params]
    for t in params:
        t.grad = None
    ref = F.layer_norm(F.linear(x, w, b).view(T, N // G, G), (G,), gamma, beta, 1e-5).view(T, N)
    ref = ref + r if res else ref
    ref.backward(gy)
    assert _close(y.detach(), ref.detach(), 1e-4)
    for a, t in zip(got, params):
        assert _close(a, t.grad, 1e-4)


def test_mlp_ln_node_matches_torch_autograd(fake, monkeypatch):
    """x + LN(fc2(GELU(fc1 x)) + b2): the residual's gradient enters fc1's input-gradient GEMM as its aux operand (mode 1),
    fc2's bias gradient comes out of the LayerNorm-backward pass."""
    monkeypatch.setattr(ops, "_MLP_COMPACT", "0")
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 6, 32, generator=g, requires_grad=True)
    w1 = (torch.randn(64, 32, generator=g) / 6).requires_grad_(True)
    b1 = torch.randn(64, generator=g, requires_grad=True)
    w2 = (torch.randn(32, 64, generator=g) / 8).requires_grad_(True)
    b2 = torch.randn(32, generator=g, requires_grad=True)
    gamma = (1 + 0.3 * torch.randn(32, generator=g)).requires_grad_(True)
    beta = (0.2 * torch.randn(32, generator=g)).requires_grad_(True)
    gy = torch.randn(2, 6, 32, generator=g)
    params = [x, w1, b1, w2, b2, gamma, beta]
    y = ops._MlpFn.apply(x, w1, b1, w2, 0.0, 0, False, b2, gamma, beta, 1e-5)
    y.backward(gy)
    assert fake.calls == ["gemm3:2", "gemm3_ln+aux+save", "ln_bwd", "wgrad", "mlp_dgrad_gelu", "gemm3:1", "wgrad"]
    got = [t.grad.clone() for t in params]
    for t in params:
        t.grad = None
    ref = x + F.layer_norm(F.linear(F.gelu(F.linear(x, w1, b1)), w2, b2), (32,), gamma, beta, 1e-5)
    ref.backward(gy)
    assert _close(y.detach(), ref.detach(), 1e-4)
    for a, t in zip(got, params):
        assert _close(a, t.grad, 1e-4)


def test_ln_linear_node_matches_torch_autograd(fake):
    """linear(LN(x)) (PatchMerging): one forward launch on the raw rows; the backward rebuilds the normalised rows for the
    weight gradient and runs dgrad + LayerNorm backward for x, gamma, beta."""
    g = torch.Generator().manual_seed(6)
    T, K, N = 9, 160, 16
    x = (torch.randn(T, K, generator=g) + 0.5).requires_grad_(True)
    gamma = (1 + 0.3 * torch.randn(K, generator=g)).requires_grad_(True)
    beta = (0.2 * torch.randn(K, generator=g)).requires_grad_(True)
    w = (torch.randn(N, K, generator=g) / 12).requires_grad_(True)
    gy = torch.randn(T, N, generator=g)
    params = [x, gamma, beta, w]
    y = ops._LnLinearFn.apply(x, gamma, beta, w, 1e-5)
    y.backward(gy)
    assert fake.calls == ["gemm3_lnin", "ln_fwd", "wgrad", "gemm3:0", "ln_bwd"]
    got = [t.grad.clone() for t in params]
    for t in params:
        t.grad = None
    ref = F.linear(F.layer_norm(x, (K,), gamma, beta, 1e-5), w)
    ref.backward(gy)
    assert _close(y.detach(), ref.detach(), 1e-4)
    for a, t in zip(got, params):
        assert _close(a, t.grad, 2e-4)


def test_zero_arena_serves_aligned_slices_and_measures_first():
    ar = ops.ZeroArena("cpu")
    prev = ops.use_zero_arena(ar)
    try:
        ar.begin()
        a = ops.zeros((3, 5), "cpu")          # first pass: no buffer yet, plain zeros, but the demand is recorded
        b = ops.zeros(70, "cpu")
        assert ar.buf is None and ar.need == 64 + 128 and float(a.abs().sum() + b.abs().sum()) == 0.0
        ar.begin()                            # allocates what the first pass asked for
        a = ops.zeros((3, 5), "cpu")
        b = ops.zeros(70, "cpu")
        assert a.data_ptr() == ar.buf.data_ptr() and b.data_ptr() == ar.buf.data_ptr() + 64 * 4
        a.add_(1.0)
        b.add_(2.0)
        c = ops.zeros(4, "cpu")               # beyond the measured demand: falls back to torch.zeros
        assert float(c.sum()) == 0.0 and c.data_ptr() != ar.buf.data_ptr()
        ar.begin()                            # one memset clears every slice
        assert float(ar.buf.abs().sum()) == 0.0
    finally:
        ops.use_zero_arena(prev)
    assert ops.zeros(3, "cpu").shape == (3,)


def test_refresh_weight_splits_marks_the_cache_current(fake, monkeypatch):
    """One batched launch covers every live cached operand; afterwards split_weight returns the cached buffers without
    launching, also for parameters that were changed without a version bump."""
    w1, w2 = torch.nn.Parameter(torch.randn(8, 12)), torch.nn.Parameter(torch.randn(16, 8))
    a, b, c = ops.split_weight(w1), ops.split_weight(w1, transposed=True), ops.split_weight(w2)
    assert fake.splits == 3
    w1.data.mul_(2.0)
    ops._on_optimizer_step()                  # what a torch optimizer's post-step hook does: a new epoch
    monkeypatch.setattr(torch.Tensor, "to", lambda self, *a_, **k: self, raising=False)
    n = ops.refresh_weight_splits("cpu")
    assert n == 3 and fake.calls[-1] == "split_batch:3"
    assert ops.split_weight(w1) is a and ops.split_weight(w1, transposed=True) is b and ops.split_weight(w2) is c
    assert fake.splits == 3                   # no per-matrix launches after the batched refresh
    del w2
    import gc

    gc.collect()
    assert ops.refresh_weight_splits("cpu") == 2
