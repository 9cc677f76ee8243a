"""GPU parity of the LayerNorm (+ fused residual) kernels through the C-ABI against torch's fp32 layer_norm
(floating-point kernel: plain PyTorch fp32 reference) and against the reference-generated op fixtures.

Tolerance: 1e-5 relative L2 forward, 1e-4 backward (fp32 everywhere; only the reduction order differs).
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.util import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("C", [96, 192, 384, 768, 1536, 48, 24, 64, 16, 8, 2, 100])
@pytest.mark.parametrize("rows,with_res", [(1001, True), (64, False), (3, True)])
def test_layernorm_fwd_bwd_vs_torch(C, rows, with_res):
    from heal_swin_b200 import ops

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(C * 7 + rows)
    x = (torch.randn(rows, C, generator=g) * 2 + 0.5).to(dev).requires_grad_(True)
    res = torch.randn(rows, C, generator=g).to(dev).requires_grad_(True) if with_res else None
    norm = torch.nn.LayerNorm(C).to(dev)
    with torch.no_grad():
        norm.weight.copy_(1 + 0.3 * torch.randn(C, generator=g))
        norm.bias.copy_(0.2 * torch.randn(C, generator=g))
    wgt = torch.randn(rows, C, generator=g).to(dev)

    y = ops.layer_norm(x, norm, residual=res)
    (y * wgt).sum().backward()
    got = [y.detach(), x.grad.clone(), norm.weight.grad.clone(), norm.bias.grad.clone()]
    if with_res:
        got.append(res.grad.clone())

    x2 = x.detach().clone().requires_grad_(True)
    r2 = res.detach().clone().requires_grad_(True) if with_res else None
    norm.weight.grad = norm.bias.grad = None
    y2 = F.layer_norm(x2, (C,), norm.weight, norm.bias, norm.eps)
    if with_res:
        y2 = r2 + y2
    (y2 * wgt).sum().backward()
    want = [y2.detach(), x2.grad, norm.weight.grad, norm.bias.grad] + ([r2.grad] if with_res else [])
    tols = [1e-5, 1e-4, 1e-4, 1e-4, 1e-6]
    for i, (a, b, tol) in enumerate(zip(got, want, tols)):
        if i == 1 and C == 2:
            # C = 2: the normalised row is (+1, -1) whatever x is, so dx is pure cancellation amplified by rstd (up to
            # 1/sqrt(eps) when the two channels nearly coincide): ill-conditioned in any fp32 implementation, not compared
            assert torch.isfinite(a).all()
            continue
        assert rel_err(a.cpu(), b.cpu()) < tol


def test_layernorm_keeps_leading_shape_and_3d_input():
    from heal_swin_b200 import ops

    dev = torch.device("cuda:0")
    x = torch.randn(2, 50, 96, device=dev)
    norm = torch.nn.LayerNorm(96).to(dev)
    y = ops.layer_norm(x, norm)
    assert y.shape == x.shape
    assert rel_err(y.cpu(), F.layer_norm(x, (96,), norm.weight, norm.bias, norm.eps).detach().cpu()) < 1e-5


def test_layernorm_full_size_properties():
    """BASELINE configs[1] stage-0 size (8 x 196608 rows of 96): rows are normalised independently, so (a) every output
    row has zero mean / unit variance with the identity affine, (b) permuting rows permutes the output."""
    from heal_swin_b200 import ops

    dev = torch.device("cuda:0")
    rows, C = 8 * 196608, 96
    x = torch.randn(rows, C, device=dev) * 3 + 1
    norm = torch.nn.LayerNorm(C).to(dev)
    y = ops.layer_norm(x, norm)
    assert float(y.mean(1).abs().max()) < 1e-5
    assert float((y.var(1, unbiased=False) - 1).abs().max()) < 1e-3
    perm = torch.randperm(rows, device=dev)
    assert torch.equal(ops.layer_norm(x[perm], norm), y[perm])


@pytest.mark.parametrize("C", [96, 384, 768, 1536, 16, 100])
def test_layernorm_with_pre_bias_and_residual(C):
    """y = res + LN(x + pre_bias): the fused form used after proj / fc2 in the v2 placement; d(pre_bias) = colsum(dx)."""
    from heal_swin_b200 import ops

    dev = torch.device("cuda:0")
    rows = 777
    g = torch.Generator().manual_seed(C)
    x = torch.randn(rows, C, generator=g).to(dev).requires_grad_(True)
    res = torch.randn(rows, C, generator=g).to(dev).requires_grad_(True)
    pb = torch.randn(C, generator=g).to(dev).requires_grad_(True)
    norm = torch.nn.LayerNorm(C).to(dev)
    with torch.no_grad():
        norm.weight.copy_(1 + 0.3 * torch.randn(C, generator=g))
        norm.bias.copy_(0.2 * torch.randn(C, generator=g))
    wgt = torch.randn(rows, C, generator=g).to(dev)
    y = ops.layer_norm(x, norm, residual=res, pre_bias=pb)
    (y * wgt).sum().backward()
    got = [y.detach().clone(), x.grad.clone(), pb.grad.clone(), norm.weight.grad.clone(), norm.bias.grad.clone()]
    for t in (x, res, pb, norm.weight, norm.bias):
        t.grad = None
    y2 = res + F.layer_norm(x + pb, (C,), norm.weight, norm.bias, norm.eps)
    (y2 * wgt).sum().backward()
    want = [y2.detach(), x.grad, pb.grad, norm.weight.grad, norm.bias.grad]
    for a, b, tol in zip(got, want, [1e-5, 1e-4, 1e-4, 1e-4, 1e-4]):
        assert rel_err(a.cpu(), b.cpu()) < tol


@pytest.mark.parametrize("C", [384, 768, 1536, 3072, 64, 8])
@pytest.mark.parametrize("rows", [1000, 7])
def test_bias_gelu_fwd_bwd_vs_torch(C, rows):
    from heal_swin_b200 import ops

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(C + rows)
    z = (torch.randn(rows, C, generator=g) * 2).to(dev).requires_grad_(True)
    b = torch.randn(C, generator=g).to(dev).requires_grad_(True)
    wgt = torch.randn(rows, C, generator=g).to(dev)
    h = ops.bias_gelu(z, b)
    (h * wgt).sum().backward()
    got = [h.detach().clone(), z.grad.clone(), b.grad.clone()]
    z.grad = b.grad = None
    h2 = F.gelu(z + b)
    (h2 * wgt).sum().backward()
    for a, w, tol in zip(got, [h2.detach(), z.grad, b.grad], [1e-5, 1e-5, 1e-4]):
        assert rel_err(a.cpu(), w.cpu()) < tol


# ---------------------------------------------------------------------------------------------- fused dropout / drop-path
M32 = 0xFFFFFFFF


def _mix32(x):
    x = x.astype(np.uint64)
    x ^= x >> 16
    x = (x * 0x7FEB352D) & M32
    x ^= x >> 15
    x = (x * 0x846CA68B) & M32
    x ^= x >> 16
    return x


def host_elem_mask(seed, rows, C, p):
    """(rows, C) float32 multiplier of the element-wise dropout the kernels apply (csrc/hs_common.h: drop_row_key /
    drop_keep_elem): 0 where dropped, 1/(1-p) where kept."""
    thresh = min(int(float(np.float32(p)) * 4294967296.0), M32)
    r = np.arange(rows, dtype=np.uint64)
    key = _mix32((np.uint64(seed & M32) ^ _mix32(((seed >> 32) + r) & M32) ^ (r >> 32)) & M32)
    e = (np.arange(C, dtype=np.uint64) * 0x9E3779B9) & M32
    keep = _mix32((key[:, None] ^ e[None, :]) & M32) >= thresh
    return torch.from_numpy(keep.astype(np.float32) * (np.float32(1.0) / (np.float32(1.0) - np.float32(p))))


@pytest.mark.parametrize("C", [96, 384, 100])
def test_layernorm_fused_dropout_droppath_matches_torch_with_the_same_mask(C):
    from heal_swin_b200 import ops

    dev = torch.device("cuda:0")
    B, n, p, seed = 4, 150, 0.2, 0x2BCDEF0123456789
    rows = B * n
    g = torch.Generator().manual_seed(C)
    x = torch.randn(B, n, C, generator=g).to(dev).requires_grad_(True)
    res = torch.randn(B, n, C, generator=g).to(dev).requires_grad_(True)
    pb = torch.randn(C, generator=g).to(dev).requires_grad_(True)
    scale = torch.tensor([0.0, 1.25, 1.25, 0.0]).to(dev)  # a stochastic-depth draw with keep = 0.8
    norm = torch.nn.LayerNorm(C).to(dev)
    with torch.no_grad():
        norm.weight.copy_(1 + 0.3 * torch.randn(C, generator=g))
        norm.bias.copy_(0.2 * torch.randn(C, generator=g))
    wgt = torch.randn(B, n, C, generator=g).to(dev)
    y = ops.layer_norm(x, norm, residual=res, pre_bias=pb, row_scale=scale, in_drop=p, seed=seed)
    (y * wgt).sum().backward()
    got = [y.detach().clone(), x.grad.clone(), pb.grad.clone(), norm.weight.grad.clone(), norm.bias.grad.clone(), res.grad.clone()]
    for t in (x, res, pb, norm.weight, norm.bias):
        t.grad = None
    mult = host_elem_mask(seed, rows, C, p).view(B, n, C).to(dev)
    y2 = res + scale.view(B, 1, 1) * F.layer_norm((x + pb) * mult, (C,), norm.weight, norm.bias, norm.eps)
    (y2 * wgt).sum().backward()
    want = [y2.detach(), x.grad, pb.grad, norm.weight.grad, norm.bias.grad, res.grad]
    for a, b, tol in zip(got, want, [1e-5, 1e-4, 1e-4, 1e-4, 1e-4, 1e-6]):
        assert rel_err(a.cpu(), b.cpu()) < tol


@pytest.mark.parametrize("C", [384, 64])
def test_bias_gelu_fused_dropout_matches_torch_with_the_same_mask(C):
    from heal_swin_b200 import ops

    dev = torch.device("cuda:0")
    rows, p, seed = 999, 0.1, 77
    g = torch.Generator().manual_seed(C)
    z = (torch.randn(rows, C, generator=g) * 2).to(dev).requires_grad_(True)
    b = torch.randn(C, generator=g).to(dev).requires_grad_(True)
    wgt = torch.randn(rows, C, generator=g).to(dev)
    h = ops.bias_gelu(z, b, drop=p, seed=seed)
    (h * wgt).sum().backward()
    got = [h.detach().clone(), z.grad.clone(), b.grad.clone()]
    z.grad = b.grad = None
    mult = host_elem_mask(seed, rows, C, p).to(dev)
    h2 = F.gelu(z + b) * mult
    (h2 * wgt).sum().backward()
    kept = float((mult > 0).float().mean())
    assert abs(kept - (1 - p)) < 0.01
    for a, w, tol in zip(got, [h2.detach(), z.grad, b.grad], [1e-5, 1e-5, 1e-4]):
        assert rel_err(a.cpu(), w.cpu()) < tol
