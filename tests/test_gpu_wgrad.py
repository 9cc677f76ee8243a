"""GPU parity of the tcgen05 weight-gradient kernel (dW = dY^T X, TF32 operands, fp32 accumulation) through the C-ABI
against the fp32 torch product.  Tolerance 2e-3 relative L2 (TF32 operands); exactness checks use TF32-representable data."""
import pytest
import torch

from tests.util import rel_err

pytestmark = pytest.mark.gpu
TOL = 2e-3


def _wgrad(dy, x):
    import ctypes as C

    from heal_swin_b200 import _lib
    from heal_swin_b200._lib import check, current_stream, lib, ptr

    T, N = dy.shape
    K = x.shape[1]
    assert lib.hs_linear_wgrad_supported(T, N, K)
    dw = torch.zeros(N, K, device=dy.device)
    check(lib.hs_linear_wgrad(ptr(dy), ptr(x), ptr(dw), None, T, N, K, 0, current_stream()))
    return dw


@pytest.mark.parametrize("T,N,K", [(8192, 288, 96), (8192, 96, 96), (8192, 384, 96), (8192, 96, 384), (5000, 576, 192),
                                   (4096 + 37, 192, 768), (8192, 32, 64), (20000, 100, 64), (8192, 1152, 256),
                                   # smaller feature dimension in (256, 512]: 32-token stages, two MMAs per K step
                                   (8192, 1152, 384), (8192, 384, 384), (6000, 384, 1536), (8192, 640, 512),
                                   (8192, 288, 320),
                                   # smaller feature dimension in (512, 1024]: two launches over its column halves (stage 3)
                                   (24576, 2304, 768), (8192, 768, 768), (6000, 768, 3072), (8192, 1536, 1024),
                                   # smaller feature dimension not a multiple of 32 (patch embedding: 4 * 4 * 3 = 48 inputs):
                                   # zero-filled last slab
                                   (8192, 96, 48), (8192, 48, 96), (5000, 128, 36), (8192, 1536, 200), (8192, 96, 12), (8192, 12, 96),
                                   # CTA pairs (>= 6 slabs of the smaller operand, even number of 128-row blocks)
                                   (16384, 768, 192), (16384, 192, 768), (8192, 1536, 384), (6000, 3072, 768)])
def test_wgrad_matches_fp32_product(T, N, K):
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(T + N + K)
    dy = torch.randn(T, N, generator=g).to(dev)
    x = torch.randn(T, K, generator=g).to(dev)
    got = _wgrad(dy, x)
    want = (dy.double().t() @ x.double()).float()
    assert rel_err(got.cpu(), want.cpu()) < TOL


@pytest.mark.parametrize("N,K", [(288, 96), (96, 48), (96, 12)])  # (96, 12): the patch embedding, ragged slab + fused bias gradient
def test_wgrad_exact_on_tf32_representable_data_and_accumulates(N, K):
    """Small integers are exact in TF32 and the sums stay below 2^24: the result must be bit-exact once the truncation
    compensation is switched off; a second call accumulates (+=)."""
    import ctypes as C

    from heal_swin_b200 import _lib
    from heal_swin_b200._lib import check, current_stream, lib, ptr

    dev = torch.device("cuda:0")
    T = 16384
    g = torch.Generator().manual_seed(1)
    dy = torch.randint(-4, 5, (T, N), generator=g).float().to(dev)
    x = torch.randint(-4, 5, (T, K), generator=g).float().to(dev)
    dw = torch.zeros(N, K, device=dev)
    db = torch.zeros(N, device=dev)
    assert lib.hs_linear_wgrad_supported(T, N, K) == 2
    check(lib.hs_linear_wgrad(ptr(dy), ptr(x), ptr(dw), ptr(db), T, N, K, _lib.ATTN_NO_TRUNC_COMP, current_stream()))
    want = (dy.double().t() @ x.double()).float()
    assert torch.equal(dw, want)
    assert torch.equal(db, dy.double().sum(0).float())  # the fused bias gradient (column sums of dy)
    check(lib.hs_linear_wgrad(ptr(dy), ptr(x), ptr(dw), ptr(db), T, N, K, _lib.ATTN_NO_TRUNC_COMP, current_stream()))
    assert torch.equal(dw, 2 * want)


def test_linear_autograd_uses_the_kernels_and_matches():
    from heal_swin_b200 import ops

    dev = torch.device("cuda:0")
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        g = torch.Generator().manual_seed(3)
        x = torch.randn(4, 4096, 96, generator=g).to(dev).requires_grad_(True)
        lin = torch.nn.Linear(96, 288).to(dev)
        wgt = torch.randn(4, 4096, 288, generator=g).to(dev)
        ops.STATS.reset()
        y = ops.linear(x, lin.weight, lin.bias)
        (y * wgt).sum().backward()
        n = ops.STATS.by_name  # forward + input gradient on the bf16x3 GEMM (one weight split each), the wgrad kernel
        assert n == {"weight_split": 2, "gemm3": 2, "linear_wgrad": 1}, n
        got = (lin.weight.grad.clone(), lin.bias.grad.clone(), x.grad.clone())
        lin.weight.grad = lin.bias.grad = x.grad = None
        torch.backends.cuda.matmul.allow_tf32 = False
        y2 = torch.nn.functional.linear(x, lin.weight, lin.bias)
        (y2 * wgt).sum().backward()
        for a, b in zip(got, (lin.weight.grad, lin.bias.grad, x.grad)):
            assert rel_err(a.cpu(), b.cpu()) < TOL
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def test_forked_linear_folds_the_shortcut_gradient_into_the_dgrad():
    """ops.linear(..., fork=True) returns (y, shortcut): gradients must equal those of using x twice."""
    from heal_swin_b200 import ops

    dev = torch.device("cuda:0")
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        g = torch.Generator().manual_seed(4)
        x = torch.randn(2, 4096, 96, generator=g).to(dev).requires_grad_(True)
        lin = torch.nn.Linear(96, 288).to(dev)
        wy = torch.randn(2, 4096, 288, generator=g).to(dev)
        ws = torch.randn(2, 4096, 96, generator=g).to(dev)
        y, shortcut = ops.linear(x, lin.weight, lin.bias, fork=True)
        assert shortcut.data_ptr() == x.data_ptr() and shortcut.grad_fn is y.grad_fn
        ((y * wy).sum() + (shortcut * ws).sum()).backward()
        got = (lin.weight.grad.clone(), lin.bias.grad.clone(), x.grad.clone())
        lin.weight.grad = lin.bias.grad = x.grad = None
        torch.backends.cuda.matmul.allow_tf32 = False
        ((torch.nn.functional.linear(x, lin.weight, lin.bias) * wy).sum() + (x * ws).sum()).backward()
        for a, b in zip(got, (lin.weight.grad, lin.bias.grad, x.grad)):
            assert rel_err(a.cpu(), b.cpu()) < TOL
        # only one of the two outputs used
        x.grad = None
        y, shortcut = ops.linear(x, lin.weight, lin.bias, fork=True)
        (shortcut * ws).sum().backward()
        assert torch.equal(x.grad, ws)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def test_wgrad_full_size_linearity():
    """BASELINE stage-0 token count (8 x 196608): dW is linear in dY (size-independent property)."""
    dev = torch.device("cuda:0")
    T, N, K = 8 * 196608, 288, 96
    x = torch.randn(T, K, device=dev)
    d1 = torch.randn(T, N, device=dev)
    d2 = torch.randn(T, N, device=dev)
    w1, w2, w12 = _wgrad(d1, x), _wgrad(d2, x), _wgrad(d1 + 0.5 * d2, x)
    assert rel_err((w1 + 0.5 * w2).cpu(), w12.cpu()) < TOL
