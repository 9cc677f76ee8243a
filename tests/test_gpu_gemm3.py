"""GPU parity of the bf16x3 tensor-core GEMM (csrc/hs_gemm3_tc.cu) through the C-ABI against an fp64 product.
Tolerance 2e-5 relative L2: every operand is split into two bf16 terms (residual <= 2^-17) and the lo*lo term is dropped,
so each product is good to ~2^-16 -- fp32-class, unlike TF32 (2e-4 on the same data).  Covers all four epilogue modes,
ragged token counts (TMA clips the last 128-row tile), feature dimensions that are not multiples of 32 (zero-filled
chunk tails, clipped stores), resident and streamed weight chunks, several column chunks, and the dropout masks."""
import ctypes as C
import math

import pytest
import torch

from tests.util import rel_err

pytestmark = pytest.mark.gpu
TOL = 2e-5


def _split(w, transposed=False, prec=0):
    from heal_swin_b200._lib import check, current_stream, lib, ptr

    w = w.contiguous()
    n, k = w.shape
    rows, cols = (k, n) if transposed else (n, k)
    out = torch.empty((rows, 2 * ((cols + 31) // 32 * 32)), device=w.device, dtype=torch.bfloat16)
    check(lib.hs_weight_split(ptr(w), rows, cols, k, 1 if transposed else 0, 1 if prec == 1 else 0, ptr(out), current_stream()))
    return out


def _gemm3(a, ws, N, bias=None, aux=None, mode=0, drop=0.0, seed=0, colsum=None, prec=0):
    from heal_swin_b200._lib import check, current_stream, lib, ptr

    T, K = a.shape
    d = torch.full((T, N), float("nan"), device=a.device)
    d2 = torch.full((T, N), float("nan"), device=a.device) if mode == 2 else None
    check(lib.hs_gemm3(ptr(a), ptr(ws), ptr(bias), ptr(aux), ptr(d), ptr(d2), ptr(colsum), T, N, K, mode, prec, C.c_float(drop),
                       C.c_uint64(seed), current_stream()))
    return (d, d2) if mode == 2 else d


def _data(T, N, K, seed=0):
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(seed + T + 3 * N + 7 * K)
    a = torch.randn(T, K, generator=g).to(dev)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    return a, w, b


def test_weight_split_layout_and_accuracy():
    """hi + lo reproduces the weight to 2^-16; layout (rows, chunks, [hi 32 | lo 32]); transposed form; zero padding."""
    dev = torch.device("cuda:0")
    w = torch.randn(40, 72, generator=torch.Generator().manual_seed(0)).to(dev)
    for transposed in (False, True):
        m = w.t() if transposed else w
        rows, cols = m.shape
        out = _split(w, transposed).view(rows, -1, 64).float()
        assert out.shape[1] == (cols + 31) // 32
        hi, lo = out[:, :, :32].reshape(rows, -1), out[:, :, 32:].reshape(rows, -1)
        assert torch.equal(hi[:, :cols], m.bfloat16().float())  # hi is the round-to-nearest bf16
        assert float((hi + lo)[:, :cols].sub(m).abs().max() / m.abs().max()) < 2.0 ** -16
        assert float(hi[:, cols:].abs().max()) == 0.0 and float(lo[:, cols:].abs().max()) == 0.0


@pytest.mark.parametrize("T,N,K", [
    (128, 32, 32), (300, 96, 96), (1000, 288, 64),        # ragged tokens, two column chunks (160 + 128)
    (4096, 384, 96), (4096, 96, 384),                     # resident W / streamed W (12 K slices)
    (2048, 1152, 384), (1024, 768, 3072),                 # 5 column chunks; 96 K chunks
    (777, 144, 48), (512, 96, 12), (640, 200, 100),       # feature dims that are not multiples of 32
])
def test_plain_and_add_match_fp64(T, N, K):
    a, w, b = _data(T, N, K)
    want = a.double() @ w.double().t()
    got = _gemm3(a, _split(w), N, b)
    assert rel_err(got.cpu(), (want + b.double()).cpu()) < TOL
    got = _gemm3(a, _split(w), N)
    assert rel_err(got.cpu(), want.cpu()) < TOL
    aux = torch.randn(T, N, device=a.device)
    got = _gemm3(a, _split(w), N, b, aux, mode=1)
    assert rel_err(got.cpu(), (want + b.double() + aux.double()).cpu()) < TOL
    # input-gradient form: dy @ w through the transposed split
    dy = torch.randn(T, N, device=a.device)
    got = _gemm3(dy, _split(w, transposed=True), K)
    assert rel_err(got.cpu(), (dy.double() @ w.double()).cpu()) < TOL


@pytest.mark.parametrize("T,N,K", [(300, 96, 96), (4096, 384, 96), (2048, 768, 192), (1000, 1536, 384), (600, 192, 48)])
def test_gelu_epilogues_match_fp64(T, N, K):
    a, w, b = _data(T, N, K, seed=1)
    z_want = a.double() @ w.double().t()
    z, h = _gemm3(a, _split(w), N, b, mode=2)
    assert rel_err(z.cpu(), z_want.cpu()) < TOL  # z is the bias-free fc1 output, as hs_bias_gelu_fwd expects
    assert rel_err(h.cpu(), torch.nn.functional.gelu(z_want + b.double()).cpu()) < TOL
    zz = torch.randn(T, N, device=a.device)
    got = _gemm3(a, _split(w), N, b, zz, mode=3)
    u = (zz.double() + b.double()).requires_grad_(True)
    torch.nn.functional.gelu(u).sum().backward()
    assert rel_err(got.cpu(), (z_want * u.grad).cpu()) < TOL


@pytest.mark.parametrize("T,N,K,mode", [(1000, 96, 288, 0), (4096, 384, 1152, 1), (777, 96, 100, 0), (2048, 768, 3072, 1)])
def test_column_sums_ride_along(T, N, K, mode):
    """colsum += column sums of the A operand (the bias gradient of a linear whose output gradient is A), accumulated once
    although several column-chunk CTAs convert the same A tile; rows beyond T and columns beyond K contribute nothing."""
    a, w, b = _data(T, N, K, seed=3)
    aux = torch.randn(T, N, device=a.device) if mode == 1 else None
    cs = torch.full((K,), 2.0, device=a.device)
    got = _gemm3(a, _split(w), N, None, aux, mode=mode, colsum=cs)
    want = a.double() @ w.double().t() + (aux.double() if mode == 1 else 0)
    assert rel_err(got.cpu(), want.cpu()) < TOL
    assert rel_err(cs.cpu(), (a.double().sum(0) + 2.0).cpu()) < 1e-5


def test_dropout_masks_match_the_bias_gelu_kernels():
    """Modes 2 / 3 with drop > 0 use the same counter-based mask as hs_bias_gelu_fwd / bwd (pure function of seed, row,
    column): the fused epilogues must reproduce the two-launch results element for element (up to GEMM rounding)."""
    from heal_swin_b200._lib import check, current_stream, lib, ptr

    T, N, K = 1500, 384, 96
    a, w, b = _data(T, N, K, seed=2)
    drop, seed = 0.25, 0x1234_5678_9ABC
    z, h = _gemm3(a, _split(w), N, b, mode=2, drop=drop, seed=seed)
    h_ref = torch.empty_like(z)
    check(lib.hs_bias_gelu_fwd(ptr(z), ptr(b), C.c_float(drop), C.c_uint64(seed), ptr(h_ref), T, N, current_stream()))
    assert torch.equal(h == 0, h_ref == 0) and 0.2 < float((h == 0).float().mean()) < 0.3
    assert rel_err(h.cpu(), h_ref.cpu()) < 1e-6
    dh = a.double() @ w.double().t()  # any (T, N) "hidden gradient": reuse the product
    got = _gemm3(a, _split(w), N, b, z, mode=3, drop=drop, seed=seed)
    dz_ref = torch.empty_like(z)
    check(lib.hs_bias_gelu_bwd(ptr(dh.float().contiguous()), ptr(z), ptr(b), C.c_float(drop), C.c_uint64(seed), ptr(dz_ref),
                               None, T, N, current_stream()))
    assert torch.equal(got == 0, dz_ref == 0)
    assert rel_err(got.cpu(), dz_ref.cpu()) < TOL


def test_linearity_and_exactness_on_bf16_data_at_full_size():
    """BASELINE configs[1] stage-0 qkv shape (1.57 M tokens): operands that are exactly representable in bf16 with small
    integer values give exact fp32 sums -- the result must be bit-exact, on every tile of the full-size launch."""
    dev = torch.device("cuda:0")
    T, N, K = 8 * 196608, 288, 96
    g = torch.Generator(device=dev).manual_seed(5)
    a = torch.randint(-8, 9, (T, K), generator=g, device=dev).float()
    w = torch.randint(-8, 9, (N, K), generator=g, device=dev).float()
    got = _gemm3(a, _split(w), N)
    torch.backends.cuda.matmul.allow_tf32 = False
    want = a @ w.t()  # exact: |sum| <= 96 * 64 < 2^24
    assert torch.equal(got, want)


def test_cat_linear_matches_the_concatenated_linear():
    """ops.cat_linear (skip connections of the decoder, swin_hp_transformer.py:772-775): two accumulating GEMMs over the
    column blocks of the weight == F.linear(cat(x1, x2)), forward and all gradients."""
    from heal_swin_b200 import ops

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(11)
    for T, C in ((4096, 96), (1000, 384)):
        x1 = torch.randn(2, T // 2, C, generator=g).to(dev).requires_grad_(True)
        x2 = torch.randn(2, T // 2, C, generator=g).to(dev).requires_grad_(True)
        lin = torch.nn.Linear(2 * C, C).to(dev)
        gy = torch.randn(2, T // 2, C, generator=g).to(dev)
        ops.STATS.reset()
        y = ops.cat_linear(x1, x2, lin.weight, lin.bias)
        y.backward(gy)
        assert ops.STATS.by_name.get("gemm3") == 4  # two forward, two input-gradient GEMMs; no concatenation
        got = [y.detach().clone()] + [t.grad.clone() for t in (x1, x2, lin.weight, lin.bias)]
        for t in (x1, x2, lin.weight, lin.bias):
            t.grad = None
        torch.backends.cuda.matmul.allow_tf32 = False
        ref = torch.nn.functional.linear(torch.cat([x1, x2], -1), lin.weight, lin.bias)
        ref.backward(gy)
        want = [ref.detach()] + [t.grad for t in (x1, x2, lin.weight, lin.bias)]
        # the weight gradient is the TF32 kernel where T >= 4096; the input gradients are TF32 where bf16x3 would make them
        # tensor-bound (C = 384: ops._dgrad_prec)
        tols = [2e-5, 1e-3, 1e-3, 2e-3, 2e-5]
        for a, b, tol in zip(got, want, tols):
            assert rel_err(a.cpu(), b.cpu()) < tol


@pytest.mark.parametrize("T,N,K", [(1000, 384, 1152), (4096, 1536, 384), (512, 96, 100), (2048, 768, 3072)])
def test_single_mma_precisions(T, N, K):
    """HS_GEMM_TF32 (one TF32 MMA, A read as fp32 without conversion, weight operand = fp32 rounded to TF32) and
    HS_GEMM_BF16 (one bf16 MMA on the hi terms).  Tolerances: TF32 truncates A to 10 mantissa bits (relative 2^-11 on
    average, biased low) -> 1e-3; bf16 operands (2^-9 each) -> 6e-3.  Column sums and the add epilogue ride along."""
    a, w, b = _data(T, N, K, seed=4)
    want = a.double() @ w.double().t() + b.double()
    aux = torch.randn(T, N, device=a.device)
    cs = torch.zeros(K, device=a.device)
    got = _gemm3(a, _split(w, prec=1), N, b, aux, mode=1, colsum=cs, prec=1)
    assert rel_err(got.cpu(), (want + aux.double()).cpu()) < 1e-3
    assert rel_err(cs.cpu(), a.double().sum(0).cpu()) < 1e-5
    got = _gemm3(a, _split(w, prec=1), N, b, prec=1)
    e_tf32 = rel_err(got.cpu(), want.cpu())
    got = _gemm3(a, _split(w), N, b, prec=2)
    e_bf16 = rel_err(got.cpu(), want.cpu())
    got = _gemm3(a, _split(w), N, b)
    e_x3 = rel_err(got.cpu(), want.cpu())
    assert e_x3 < TOL < e_tf32 < 1e-3 < e_bf16 < 6e-3, (e_x3, e_tf32, e_bf16)


@pytest.mark.parametrize("T,N,K,prec", [
    (8 * 196608, 96, 288, 0),    # stage-0 qkv input gradient: 9 K chunks per tile, resident W (the shape that exposed the
                                 # converter-team parity hazard: it needs a long persistent loop to show)
    (8 * 196608, 96, 384, 0),    # stage-0 fc2: streamed W
    (8 * 49152, 768, 192, 0),    # stage-1 fc1: four column chunks
    (8 * 12288, 1536, 384, 0),   # stage-2 fc1: tensor-bound, A split in place in shared memory, 256-column stages
    (8 * 12288, 384, 1536, 1),   # stage-2 fc1 input gradient: single TF32 MMA
    (8 * 3072, 768, 3072, 0),    # stage-3 fc2
    (8 * 12288, 384, 1152, 2),   # bf16-operand mode
])
def test_full_size_launches_are_exact_on_small_integer_data(T, N, K, prec):
    """BASELINE configs[1] shapes at their full token counts (long persistent loops, every pipeline barrier cycling
    thousands of times): on small-integer data every product and every partial sum is exact in all three precisions, so
    the result must be BIT-EXACT on every tile -- any protocol slip (a stale or half-converted chunk) shows."""
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(T % 1000 + N + K)
    a = torch.randint(-4, 5, (T, K), generator=g, device=dev).float()
    w = torch.randint(-4, 5, (N, K), generator=g, device=dev).float()
    for _ in range(3):
        got = _gemm3(a, _split(w, prec=prec), N, prec=prec)
    torch.backends.cuda.matmul.allow_tf32 = False
    want = a @ w.t()  # exact: |sum| <= 3072 * 16 < 2^24
    assert torch.equal(got, want)


@pytest.mark.parametrize("T", [256, 384, 300, 129, 1000, 2048 + 17])
@pytest.mark.parametrize("N,K", [(1152, 384), (384, 1536), (200, 512)])
def test_cta_pair_regime_small_and_ragged_token_counts(T, N, K):
    """The tensor-bound regime (streamed weights, K >= 256) runs as CTA pairs (tcgen05.mma.cta_group::2, M = 256): two
    neighbouring 128-token tiles per cluster, each CTA staging half of every weight slice.  Token counts that leave the
    second tile of the last pair partly or wholly outside the tensor, an odd number of tiles, a ragged last column chunk
    (N = 200), both operand precisions, every epilogue mode and the column sums."""
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(T + N + K)
    a = torch.randn(T, K, generator=g).to(dev)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    aux = torch.randn(T, N, generator=g).to(dev)
    ref = a.double() @ w.double().t()
    for prec, tol in ((0, TOL), (1, 2e-3)):
        ws = _split(w, prec=prec)
        cs = torch.zeros(K, device=dev)
        d0 = _gemm3(a, ws, N, bias, prec=prec, colsum=cs)
        assert rel_err(d0.cpu(), (ref + bias.double()).cpu()) < tol
        assert rel_err(cs.cpu(), a.double().sum(0).cpu()) < 1e-5
        d1 = _gemm3(a, ws, N, None, aux, mode=1, prec=prec)
        assert rel_err(d1.cpu(), (ref + aux.double()).cpu()) < tol
    ws = _split(w)
    z, h = _gemm3(a, ws, N, bias, mode=2)
    assert rel_err(z.cpu(), ref.cpu()) < TOL
    assert rel_err(h.cpu(), torch.nn.functional.gelu(ref + bias.double()).cpu()) < TOL
    gg = _gemm3(a, ws, N, bias, aux, mode=3)
    u = (aux.double() + bias.double()).requires_grad_(True)
    torch.nn.functional.gelu(u).sum().backward()
    assert rel_err(gg.cpu(), (ref * u.grad).cpu()) < TOL


# ---------------------------------------------------------------------------------------------------------------------
# LayerNorm in the GEMM's epilogue (hs_gemm3_ln): PatchExpand's Linear -> view -> LayerNorm and the
# `shortcut + norm(branch)` tail of a v2 block (swin_hp_transformer.py:420-430, 333-338)

def _gemm3_ln(a, ws, N, G, bias, gamma, beta, aux=None, save=True, eps=1e-5):
    from heal_swin_b200._lib import check, current_stream, lib, ptr

    T, K = a.shape
    y = torch.full((T, N), float("nan"), device=a.device)
    pre = torch.full((T, N), float("nan"), device=a.device) if save else None
    mean = torch.full((T * (N // G),), float("nan"), device=a.device) if save else None
    rstd = torch.full((T * (N // G),), float("nan"), device=a.device) if save else None
    check(lib.hs_gemm3_ln(ptr(a), ptr(ws), ptr(bias), ptr(gamma), ptr(beta), ptr(aux), ptr(pre), ptr(y), ptr(mean), ptr(rstd),
                          T, N, K, G, C.c_float(eps), 0, current_stream()))
    return y, pre, mean, rstd


LN_TOL = 2e-5  # the products are good to 2^-16; the statistics are exact fp32 (pivoted one-pass sums merged pairwise)


@pytest.mark.parametrize("T,N,K,G", [
    (300, 96, 96, 96),        # block tail at stage 0 (proj): one LN group, three slabs over two epilogue groups, ragged tokens
    (4096, 96, 384, 96),      # block tail at stage 0 (fc2), resident W with 12 K slices
    (1000, 192, 192, 192),    # stage 1: one group of six slabs
    (2048, 192, 768, 192),    # stage 1 fc2: streamed W
    (1024, 384, 192, 96),     # PatchExpand 1 -> 0: four groups, two per column chunk
    (640, 768, 384, 192),     # PatchExpand 2 -> 1: four chunks of one group
    (512, 128, 128, 128),     # C = 128 (BASELINE configs[3]) stage 0
    (256, 96, 64, 32),        # one slab per group: an epilogue group that owns no slab of a group
    (200, 288, 96, 96),       # three groups: ragged last chunk holds a single group
    (384, 64, 32, 64),
])
@pytest.mark.parametrize("with_aux", [False, True])
def test_ln_epilogue_matches_fp64(T, N, K, G, with_aux):
    a, w, b = _data(T, N, K, seed=5)
    dev = a.device
    g = torch.Generator().manual_seed(T + N)
    gamma = (1.0 + 0.3 * torch.randn(G, generator=g)).to(dev)
    beta = (0.2 * torch.randn(G, generator=g)).to(dev)
    aux = torch.randn(T, N, generator=g).to(dev) if with_aux else None
    a = a + 0.7  # a non-zero row mean: the one-pass variance must not cancel
    pre_w = a.double() @ w.double().t() + b.double()
    v = pre_w.view(T * (N // G), G)
    mean_w, var_w = v.mean(1), v.var(1, unbiased=False)
    y_w = ((v - mean_w[:, None]) / torch.sqrt(var_w[:, None] + 1e-5) * gamma.double() + beta.double()).view(T, N)
    if with_aux:
        y_w = y_w + aux.double()
    for save in (True, False):
        y, pre, mean, rstd = _gemm3_ln(a, _split(w), N, G, b, gamma, beta, aux, save)
        assert rel_err(y.cpu(), y_w.cpu()) < LN_TOL, (save, rel_err(y.cpu(), y_w.cpu()))
        if save:
            assert rel_err(pre.cpu(), pre_w.cpu()) < TOL
            assert rel_err(mean.cpu(), mean_w.cpu()) < 1e-5
            assert rel_err(rstd.cpu(), (1.0 / torch.sqrt(var_w + 1e-5)).cpu()) < 1e-5
    # no bias
    y, _, _, _ = _gemm3_ln(a, _split(w), N, G, None, gamma, beta, aux, False)
    v = (pre_w - b.double()).view(T * (N // G), G)
    y_w = ((v - v.mean(1, keepdim=True)) / torch.sqrt(v.var(1, unbiased=False, keepdim=True) + 1e-5) * gamma.double()
           + beta.double()).view(T, N)
    if with_aux:
        y_w = y_w + aux.double()
    assert rel_err(y.cpu(), y_w.cpu()) < LN_TOL


def test_ln_epilogue_full_size_is_deterministic_and_matches_two_launches():
    """BASELINE stage-0 shape (T = 1.57 M tokens, C = 96): two runs are bit-identical and equal the unfused pair
    (hs_gemm3 + hs_layernorm_fwd with the residual) to fp32 rounding."""
    from heal_swin_b200 import ops

    dev = torch.device("cuda:0")
    T, Cc = 8 * 196608, 96
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(T, Cc, device=dev, generator=g)
    sc = torch.randn(T, Cc, device=dev, generator=g)
    lin = torch.nn.Linear(Cc, Cc).to(dev)
    norm = torch.nn.LayerNorm(Cc).to(dev)
    with torch.no_grad():
        norm.weight.add_(0.1 * torch.randn(Cc, device=dev))
        norm.bias.add_(0.1 * torch.randn(Cc, device=dev))
        assert ops.linear_ln_supported(x, lin.weight, norm)
        y1 = ops.linear_ln(x, lin.weight, lin.bias, norm, residual=sc)
        y2 = ops.linear_ln(x, lin.weight, lin.bias, norm, residual=sc)
        assert torch.equal(y1, y2)
        want = ops.layer_norm(ops.linear(x, lin.weight), norm, residual=sc, pre_bias=lin.bias)
        assert float((y1 - want).abs().max()) < 2e-5


@pytest.mark.parametrize("T,Cin,Cout,G,res", [(777, 96, 96, 96, True), (512, 192, 384, 96, False), (384, 384, 192, 192, True)])
def test_linear_ln_autograd_matches_torch_fp64(T, Cin, Cout, G, res):
    """ops.linear_ln forward and all five gradients (x, W, b, gamma, beta) + the residual's against fp64 autograd."""
    from heal_swin_b200 import ops

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(T)
    x = torch.randn(T, Cin, generator=g).to(dev).requires_grad_()
    r = torch.randn(T, Cout, generator=g).to(dev).requires_grad_() if res else None
    lin = torch.nn.Linear(Cin, Cout).to(dev)
    norm = torch.nn.LayerNorm(G).to(dev)
    with torch.no_grad():
        norm.weight.add_(0.2 * torch.randn(G, device=dev))
        norm.bias.add_(0.2 * torch.randn(G, device=dev))
    assert ops.linear_ln_supported(x, lin.weight, norm)
    y = ops.linear_ln(x, lin.weight, lin.bias, norm, residual=r)
    dy = torch.randn(T, Cout, generator=g).to(dev)
    params = [x, lin.weight, lin.bias, norm.weight, norm.bias] + ([r] if res else [])
    got = torch.autograd.grad(y, params, dy)
    p64 = [p.detach().double().requires_grad_() for p in params]
    pre = torch.nn.functional.linear(p64[0], p64[1], p64[2])
    y64 = torch.nn.functional.layer_norm(pre.view(T, Cout // G, G), (G,), p64[3], p64[4], 1e-5).view(T, Cout)
    if res:
        y64 = y64 + p64[5]
    want = torch.autograd.grad(y64, p64, dy.double())
    assert rel_err(y.detach().cpu(), y64.detach().cpu()) < LN_TOL
    for name, a_, b_ in zip(["x", "W", "b", "gamma", "beta", "res"], got, want):
        tol = 2e-3 if name == "W" else 2e-4  # weight gradients are TF32 tensor-core sums (tests/test_gpu_wgrad.py)
        assert rel_err(a_.cpu(), b_.cpu()) < tol, (name, rel_err(a_.cpu(), b_.cpu()))


# ---------------------------------------------------------------------------------------------------------------------
# LayerNorm folded into the GEMM from the input side (hs_gemm3_lnin): PatchMerging (swin_hp_transformer.py:378-395)

@pytest.mark.parametrize("T,N,K", [
    (300, 96, 192),       # tiny config (C = 48): 6 chunks per tile, ragged tokens
    (4096, 192, 384),     # stage 0 -> 1: A through tensor memory, 12 chunks
    (1000, 384, 768),     # stage 1 -> 2: in-place split ("ss"), CTA pairs, ragged tokens
    (1536, 768, 1536),    # stage 2 -> 3
    (640, 200, 224),      # odd chunk count (7): the two converter teams swap parity every tile; N not a multiple of 32
    (129, 64, 160),
])
def test_ln_prologue_matches_fp64(T, N, K):
    from heal_swin_b200._lib import check, current_stream, lib, ptr

    a, w, _ = _data(T, N, K, seed=9)
    dev = a.device
    g = torch.Generator().manual_seed(T + K)
    gamma = (1.0 + 0.3 * torch.randn(K, generator=g)).to(dev)
    beta = (0.2 * torch.randn(K, generator=g)).to(dev)
    a = a * (0.5 + torch.rand(T, 1, generator=g).to(dev)) + 1.5 * torch.randn(T, 1, generator=g).to(dev)  # row mean / scale vary
    assert lib.hs_gemm3_lnin_supported(T, N, K)
    wg = (w * gamma).contiguous()
    wsum, b0 = wg.sum(1), w @ beta
    d = torch.full((T, N), float("nan"), device=dev)
    mean = torch.full((T,), float("nan"), device=dev)
    rstd = torch.full((T,), float("nan"), device=dev)
    check(lib.hs_gemm3_lnin(ptr(a), ptr(_split(wg)), ptr(wsum), ptr(b0), ptr(d), ptr(mean), ptr(rstd), T, N, K,
                            C.c_float(1e-5), 0, current_stream()))
    xn = torch.nn.functional.layer_norm(a.double(), (K,), gamma.double(), beta.double(), 1e-5)
    want = xn @ w.double().t()
    # the epilogue subtracts mean * s from the raw product: where |mean| >> std the result carries the product's 2^-16
    # relative error amplified by that ratio (here <= ~6), still well inside the network's 1e-3
    assert rel_err(d.cpu(), want.cpu()) < 1e-4, rel_err(d.cpu(), want.cpu())
    assert rel_err(mean.cpu(), a.double().mean(1).cpu()) < 1e-5
    assert rel_err(rstd.cpu(), (1.0 / torch.sqrt(a.double().var(1, unbiased=False) + 1e-5)).cpu()) < 1e-5


def test_patch_merging_module_matches_torch_fp64():
    """PatchMerging through ops.ln_linear: output and gradients (x, norm, reduction) against fp64 autograd."""
    from heal_swin_b200 import ops
    from heal_swin_b200.models_torch.swin_hp_transformer import PatchMerging

    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    m = PatchMerging(96).to(dev)
    with torch.no_grad():
        m.norm.weight.add_(0.2 * torch.randn(384, device=dev))
        m.norm.bias.add_(0.2 * torch.randn(384, device=dev))
    x = torch.randn(2, 4 * 1100, 96, device=dev, requires_grad=True)
    assert ops.ln_linear_supported(x.view(2, 1100, 384), m.norm, m.reduction.weight)
    y = m(x)
    gy = torch.randn_like(y)
    params = [x, m.norm.weight, m.norm.bias, m.reduction.weight]
    got = torch.autograd.grad(y, params, gy)
    p64 = [p.detach().double().requires_grad_() for p in params]
    y64 = torch.nn.functional.linear(
        torch.nn.functional.layer_norm(p64[0].view(2, 1100, 384), (384,), p64[1], p64[2], 1e-5), p64[3])
    want = torch.autograd.grad(y64, p64, gy.double())
    assert rel_err(y.detach().cpu(), y64.detach().cpu()) < 1e-4
    for name, a_, b_ in zip(["x", "gamma", "beta", "W"], got, want):
        assert rel_err(a_.cpu(), b_.cpu()) < (2e-3 if name == "W" else 3e-4), (name, rel_err(a_.cpu(), b_.cpu()))


def test_batched_weight_split_equals_individual_splits():
    """ops.refresh_weight_splits (one hs_weight_split_batch launch over every cached operand) writes bit-identical
    operands to the per-matrix kernel, for forward / transposed / column-block / TF32-format entries, and picks up
    parameter updates that did not bump the version counter (fused optimizers)."""
    from heal_swin_b200 import _lib, ops

    dev = torch.device("cuda:0")
    ops.invalidate_weight_splits()
    torch.manual_seed(0)
    ws = [torch.nn.Parameter(torch.randn(n, k, device=dev)) for n, k in [(96, 96), (288, 96), (200, 100), (96, 384), (384, 192)]]
    outs = []
    for w in ws:
        outs.append(ops.split_weight(w))
        outs.append(ops.split_weight(w, transposed=True))
        outs.append(ops.split_weight(w, transposed=True, prec=_lib.PREC_TF32))
    outs.append(ops.split_weight(ws[4], cols=(96, 96)))
    with torch.no_grad():
        for w in ws:
            w.data.mul_(1.5).add_(0.25)  # (no version bump through .data)
    n = ops.refresh_weight_splits(dev)
    assert n == len(outs)
    got = [o.clone() for o in outs]
    ops.invalidate_weight_splits()
    want = []
    for w in ws:
        want.append(ops.split_weight(w))
        want.append(ops.split_weight(w, transposed=True))
        want.append(ops.split_weight(w, transposed=True, prec=_lib.PREC_TF32))
    want.append(ops.split_weight(ws[4], cols=(96, 96)))
    for a_, b_ in zip(got, want):
        assert torch.equal(a_.view(torch.int16), b_.view(torch.int16))


# ---------------------------------------------------------------------------------------------------------------------
# Compact MLP activation record: the forward saves g' = GELU'(fc1(x) + b1) * dropmask as FP16 instead of z (hs_gemm3 mode 6),
# the backward multiplies by it (mode 7; hs_mlp_dgrad_gelu with HS_MLP_GRAD16: tests/test_gpu_mlp.py)

def _gelu_grad64(u):
    return 0.5 * (1 + torch.erf(u / math.sqrt(2))) + u * torch.exp(-0.5 * u * u) / math.sqrt(2 * math.pi)


def _gemm3_raw(a, ws, N, bias, aux, d, d2, mode, prec=0, drop=0.0, seed=0):
    from heal_swin_b200._lib import check, current_stream, lib, ptr

    T, K = a.shape
    check(lib.hs_gemm3(ptr(a), ptr(ws), ptr(bias), ptr(aux), ptr(d), ptr(d2), None, T, N, K, mode, prec, C.c_float(drop),
                       C.c_uint64(seed), current_stream()))


@pytest.mark.parametrize("T,N,K", [(300, 96, 96), (4096, 384, 96), (2048, 768, 192), (1000, 1536, 384), (130, 64, 48)])
def test_compact_gelu_modes_match_fp64(T, N, K):
    """mode 6: h as mode 2 (2e-5) and g' to FP16 rounding (2^-11 of its magnitude); mode 7: acc * g' in every precision
    the input gradients use, ragged token counts included (rows beyond T are neither written nor read)."""
    a, w, b = _data(T, N, K, seed=11)
    dev = a.device
    u = a.double() @ w.double().t() + b.double()
    gp = torch.full((T + 3, N), float("nan"), device=dev, dtype=torch.float16)  # (guard rows: must stay untouched)
    h = torch.full((T, N), float("nan"), device=dev)
    _gemm3_raw(a, _split(w), N, b, None, gp, h, mode=6)
    assert rel_err(h.cpu(), torch.nn.functional.gelu(u).cpu()) < TOL
    assert bool(torch.isnan(gp[T:]).all())
    g64 = _gelu_grad64(u)
    assert float((gp[:T].double() - g64).abs().max()) < 2.0 ** -11 * 1.2 + 1e-5
    # mode 7 consumes exactly that tensor
    for prec, tol in ((0, TOL), (1, 1e-3)):
        out = torch.full((T, N), float("nan"), device=dev)
        a2, wz, _ = _data(T, N, K, seed=13)
        _gemm3_raw(a2, _split(wz, prec=prec), N, None, gp, out, None, mode=7, prec=prec)
        want = (a2.double() @ wz.double().t()) * gp[:T].double()
        assert rel_err(out.cpu(), want.cpu()) < tol, (prec, rel_err(out.cpu(), want.cpu()))


def test_compact_gelu_dropout_mask_is_shared_by_h_and_the_derivative():
    """With drop > 0 the mask (and the 1 / (1 - p) scale) is folded into BOTH outputs of mode 6, identically: h == 0 exactly
    where g' == 0, at the rate p, and the kept entries carry the scale -- the backward needs neither seed nor mask."""
    T, N, K = 1500, 384, 96
    a, w, b = _data(T, N, K, seed=2)
    a = a + 3.0  # GELU and GELU' > 0 almost everywhere: zero <=> dropped
    w = w.abs()
    b = b.abs()
    drop, seed = 0.25, 0x1234_5678_9ABC
    gp = torch.empty((T, N), device=a.device, dtype=torch.float16)
    h = torch.empty((T, N), device=a.device)
    _gemm3_raw(a, _split(w), N, b, None, gp, h, mode=6, drop=drop, seed=seed)
    z, h2 = _gemm3(a, _split(w), N, b, mode=2, drop=drop, seed=seed)   # the fp32 form with the same seed
    assert torch.equal(h == 0, h2 == 0) and rel_err(h.cpu(), h2.cpu()) < 1e-6
    assert torch.equal(h == 0, gp == 0) and 0.2 < float((h == 0).float().mean()) < 0.3
    u = a.double() @ w.double().t() + b.double()
    keep = (h != 0).double()
    assert float((gp.double() - _gelu_grad64(u) * keep / (1 - drop)).abs().max()) < 2.0 ** -10 * 1.6
