"""torch.autograd.Function wrappers around the C-ABI device kernels.

PyTorch is only the plumbing here (device memory, streams, autograd graph); every op below
launches hand-written CUDA from libhealswin_b200 on torch's current stream and raises if the
tensors are not on a CUDA device.
"""
import ctypes as C
import os
import weakref

import torch

from . import _lib
from ._lib import check, current_stream, lib, ptr, require_cuda


class _Stats:
    """Launch accounting for bench.py: how many of this library's kernels were launched, and
    (when ``timing`` is on) CUDA-event pairs around named launches on the launching stream."""

    def __init__(self):
        self.launches = 0
        self.by_name = {}
        self.timing = False
        self.events = {}

    def reset(self):
        self.launches = 0
        self.by_name = {}
        self.events = {}

    def launch(self, name, fn, *args, tag=None):
        self.launches += 1
        self.by_name[name] = self.by_name.get(name, 0) + 1
        if self.timing and tag is not None:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            check(fn(*args))
            b.record()
            self.events.setdefault((name, tag), []).append((a, b))
        else:
            check(fn(*args))

    def elapsed_ms(self):
        """{(name, tag): [ms, ...]} -- call after torch.cuda.synchronize()."""
        return {k: [a.elapsed_time(b) for a, b in v] for k, v in self.events.items()}


STATS = _Stats()

# Precision of the attention matmuls (q k^T and p v).  "tf32": tcgen05 tensor-core kernels, TF32 operands with
# fp32 accumulation (what torch.backends.cuda.matmul.allow_tf32 gives the reference's q @ k^T / attn @ v on GPU);
# "fp32": exact-fp32 CUDA-core kernels (also used automatically for shapes the tensor-core kernels do not cover).
_ATTN_PRECISION = os.environ.get("HEALSWIN_ATTN_PRECISION", "tf32")


def set_attention_precision(mode: str) -> None:
    global _ATTN_PRECISION
    assert mode in ("tf32", "fp32"), mode
    _ATTN_PRECISION = mode


def get_attention_precision() -> str:
    return _ATTN_PRECISION


class ZeroArena:
    """One flat fp32 buffer that serves the small zero-initialised accumulators of a training step (weight / bias / norm
    gradients, loss sums: ~360 per step of the N_side=256 network) from ONE memset instead of one fill kernel each.
    Used by heal_swin_b200.graph.GraphedTrainStep: ``begin()`` zeroes the buffer and rewinds it, every ``zeros()`` below
    carves the next 256-byte-aligned slice.  A first pass with an empty arena only measures the demand."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.buf = None
        self.off = 0
        self.need = 0

    def begin(self):
        if self.buf is None and self.need:
            self.buf = torch.empty(self.need, device=self.device, dtype=torch.float32)
        if self.buf is not None:
            self.buf.zero_()
        self.off = 0
        self.need = 0

    def take(self, n):
        a = (n + 63) // 64 * 64
        self.need += a
        if self.buf is None or self.off + a > self.buf.numel():
            return None
        v = self.buf[self.off:self.off + n]
        self.off += a
        return v


_ARENA = [None]


def use_zero_arena(arena):
    """Install (or, with None, remove) the arena ``zeros`` serves from; returns the previous one."""
    prev, _ARENA[0] = _ARENA[0], arena
    return prev


def zeros(shape, device):
    """fp32 zeros for a gradient accumulator: a slice of the active ZeroArena, else ``torch.zeros``."""
    ar = _ARENA[0]
    if ar is not None and ar.device == torch.device(device):
        n = 1
        for d in (shape if isinstance(shape, (tuple, list, torch.Size)) else (shape,)):
            n *= int(d)
        v = ar.take(n)
        if v is not None:
            return v.view(shape)
    return torch.zeros(shape, device=device, dtype=torch.float32)


def _f32c(t):
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class GatherRows(torch.autograd.Function):
    """out[b, p] = x[b, idx[p]] (shifter.shift / shift_back); backward gathers with the inverse."""

    @staticmethod
    def forward(ctx, x, idx_i32, inv_i32):
        require_cuda(x, idx_i32)
        x = _f32c(x)
        B, N, Cc = x.shape
        out = torch.empty_like(x)
        STATS.launch("gather_rows", lib.hs_gather_rows, ptr(x), ptr(idx_i32), ptr(out), B, N, Cc, current_stream(),
                     tag=(B, N, Cc))
        ctx.inv = inv_i32
        return out

    @staticmethod
    def backward(ctx, g):
        g = _f32c(g)
        B, N, Cc = g.shape
        out = torch.empty_like(g)
        STATS.launch("gather_rows", lib.hs_gather_rows, ptr(g), ptr(ctx.inv), ptr(out), B, N, Cc, current_stream())
        return out, None, None


_SEED_COUNTERS = {}   # device -> int64 tensor [1]: the replay counter indirect seeds refer to
_SEED_CALL_ID = [0]


def seed_counter(device):
    """The device-side counter behind indirect dropout seeds (csrc/hs_common.h: resolve_seed).  A captured training step
    increments it once per replay (heal_swin_b200/graph.py), which gives every replay fresh masks."""
    dev = torch.device(device)
    t = _SEED_COUNTERS.get(dev)
    if t is None:
        base = int(torch.randint(0, 2**31, (1,)).item()) + (int(os.environ.get("RANK", "0")) << 32)
        t = _SEED_COUNTERS[dev] = torch.tensor([base], device=dev, dtype=torch.int64)
    return t


def _next_dropout_seed() -> int:
    """64-bit seed of one dropout mask.  Eagerly: drawn from torch's default CPU generator -- reproducible under
    torch.manual_seed, replayed identically when torch.utils.checkpoint re-runs the forward (it restores the RNG state),
    no device sync; the rank is mixed in so that data-parallel shards do not share masks.  Inside a CUDA-graph capture a
    host integer would be baked into the graph and every replay would repeat the mask: there the seed is INDIRECT (bit
    63 set, a call id in bits 48-62, the address of the device-side replay counter in bits 0-47)."""
    if torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
        _SEED_CALL_ID[0] = (_SEED_CALL_ID[0] + 1) & 0x7FFF
        ptr_ = seed_counter(torch.device("cuda", torch.cuda.current_device())).data_ptr()
        assert ptr_ < (1 << 48)
        return (1 << 63) | (_SEED_CALL_ID[0] << 48) | ptr_
    x = int(torch.randint(0, 2**62, (1,)).item()) ^ (int(os.environ.get("RANK", "0")) * 0xD1B54A32D192ED03)
    return x & 0x7FFFFFFFFFFFFFFF


def _pick_seed(seed, p):
    """A fresh seed when dropout is active and none was given; a caller's seed is a plain VALUE (bit 63 is reserved for
    the indirect seeds of graph captures and is cleared)."""
    if seed is None:
        return _next_dropout_seed() if p > 0.0 else 0
    return int(seed) & 0x7FFFFFFFFFFFFFFF


class WindowAttnCore(torch.autograd.Function):
    """shift -> window_partition -> softmax(q k^T [cos] + bias + mask) v -> window_reverse -> shift_back
    on a packed (B, N, 3C) qkv tensor (swin_hp_transformer.py:136-171, 319-330)."""

    @staticmethod
    def forward(ctx, qkv, bias_table, logit_scale, src, groups, dense_mask, rel_index_i32,
                scale, num_heads, window_size, use_cos, attn_drop=0.0, seed=0):
        require_cuda(qkv)
        qkv = _f32c(qkv)
        B, N, C3 = qkv.shape
        Cc = C3 // 3
        H, ws = int(num_heads), int(window_size)
        stream = current_stream()
        bias = None
        if bias_table is not None:
            table = _f32c(bias_table)
            T = table.shape[0]
            assert rel_index_i32.numel() == ws * ws, (
                f"relative position index is {tuple(rel_index_i32.shape)} but the window has {ws} tokens"
            )
            bias = torch.empty((H, ws, ws), device=qkv.device, dtype=torch.float32)
            STATS.launch("rel_bias_expand", lib.hs_rel_bias_expand, ptr(table), ptr(rel_index_i32), ptr(bias), T, H, ws, stream)
        ls = _f32c(logit_scale.reshape(-1)) if (use_cos and logit_scale is not None) else None
        mask = _f32c(dense_mask) if dense_mask is not None else None
        out = torch.empty((B, N, Cc), device=qkv.device, dtype=torch.float32)
        # log2-domain log-sum-exp per (head, token): lets the backward skip the softmax-statistics pass
        # planes: 0 = log-sum-exp; 1, 2 = 1/|q|, 1/|k| (cos attention), reused by the backward
        lse = (torch.empty((3 if use_cos else 1, H, B * N), device=qkv.device, dtype=torch.float32)
               if ctx.needs_input_grad[0] else None)
        flags = ((_lib.ATTN_COS if use_cos else 0) | (_lib.ATTN_NO_TC if _ATTN_PRECISION == "fp32" else 0)
                 | (_lib.ATTN_NO_TRUNC_COMP if os.environ.get("HEALSWIN_NO_TRUNC_COMP") == "1" else 0))
        STATS.launch("window_attn_fwd", lib.hs_window_attn_fwd, ptr(qkv), ptr(src), ptr(groups), ptr(mask), ptr(bias),
                     ptr(ls), C.c_float(scale), C.c_float(attn_drop), C.c_uint64(seed), ptr(out), ptr(lse), B, N, Cc, H, ws,
                     flags, stream, tag=(B, N, Cc, H, ws))
        ctx.save_for_backward(qkv, bias, ls, src, groups, mask, rel_index_i32, out, lse)
        ctx.drop = (float(attn_drop), int(seed))
        ctx.meta = (scale, H, ws, flags, bias_table is not None,
                    None if bias_table is None else tuple(bias_table.shape),
                    None if logit_scale is None else tuple(logit_scale.shape))
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, bias, ls, src, groups, mask, rel_index, out, lse = ctx.saved_tensors
        scale, H, ws, flags, has_table, table_shape, ls_shape = ctx.meta
        dout = _f32c(dout)
        B, N, C3 = qkv.shape
        Cc = C3 // 3
        stream = current_stream()
        dqkv = torch.empty_like(qkv)
        need_table = has_table and ctx.needs_input_grad[1]
        need_ls = (ls is not None) and ctx.needs_input_grad[2]
        dbias = zeros((H, ws, ws), qkv.device) if need_table else None
        dls = zeros((H,), qkv.device) if need_ls else None
        STATS.launch("window_attn_bwd", lib.hs_window_attn_bwd, ptr(qkv), ptr(out), ptr(lse), ptr(dout), ptr(src), ptr(groups), ptr(mask),
                     ptr(bias), ptr(ls), C.c_float(scale), C.c_float(ctx.drop[0]), C.c_uint64(ctx.drop[1]), ptr(dqkv),
                     ptr(dbias), ptr(dls), B, N, Cc, H, ws,
                     flags, stream, tag=(B, N, Cc, H, ws))
        dtable = None
        if need_table:
            dtable = zeros(table_shape, qkv.device)
            STATS.launch("rel_bias_reduce", lib.hs_rel_bias_reduce, ptr(dbias), ptr(rel_index), ptr(dtable),
                         table_shape[0], H, ws, stream)
        if need_ls:
            dls = dls.reshape(ls_shape)
        return dqkv, dtable, dls, None, None, None, None, None, None, None, None, None, None


def window_attention_core(qkv, bias_table, logit_scale, src, groups, dense_mask, rel_index_i32,
                          scale, num_heads, window_size, use_cos, attn_drop=0.0, seed=None):
    """``attn_drop`` > 0 applies dropout to the attention probabilities inside the kernel (training mode of
    ``nn.Dropout(attn_drop)``, swin_hp_transformer.py:167-169); ``seed`` fixes the mask (default: a fresh one)."""
    attn_drop = float(attn_drop)
    seed = _pick_seed(seed, attn_drop)
    return WindowAttnCore.apply(qkv, bias_table, logit_scale, src, groups, dense_mask, rel_index_i32,
                                float(scale), num_heads, window_size, bool(use_cos), attn_drop, int(seed or 0))


class LayerNormFn(torch.autograd.Function):
    """y = residual + row_scale * (LayerNorm(dropout(x + pre_bias)) * weight + bias) over the last dim (pre_bias,
    residual, row_scale and the dropout optional), csrc/hs_layernorm.cu."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, pre_bias, eps, row_scale, in_drop, seed):
        require_cuda(x, weight, bias, residual, pre_bias, row_scale)
        shape = x.shape
        Cc = shape[-1]
        x2 = _f32c(x).reshape(-1, Cc)
        rows = x2.shape[0]
        w, b = _f32c(weight), _f32c(bias)
        res2 = _f32c(residual).reshape(-1, Cc) if residual is not None else None
        pb = _f32c(pre_bias) if pre_bias is not None else None
        rsc = _f32c(row_scale).reshape(-1) if row_scale is not None else None
        rps = rows // rsc.numel() if rsc is not None else 0
        assert rsc is None or rps * rsc.numel() == rows, "row_scale must divide the rows evenly"
        y = torch.empty_like(x2)
        mean = torch.empty(rows, device=x2.device, dtype=torch.float32)
        rstd = torch.empty(rows, device=x2.device, dtype=torch.float32)
        STATS.launch("layernorm_fwd", lib.hs_layernorm_fwd, ptr(x2), ptr(pb), ptr(res2), ptr(w), ptr(b), ptr(rsc), rps,
                     C.c_float(in_drop), C.c_uint64(seed), ptr(y), ptr(mean), ptr(rstd), rows, Cc, C.c_float(eps),
                     current_stream(), tag=(rows, Cc, int(res2 is not None)))
        ctx.save_for_backward(x2, w, mean, rstd, pb, rsc)
        ctx.has_res = residual is not None
        ctx.shape = shape
        ctx.extra = (rps, float(in_drop), int(seed))
        return y.view(shape)

    @staticmethod
    def backward(ctx, dy):
        x2, w, mean, rstd, pb, rsc = ctx.saved_tensors
        rps, in_drop, seed = ctx.extra
        rows, Cc = x2.shape
        dy2 = _f32c(dy).reshape(rows, Cc)
        dx = torch.empty_like(x2)
        need_w, need_b = ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        need_pb = pb is not None and ctx.needs_input_grad[4]
        dw = zeros(Cc, x2.device) if need_w else None
        db = zeros(Cc, x2.device) if need_b else None
        dpb = zeros(Cc, x2.device) if need_pb else None
        STATS.launch("layernorm_bwd", lib.hs_layernorm_bwd, ptr(dy2), ptr(x2), ptr(pb), ptr(mean), ptr(rstd), ptr(w),
                     ptr(rsc), rps, C.c_float(in_drop), C.c_uint64(seed), ptr(dx), ptr(dw), ptr(db), ptr(dpb), rows, Cc,
                     current_stream(), tag=(rows, Cc))
        dres = dy if ctx.has_res else None
        return dx.view(ctx.shape), dw, db, dres, dpb, None, None, None, None


class _CrossEntropyFn(torch.autograd.Function):
    """Mean cross-entropy over (B, K, P) logits and (B, P) class ids with the gradient produced in the same pass
    (csrc/hs_cross_entropy.cu): the backward is one scaling of the saved gradient."""

    @staticmethod
    def forward(ctx, logits, target, ignore_index):
        require_cuda(logits, target)
        x = _f32c(logits)
        B, K = x.shape[0], x.shape[1]
        P = x.numel() // (B * K)
        t = target.contiguous()
        assert t.numel() == B * P and t.dtype in (torch.uint8, torch.int64), "targets: (B, P) uint8 or int64"
        dl = torch.empty_like(x)
        acc = zeros(2, x.device)
        STATS.launch("cross_entropy", lib.hs_cross_entropy, ptr(x), ptr(t), t.element_size(), ptr(dl), ptr(acc), B, K, P,
                     int(ignore_index), current_stream(), tag=(B * P, K))
        ctx.save_for_backward(dl, acc)
        return acc[0] / acc[1]

    @staticmethod
    def backward(ctx, g):
        dl, acc = ctx.saved_tensors
        return dl.mul_(g / acc[1]), None, None  # (the saved gradient is consumed: one backward per forward)


def cross_entropy(logits, target, ignore_index=-100):
    """``F.cross_entropy(logits, target)`` (mean reduction, no class weights, no label smoothing) for the network output
    (B, K, *spatial) and class ids (B, *spatial); uint8 targets are accepted as they come from the data pipeline."""
    if (logits.is_cuda and logits.dtype == torch.float32 and logits.dim() >= 3 and target.dtype in (torch.uint8, torch.int64)
            and lib.hs_cross_entropy_supported(logits.shape[1])):
        return _CrossEntropyFn.apply(logits, target, ignore_index)
    return torch.nn.functional.cross_entropy(logits, target.long(), ignore_index=ignore_index)


class CrossEntropyLoss(torch.nn.Module):
    """Drop-in for ``nn.CrossEntropyLoss()`` with default arguments (model_lightning_swin_hp.py:45)."""

    def __init__(self, ignore_index=-100):
        super().__init__()
        self.ignore_index = ignore_index

    def forward(self, logits, target):
        return cross_entropy(logits, target, self.ignore_index)


def _fusable_norm(norm, x):
    return (isinstance(norm, torch.nn.LayerNorm) and norm.elementwise_affine and norm.bias is not None
            and len(norm.normalized_shape) == 1 and norm.normalized_shape[0] == x.shape[-1])


def layer_norm(x, norm, residual=None, pre_bias=None, row_scale=None, in_drop=0.0, seed=None):
    """``residual + row_scale * norm(dropout(x + pre_bias))`` for an ``nn.LayerNorm`` module ``norm`` (affine, normalising
    the last dim) in one launch.  ``row_scale``: one factor per leading-dimension sample (stochastic depth), a constant
    (no gradient).  ``in_drop``: dropout probability on the LayerNorm input (counter-based mask, fresh seed by default).
    Any other norm layer is applied as the module it is."""
    in_drop = float(in_drop)
    if _fusable_norm(norm, x):
        seed = _pick_seed(seed, in_drop)
        return LayerNormFn.apply(x, norm.weight, norm.bias, residual, pre_bias, float(norm.eps), row_scale, in_drop,
                                 int(seed or 0))
    h = x if pre_bias is None else x + pre_bias
    if in_drop > 0.0:
        h = torch.nn.functional.dropout(h, in_drop, True)
    y = norm(h)
    if row_scale is not None:
        y = y * row_scale.view(-1, *([1] * (y.dim() - 1)))
    return y if residual is None else residual + y


class LnHeadFn(torch.autograd.Function):
    """logits (B, K, P) = Conv1d_1x1(LayerNorm(x (B, P, C))) in one pass, csrc/hs_ln_head.cu; the backward is one pass too
    and never materialises the normalised activation or its gradient."""

    @staticmethod
    def forward(ctx, x, gamma, beta, w, hbias, eps):
        require_cuda(x, gamma, beta, w)
        B, P, Cc = x.shape
        K = w.shape[0]
        x2 = _f32c(x).reshape(B * P, Cc)
        gamma, beta, w, hbias = _f32c(gamma), _f32c(beta), _f32c(w), _f32c(hbias)
        logits = torch.empty((B, K, P), device=x.device, dtype=torch.float32)
        mean = torch.empty((B * P,), device=x.device, dtype=torch.float32)
        rstd = torch.empty_like(mean)
        STATS.launch("ln_head_fwd", lib.hs_ln_head_fwd, ptr(x2), ptr(gamma), ptr(beta), ptr(w), ptr(hbias), ptr(logits),
                     ptr(mean), ptr(rstd), B * P, P, Cc, K, C.c_float(eps), current_stream(), tag=(B * P, Cc, K))
        ctx.save_for_backward(x2, gamma, beta, w, mean, rstd)
        ctx.meta = (B, P, Cc, K, hbias is not None)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        x2, gamma, beta, w, mean, rstd = ctx.saved_tensors
        B, P, Cc, K, has_bias = ctx.meta
        dl = _f32c(dlogits)
        dx = torch.empty_like(x2)
        s_acc = zeros((K, Cc), x2.device)
        g_acc = zeros((K,), x2.device)
        STATS.launch("ln_head_bwd", lib.hs_ln_head_bwd, ptr(dl), ptr(x2), ptr(mean), ptr(rstd), ptr(gamma), ptr(w), ptr(dx),
                     ptr(s_acc), ptr(g_acc), B * P, P, Cc, K, current_stream(), tag=(B * P, Cc, K))
        dw = gamma * s_acc + beta * g_acc[:, None]
        dgamma = (w * s_acc).sum(0)
        dbeta = (w * g_acc[:, None]).sum(0)
        return dx.view(B, P, Cc), dgamma, dbeta, dw, (g_acc if has_bias else None), None


def ln_head(x, norm, weight, bias=None):
    """``Conv1d_1x1(norm(x).transpose(1, 2))`` for x (B, P, C) and ``weight`` (K, C): logits (B, K, P).  One fused launch
    when ``norm`` is an affine LayerNorm over C and the shape is covered, else LayerNorm + linear + transpose."""
    if (x.is_cuda and x.dim() == 3 and _fusable_norm(norm, x)
            and lib.hs_ln_head_supported(x.shape[0] * x.shape[1], x.shape[2], weight.shape[0])):
        return LnHeadFn.apply(x, norm.weight, norm.bias, weight, bias, float(norm.eps))
    return linear(layer_norm(x, norm), weight, bias).permute(0, 2, 1).contiguous()


class BiasGeluFn(torch.autograd.Function):
    """h = dropout(GELU(z + bias)) (exact erf GELU), csrc/hs_bias_gelu.cu; backward also yields d(bias)."""

    @staticmethod
    def forward(ctx, z, bias, drop, seed):
        require_cuda(z, bias)
        shape = z.shape
        Cc = shape[-1]
        z2 = _f32c(z).reshape(-1, Cc)
        b = _f32c(bias) if bias is not None else None
        h = torch.empty_like(z2)
        STATS.launch("bias_gelu_fwd", lib.hs_bias_gelu_fwd, ptr(z2), ptr(b), C.c_float(drop), C.c_uint64(seed), ptr(h),
                     z2.shape[0], Cc, current_stream(), tag=(z2.shape[0], Cc))
        ctx.save_for_backward(z2, b)
        ctx.shape = shape
        ctx.drop = (float(drop), int(seed))
        return h.view(shape)

    @staticmethod
    def backward(ctx, dh):
        z2, b = ctx.saved_tensors
        rows, Cc = z2.shape
        dh2 = _f32c(dh).reshape(rows, Cc)
        dz = torch.empty_like(z2)
        need_b = b is not None and ctx.needs_input_grad[1]
        db = zeros(Cc, z2.device) if need_b else None
        STATS.launch("bias_gelu_bwd", lib.hs_bias_gelu_bwd, ptr(dh2), ptr(z2), ptr(b), C.c_float(ctx.drop[0]),
                     C.c_uint64(ctx.drop[1]), ptr(dz), ptr(db), rows, Cc, current_stream(), tag=(rows, Cc))
        return dz.view(ctx.shape), db, None, None


def bias_gelu(z, bias, drop=0.0, seed=None):
    """``dropout(GELU(z + bias))``: one fused launch for widths the kernel covers (C % 4 == 0, C <= 4096), torch ops for
    the rest (e.g. embed_dim 192 with four stages gives a hidden width of 6144)."""
    drop = float(drop)
    if not (z.is_cuda and lib.hs_bias_gelu_supported(z.numel() // z.shape[-1], z.shape[-1])):
        h = torch.nn.functional.gelu(z if bias is None else z + bias)
        return torch.nn.functional.dropout(h, drop, True) if drop > 0.0 else h
    seed = _pick_seed(seed, drop)
    return BiasGeluFn.apply(z, bias, drop, int(seed or 0))


# ----------------------------------------------------------------------------------------------------------------------
# Dense linears.  Forward and input-gradient products run on the hand-written bf16x3 tensor-core GEMM
# (csrc/hs_gemm3_tc.cu: every fp32 operand split into two bf16 terms, three MMAs, ~2^-16 relative per product -- the
# network output stays at fp32-GEMM accuracy, which the TF32 library GEMMs they replace did not: 1.1-1.4e-3 against a
# tolerance of 1e-3), with the bias, the residual-shortcut gradient and the GELU of the MLP in their epilogues; weight
# gradients on the token-split TF32 kernel (csrc/hs_wgrad_tc.cu).  HEALSWIN_GEMM=library (or set_gemm_mode) sends all
# of them through cuBLAS instead (diagnostics: torch's global allow_tf32 switch then decides their precision).

_GEMM_MODE = os.environ.get("HEALSWIN_GEMM", "bf16x3")
_CUSTOM_WGRAD = os.environ.get("HEALSWIN_CUSTOM_WGRAD", "1") == "1"
_FUSED_MLP = os.environ.get("HEALSWIN_FUSED_MLP", "1") == "1"
_TF32_MLP_DGRAD = os.environ.get("HEALSWIN_TF32_MLP_DGRAD", "1") == "1"
_MLP_COMPACT = os.environ.get("HEALSWIN_MLP_COMPACT", "1")  # "0" off | "1" tensor-bound stages | "all"


# Operand precision of the hand-written GEMMs.  "fp32" (default): bf16x3 for every forward product and for the HBM-bound
# input gradients, one TF32 MMA for the tensor-bound input gradients (stages 2-3: gradients are compared at 5e-3, and the
# weight gradients are TF32 already).  "bf16": one bf16 MMA on the hi terms everywhere -- "bf16 operands, fp32 accumulate",
# the arithmetic BASELINE configs[3] names (activations stay fp32 in HBM; tolerance stated in tests/test_gpu_bf16.py).
_GEMM_PRECISION = os.environ.get("HEALSWIN_GEMM_PRECISION", "fp32")
_TF32_DGRAD = os.environ.get("HEALSWIN_TF32_DGRAD", "1") == "1"


def set_gemm_mode(mode: str) -> None:
    """"bf16x3": the hand-written tensor-core GEMMs (default); "library": cuBLAS through torch."""
    global _GEMM_MODE
    assert mode in ("bf16x3", "library"), mode
    _GEMM_MODE = mode


def set_gemm_precision(precision: str) -> None:
    """"fp32": fp32-class products (bf16x3; TF32 for tensor-bound input gradients); "bf16": bf16 operands, fp32 accumulate."""
    global _GEMM_PRECISION
    assert precision in ("fp32", "bf16"), precision
    _GEMM_PRECISION = precision


def get_gemm_precision() -> str:
    return _GEMM_PRECISION


def _fwd_prec() -> int:
    return _lib.PREC_BF16 if _GEMM_PRECISION == "bf16" else _lib.PREC_BF16X3


def _dgrad_tensor_bound(n_out, k_contract) -> bool:
    """Whether three bf16 MMAs per product would make an input-gradient GEMM dx (T, n_out) = dy (T, k_contract) @ W
    tensor-bound (per output element 6 * k flops at ~1.2 PFLOP/s against 4 * (n + k) / n bytes at ~6 TB/s, i.e.
    n k / (n + k) > ~130: stages 2-3 of the N_side=256 network, the MLP gradients from stage 1 on)."""
    return _TF32_DGRAD and n_out * k_contract > 130 * (n_out + k_contract)


def _dgrad_prec(T, n_out, k_contract) -> int:
    """Input-gradient GEMM: fp32-class (bf16x3) where the launch is HBM-bound anyway; one TF32 MMA per product where it
    would be tensor-bound -- gradients are compared at 5e-3 and the weight gradients are TF32 already.  (Tried and
    rejected on the B200: two MMAs with dy split into bf16 hi / lo and the weight in FP16 -- a kind::f16 MMA with MIXED
    operand formats is an illegal instruction, and with both operands bf16 the weight keeps 8 bits, 4x coarser than TF32.)"""
    if _GEMM_PRECISION == "bf16":
        return _lib.PREC_BF16
    if _dgrad_tensor_bound(n_out, k_contract):
        return _lib.PREC_TF32
    return _lib.PREC_BF16X3


def get_gemm_mode() -> str:
    return _GEMM_MODE


_SPLITS = {}  # (id(weight), transposed) -> (weakref, version, data_ptr, split tensor, epoch)
_SPLIT_EPOCH = [0]


def _on_optimizer_step(*_):
    """Global torch.optim post-step hook: fused optimizers (Adam(fused=True), the one bench.py uses) update the parameters
    WITHOUT bumping their version counters, so every optimizer step starts a new epoch of the split cache."""
    _SPLIT_EPOCH[0] += 1


from torch.optim.optimizer import register_optimizer_step_post_hook as _register_post_step  # noqa: E402

_register_post_step(_on_optimizer_step)


def split_weight(weight, transposed=False, cols=None, prec=0):
    """The bf16 [hi | lo] operand of ``weight`` (N, K) for hs_gemm3: (N, 2 * ceil32(K)) for the forward, or with
    ``transposed`` (K, 2 * ceil32(N)) for the input gradient.  ``cols = (c0, n)`` takes the column block
    ``weight[:, c0:c0 + n]`` instead of the whole matrix (the two halves of a skip-concat projection), read in place through
    the row pitch.  Cached per parameter; redone when the parameter's version counter moves (load_state_dict, DDP
    broadcast, any in-place op), after every torch optimizer step (see above), and on every use inside a CUDA-graph
    capture (a replay cannot consult the host).  The buffer is reused, so its address is stable.  Call
    ``invalidate_weight_splits()`` after writing parameters in a way neither mechanism sees (``.data`` assignment from a
    custom optimizer that is not a torch.optim.Optimizer)."""
    w = _f32c(weight.detach())
    N, ld = w.shape
    c0, K = (0, ld) if cols is None else cols
    fmt = 1 if prec == _lib.PREC_TF32 else 0  # TF32 mode: fp32 rounded to the nearest TF32 (same bytes per row)
    key = (id(weight), bool(transposed), c0, K, fmt)
    ent = _SPLITS.get(key)
    ver = weight._version
    if ent is not None and ent[0]() is weight and ent[2] == w.data_ptr():
        capturing = w.is_cuda and torch.cuda.is_current_stream_capturing()
        if ent[1] == ver and ent[4] == _SPLIT_EPOCH[0] and not capturing:
            return ent[3]
        if capturing and _SPLITS_MAINTAINED[0]:  # refreshed by one batched launch before every replay
            return ent[3]
        out = ent[3]
    else:
        rows, ncols = (K, N) if transposed else (N, K)
        out = torch.empty((rows, 2 * ((ncols + 31) // 32 * 32)), device=w.device, dtype=torch.bfloat16)
    rows, ncols = (K, N) if transposed else (N, K)
    src = w if c0 == 0 else w.reshape(-1)[c0:]  # pointer to element (0, c0); the kernel walks rows with pitch ld
    STATS.launch("weight_split", lib.hs_weight_split, ptr(src), rows, ncols, ld, 1 if transposed else 0, fmt, ptr(out),
                 current_stream())
    if weight.is_leaf:  # views of a parameter (patch embedding) are new objects every call: not worth caching
        if len(_SPLITS) > 1024:
            for k in [k for k, v in _SPLITS.items() if v[0]() is None]:
                del _SPLITS[k]
        _SPLITS[key] = (weakref.ref(weight), ver, w.data_ptr(), out, _SPLIT_EPOCH[0])
    return out


def invalidate_weight_splits() -> None:
    _SPLITS.clear()
    _SPLIT_TABLES.clear()


class _SplitDesc(C.Structure):  # csrc/hs_gemm3_tc.cu: SplitDesc
    _fields_ = [("w", C.c_void_p), ("out", C.c_void_p), ("rows", C.c_int), ("cols", C.c_int), ("ld", C.c_int),
                ("transposed", C.c_int), ("format", C.c_int), ("tiles_x", C.c_int), ("tile0", C.c_int), ("pad", C.c_int)]


_SPLITS_MAINTAINED = [False]
_SPLIT_TABLES = {}  # device -> (signature, device table, n, total tiles)


def maintain_weight_splits(on: bool) -> bool:
    """While on, ``split_weight`` inside a CUDA-graph capture trusts the cached operands instead of re-launching their
    split kernels (195 launches per step of the N_side=256 network): the owner of the graph promises to call
    ``refresh_weight_splits`` before every replay.  Returns the previous setting."""
    prev, _SPLITS_MAINTAINED[0] = _SPLITS_MAINTAINED[0], bool(on)
    return prev


def refresh_weight_splits(device) -> int:
    """Re-split EVERY cached weight operand on ``device`` from the current parameter values in ONE launch
    (hs_weight_split_batch) and mark the cache entries current.  The descriptor table lives on the device and is rebuilt
    only when the set of operands or one of their addresses changes.  Returns the number of operands."""
    device = torch.device(device)
    ents, sig = [], []
    for key, ent in _SPLITS.items():
        wt = ent[0]()
        if wt is None or ent[3].device != device or wt.data_ptr() != ent[2]:
            continue
        ents.append((key, ent, wt))
        sig.append((key, ent[2], ent[3].data_ptr()))
    if not ents:
        return 0
    sig = tuple(sig)
    tab = _SPLIT_TABLES.get(device)
    if tab is None or tab[0] != sig:
        descs = (_SplitDesc * len(ents))()
        tile0 = 0
        for d, (key, ent, wt) in zip(descs, ents):
            _, transposed, c0, K, fmt = key
            N, ld = wt.shape
            rows, ncols = (K, N) if transposed else (N, K)
            d.w, d.out = wt.data_ptr() + 4 * c0, ent[3].data_ptr()
            d.rows, d.cols, d.ld, d.transposed, d.format = rows, ncols, ld, int(transposed), fmt
            d.tiles_x, d.tile0, d.pad = (ncols + 31) // 32, tile0, 0
            tile0 += d.tiles_x * ((rows + 31) // 32)
        host = torch.frombuffer(bytearray(bytes(descs)), dtype=torch.uint8)
        tab = _SPLIT_TABLES[device] = (sig, host.to(device), len(ents), tile0)
    STATS.launch("weight_split_batch", lib.hs_weight_split_batch, ptr(tab[1]), tab[2], tab[3], current_stream())
    for key, ent, wt in ents:
        _SPLITS[key] = (ent[0], wt._version, ent[2], ent[3], _SPLIT_EPOCH[0])
    return len(ents)


def _on_device(t) -> bool:
    return t.is_cuda


def gemm3_ok(x, weight) -> bool:
    """Whether ``F.linear(x, weight)`` is covered by the hand-written GEMM."""
    return bool(_GEMM_MODE == "bf16x3" and _on_device(x) and x.dtype == torch.float32 and weight.dtype == torch.float32
                and weight.dim() == 2 and weight.is_contiguous()
                and lib.hs_gemm3_supported(x.numel() // x.shape[-1], weight.shape[0], weight.shape[1]))


def _dgrad_ok(dy2, weight) -> bool:
    return bool(_GEMM_MODE == "bf16x3" and _on_device(dy2) and weight.is_contiguous() and weight.dtype == torch.float32
                and lib.hs_gemm3_supported(dy2.shape[0], weight.shape[1], weight.shape[0]))


def _gemm3(a2, wsplit, N, bias=None, aux=None, mode=_lib.GEMM_PLAIN, drop=0.0, seed=0, colsum=None, prec=None):
    """hs_gemm3 on a (T, K) activation and a split weight; returns d, or (d, d2) for GEMM_GELU.  ``colsum`` (K floats,
    zero or a running sum) receives the column sums of ``a2`` in the same pass."""
    T, K = a2.shape
    two = mode in (_lib.GEMM_GELU, _lib.GEMM_GELU_C)
    # (compact GELU mode: d is the FP16 derivative tensor g', d2 the activation h)
    d = torch.empty((T, N), device=a2.device, dtype=torch.float16 if mode == _lib.GEMM_GELU_C else torch.float32)
    d2 = torch.empty((T, N), device=a2.device, dtype=torch.float32) if two else None
    prec = _fwd_prec() if prec is None else prec
    STATS.launch("gemm3", lib.hs_gemm3, ptr(a2), ptr(wsplit), ptr(bias), ptr(aux), ptr(d), ptr(d2), ptr(colsum), T, N, K,
                 mode, prec, C.c_float(drop), C.c_uint64(seed), current_stream(), tag=(T, N, K, mode, prec))
    return (d, d2) if two else d


_FUSED_LINEAR_LN = os.environ.get("HEALSWIN_FUSED_LINEAR_LN", "1") == "1"


def _gemm3_ln(a2, wsplit, N, bias, gamma, beta, G, eps, aux=None, save=True):
    """hs_gemm3_ln: y = [aux +] LayerNorm_G(a2 @ W^T + bias) * gamma + beta in the GEMM's epilogue.  Returns
    (y, pre, mean, rstd); the last three (what hs_layernorm_bwd needs) only with ``save``."""
    T, K = a2.shape
    y = torch.empty((T, N), device=a2.device, dtype=torch.float32)
    pre = torch.empty_like(y) if save else None
    mean = torch.empty((T * (N // G),), device=a2.device, dtype=torch.float32) if save else None
    rstd = torch.empty_like(mean) if save else None
    prec = _fwd_prec()
    STATS.launch("gemm3", lib.hs_gemm3_ln, ptr(a2), ptr(wsplit), ptr(bias), ptr(gamma), ptr(beta), ptr(aux), ptr(pre), ptr(y),
                 ptr(mean), ptr(rstd), T, N, K, G, C.c_float(eps), prec, current_stream(),
                 tag=(T, N, K, 10 + (1 if aux is not None else 0) + (2 if save else 0), prec))
    return y, pre, mean, rstd


def _ln_tail_bwd(dy2, pre, mean, rstd, gamma, need_w, need_b, need_lin_bias=False):
    """Backward of the LN epilogue: d(pre) (T, N), d(gamma), d(beta) from dy over the (T N / G, G) view, and -- for G = N,
    on request -- the column sums of d(pre), i.e. the gradient of the linear's bias, from the same pass (the kernel's
    pre-bias path with a zero pre-bias)."""
    G = gamma.numel()
    rows = pre.numel() // G
    dpre = torch.empty_like(pre)
    dw = zeros(G, pre.device) if need_w else None
    db = zeros(G, pre.device) if need_b else None
    lin_bias = need_lin_bias and G == pre.shape[-1]
    pb = zeros(G, pre.device) if lin_bias else None
    dpb = zeros(G, pre.device) if lin_bias else None
    STATS.launch("layernorm_bwd", lib.hs_layernorm_bwd, ptr(dy2), ptr(pre), ptr(pb), ptr(mean), ptr(rstd), ptr(gamma),
                 None, 0, C.c_float(0.0), C.c_uint64(0), ptr(dpre), ptr(dw), ptr(db), ptr(dpb), rows, G,
                 current_stream(), tag=(rows, G))
    return dpre, dw, db, dpb


def _fusable_ln(norm, N) -> bool:
    return bool(_FUSED_LINEAR_LN and isinstance(norm, torch.nn.LayerNorm) and norm.elementwise_affine
                and norm.bias is not None and len(norm.normalized_shape) == 1 and N % norm.normalized_shape[0] == 0)


def linear_ln_supported(x, weight, norm) -> bool:
    """Whether ``linear_ln`` runs as ONE launch (LayerNorm in the GEMM epilogue): affine LayerNorm over G = a divisor of
    the output width, G a multiple of 32 and <= 192."""
    N = weight.shape[0]
    return bool(gemm3_ok(x, weight) and _fusable_ln(norm, N)
                and lib.hs_gemm3_ln_supported(x.numel() // x.shape[-1], N, weight.shape[1], norm.normalized_shape[0]))


class _LinearLnFn(torch.autograd.Function):
    """``[residual +] LayerNorm_G(F.linear(x, weight, bias))`` with the normalisation (over groups of G = gamma.numel()
    output columns), the affine and the residual add in the GEMM's epilogue (hs_gemm3_ln).  Backward: hs_layernorm_bwd on
    the saved pre-norm tensor, then the linear's input / weight / bias gradients as in _LinearFn."""

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, residual, eps):
        ctx.set_materialize_grads(False)
        N, K = weight.shape
        x2 = _f32c(x).reshape(-1, K)
        res2 = _f32c(residual).reshape(-1, N) if residual is not None else None
        save = any(ctx.needs_input_grad[:5])
        y, pre, mean, rstd = _gemm3_ln(x2, split_weight(weight), N, _f32c(bias), _f32c(gamma), _f32c(beta), gamma.numel(),
                                       eps, res2, save)
        if save:
            ctx.save_for_backward(x2, weight, gamma, pre, mean, rstd)
        ctx.meta = (x.shape, bias is not None, residual is not None)
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        xshape, has_bias, has_res = ctx.meta
        if dy is None:
            return (None,) * 7
        x2, weight, gamma, pre, mean, rstd = ctx.saved_tensors
        N, K = weight.shape
        dy2 = _f32c(dy).reshape(-1, N)
        need_b = has_bias and ctx.needs_input_grad[2]
        dpre, dgamma, dbeta, db = _ln_tail_bwd(dy2, pre, mean, rstd, _f32c(gamma), ctx.needs_input_grad[3],
                                               ctx.needs_input_grad[4], need_b)
        dx = dw = None
        need_b = need_b and db is None
        b_in_wgrad = need_b and ctx.needs_input_grad[1] and _wgrad_fuses_bias(dpre, x2)
        if ctx.needs_input_grad[0]:
            if need_b and not b_in_wgrad:
                dx, db = _dgrad(dpre, weight, None, xshape, want_bias_grad=True)
            else:
                dx = _dgrad(dpre, weight, None, xshape)
        if ctx.needs_input_grad[1]:
            dw, db2 = _wgrad(dpre, x2, need_b and db is None)
            db = db if db is not None else db2
        elif need_b and db is None:
            db = dpre.sum(0)
        return dx, dw, db, dgamma, dbeta, (dy if has_res else None), None


def linear_ln(x, weight, bias, norm, residual=None):
    """``[residual +] norm(F.linear(x, weight, bias).view(..., N // G, G)).view(..., N)`` for an affine LayerNorm over
    G = norm.normalized_shape[0] columns: one launch where ``linear_ln_supported``, else the linear and the LayerNorm
    as two launches.  The caller reshapes the result (PatchExpand: (B, N, 2C) -> (B, 4N, C/2))."""
    N = weight.shape[0]
    if linear_ln_supported(x, weight, norm) and (residual is None or residual.is_cuda):
        return _LinearLnFn.apply(x, weight, bias, norm.weight, norm.bias, residual, float(norm.eps))
    G = norm.normalized_shape[0]
    y = linear(x, weight, bias)
    y = layer_norm(y.reshape(*y.shape[:-1], N // G, G), norm).reshape(y.shape)
    return y if residual is None else residual + y


def ln_linear_supported(x, norm, weight) -> bool:
    """Whether ``ln_linear`` runs as one launch: affine LayerNorm over the whole input row (a multiple of 32, >= 160
    wide -- PatchMerging's 4C) feeding a Linear the hand-written GEMM covers."""
    K = weight.shape[1]
    return bool(_FUSED_LINEAR_LN and gemm3_ok(x, weight) and _fusable_norm(norm, x)
                and lib.hs_gemm3_lnin_supported(x.numel() // K, weight.shape[0], K))


class _LnLinearFn(torch.autograd.Function):
    """``F.linear(LayerNorm(x), weight)`` (no bias) with the normalisation folded into the GEMM (hs_gemm3_lnin): the
    product runs on the raw rows with gamma folded into the weight, the row statistics come out of the GEMM's own operand
    pass and are applied in its epilogue; the normalised (T, K) tensor is neither written nor kept for the backward (it is
    rebuilt there by one LayerNorm pass for the weight gradient)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, weight, eps):
        ctx.set_materialize_grads(False)
        N, K = weight.shape
        x2 = _f32c(x).reshape(-1, K)
        T = x2.shape[0]
        w = _f32c(weight.detach())
        g, b = _f32c(gamma.detach()), _f32c(beta.detach())
        wg = w * g                       # W diag(gamma)
        wsum = wg.sum(1)                 # s = W' 1
        b0 = (w * b).sum(1)              # W beta (elementwise: no library GEMV for a parameter-sized product)
        save = any(ctx.needs_input_grad[:4])
        d = torch.empty((T, N), device=x2.device, dtype=torch.float32)
        mean = torch.empty((T,), device=x2.device, dtype=torch.float32) if save else None
        rstd = torch.empty_like(mean) if save else None
        prec = _fwd_prec()
        STATS.launch("gemm3", lib.hs_gemm3_lnin, ptr(x2), ptr(split_weight(wg)), ptr(wsum), ptr(b0), ptr(d), ptr(mean),
                     ptr(rstd), T, N, K, C.c_float(eps), prec, current_stream(), tag=(T, N, K, 20, prec))
        if save:
            ctx.save_for_backward(x2, gamma, beta, weight, mean, rstd)
        ctx.meta = (x.shape, float(eps))
        return d.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        if dy is None:
            return (None,) * 5
        x2, gamma, beta, weight, mean, rstd = ctx.saved_tensors
        xshape, eps = ctx.meta
        N, K = weight.shape
        T = x2.shape[0]
        dy2 = _f32c(dy).reshape(T, N)
        g = _f32c(gamma)
        dx = dw = dgamma = dbeta = None
        if ctx.needs_input_grad[3]:  # dW = dy^T LayerNorm(x): the normalised rows are rebuilt (one pass) for this product
            xn = torch.empty_like(x2)
            STATS.launch("layernorm_fwd", lib.hs_layernorm_fwd, ptr(x2), None, None, ptr(g), ptr(_f32c(beta)), None, 0,
                         C.c_float(0.0), C.c_uint64(0), ptr(xn), None, None, T, K, C.c_float(eps), current_stream(),
                         tag=(T, K, 0))
            dw, _ = _wgrad(dy2, xn, False)
            del xn
        if any(ctx.needs_input_grad[:3]):
            dxn = _dgrad(dy2, weight, None, (T, K))
            dx = torch.empty_like(x2)
            dgamma = zeros(K, x2.device) if ctx.needs_input_grad[1] else None
            dbeta = zeros(K, x2.device) if ctx.needs_input_grad[2] else None
            STATS.launch("layernorm_bwd", lib.hs_layernorm_bwd, ptr(dxn), ptr(x2), None, ptr(mean), ptr(rstd), ptr(g), None, 0,
                         C.c_float(0.0), C.c_uint64(0), ptr(dx), ptr(dgamma), ptr(dbeta), None, T, K, current_stream(),
                         tag=(T, K))
            dx = dx.view(xshape)
        return dx, dgamma, dbeta, dw, None


def ln_linear(x, norm, weight, bias=None):
    """``F.linear(norm(x), weight, bias)`` for an affine LayerNorm over the last dim: one launch where
    ``ln_linear_supported`` (and no bias), else the LayerNorm and the linear as two launches."""
    if bias is None and ln_linear_supported(x, norm, weight):
        return _LnLinearFn.apply(x, norm.weight, norm.bias, weight, float(norm.eps))
    return linear(layer_norm(x, norm), weight, bias)


def _gemm_fwd(x2, weight, bias):
    """x2 (T, K) @ weight (N, K)^T + bias."""
    if gemm3_ok(x2, weight):
        return _gemm3(x2, split_weight(weight), weight.shape[0], _f32c(bias))
    return torch.nn.functional.linear(x2, weight, bias)


def _dgrad(dy2, weight, d_pass, xshape, want_bias_grad=False):
    """dy2 @ weight (+ the gradient that reached the forked shortcut output), shaped like the input.  The shortcut
    gradient is added in the GEMM's epilogue (GEMM_ADD): no accumulation pass over the activation.  With
    ``want_bias_grad`` the column sums of dy2 -- the bias gradient of this linear -- are taken in the same pass over dy2 (the
    GEMM's operand converters see every element anyway); returns (dx, db)."""
    N, K = weight.shape
    c = None if d_pass is None else _f32c(d_pass).reshape(-1, K)
    db = None
    if _dgrad_ok(dy2, weight):
        if want_bias_grad:
            db = zeros((N,), dy2.device)
        prec = _dgrad_prec(dy2.shape[0], K, N)
        dx = _gemm3(dy2, split_weight(weight, transposed=True, prec=prec), K, None, c,
                    _lib.GEMM_PLAIN if c is None else _lib.GEMM_ADD, colsum=db, prec=prec)
    else:
        dx = dy2 @ weight if c is None else torch.addmm(c, dy2, weight)
        if want_bias_grad:
            db = dy2.sum(0)
    return (dx.view(xshape), db) if want_bias_grad else dx.view(xshape)


def _wgrad_fuses_bias(dy2, x2) -> bool:
    return bool(_CUSTOM_WGRAD and _on_device(dy2)
                and lib.hs_linear_wgrad_supported(dy2.shape[0], dy2.shape[1], x2.shape[1]) == 2)


def _tc_flags() -> int:
    """Flags for the TF32 kernels: HEALSWIN_NO_TRUNC_COMP=1 switches the statistical truncation compensation off in ALL of
    them (attention, weight gradient, stage-0/1 MLP gradient) -- for operands that are exactly representable in TF32 the
    compensation is a +3.5e-4 .. 7e-4 bias rather than a correction."""
    return _lib.ATTN_NO_TRUNC_COMP if os.environ.get("HEALSWIN_NO_TRUNC_COMP") == "1" else 0


def _wgrad(dy2, x2, need_bias):
    """(dW, db or None) of a linear: the token-split tensor-core kernel where it covers the shape, else the library."""
    T, N = dy2.shape
    K = x2.shape[1]
    db = None
    cover = lib.hs_linear_wgrad_supported(T, N, K) if (_CUSTOM_WGRAD and _on_device(dy2)) else 0
    if cover:
        dw = zeros((N, K), x2.device)
        if need_bias and cover == 2:  # bias gradient in the same pass over dy
            db = zeros((N,), x2.device)
        STATS.launch("linear_wgrad", lib.hs_linear_wgrad, ptr(dy2), ptr(x2), ptr(dw), ptr(db), T, N, K, _tc_flags(),
                     current_stream(), tag=(T, N, K))
    else:
        # shapes outside the kernel (fewer than 4096 tokens, feature counts that are not multiples of 4): the library GEMM, in the same
        # arithmetic as the kernel (TF32 operands, fp32 accumulation) whatever torch's global switch says -- as fp32
        # these run as SIMT sgemm kernels, 11 ms per step of the N_side=256 network
        # (with the kernel switched off -- HEALSWIN_CUSTOM_WGRAD=0, the exact-fp32 tests -- torch's switch decides)
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = prev or _CUSTOM_WGRAD
        try:
            dw = dy2.t() @ x2
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev
    if need_bias and db is None:
        db = dy2.sum(0)
    return dw, db


class _LinearFn(torch.autograd.Function):
    """F.linear on this library's GEMMs: forward and input gradient on the bf16x3 kernel (bias in the epilogue), weight
    (and bias) gradient on the token-split kernel.  With ``fork`` the input is returned as a second output -- the
    shortcut of a residual block whose branch starts with this linear -- and the shortcut's gradient is added in the
    input-gradient GEMM's epilogue instead of a separate accumulation pass over the activation."""

    @staticmethod
    def forward(ctx, x, weight, bias, fork):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        ctx.set_materialize_grads(False)
        y = _gemm_fwd(_f32c(x).reshape(-1, weight.shape[1]), weight, _f32c(bias)).view(*x.shape[:-1], weight.shape[0])
        return (y, x.view_as(x)) if fork else y

    @staticmethod
    def backward(ctx, dy, d_pass=None):
        x, weight = ctx.saved_tensors
        N, K = weight.shape
        dx = dw = db = None
        if dy is None:
            return d_pass, None, None, None
        dy2 = _f32c(dy).reshape(-1, N)
        x2 = _f32c(x).reshape(-1, K)
        need_b = ctx.has_bias and ctx.needs_input_grad[2]
        # the bias gradient rides along with the weight-gradient kernel where that covers it, else with the input-gradient
        # GEMM (both read dy anyway); a separate reduction over dy only when neither runs
        b_in_wgrad = need_b and ctx.needs_input_grad[1] and _wgrad_fuses_bias(dy2, x2)
        if ctx.needs_input_grad[0]:
            if need_b and not b_in_wgrad:
                dx, db = _dgrad(dy2, weight, d_pass, x.shape, want_bias_grad=True)
            else:
                dx = _dgrad(dy2, weight, d_pass, x.shape)
        if ctx.needs_input_grad[1]:
            dw, db2 = _wgrad(dy2, x2, need_b and db is None)
            db = db if db is not None else db2
        elif need_b and db is None:
            db = dy2.sum(0)
        return dx, dw, db, None


class _MlpFn(torch.autograd.Function):
    """``fc2(dropout(GELU(fc1(x) + b1)))`` WITHOUT fc2's bias, as one autograd node.  Forward: fc1 with the bias add,
    GELU and dropout in its epilogue (z and h each written once, hs_gemm3 GEMM_GELU), then fc2.  Backward: the (T, 4C)
    hidden gradient is never materialised -- d(fc1 output) = (dy @ W2) * GELU'(z + b1) * mask comes out of one GEMM
    epilogue (csrc/hs_mlp_dgrad_tc.cu where it covers the shape, else hs_gemm3 GEMM_GELU_GRAD); both weight gradients
    (and fc1's bias gradient) from the token-split wgrad kernel."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, drop, seed, fork, b2=None, gamma=None, beta=None, eps=0.0):
        ctx.set_materialize_grads(False)
        K = x.shape[-1]
        x2 = _f32c(x).reshape(-1, K)
        # z: the bias-free fc1 output (fp32), or -- compact form -- the FP16 tensor g' = GELU'(z + b1) * dropmask, which is
        # all the backward needs from it at half the bytes.  Used where it pays (same-box A/B of the N_side=256 step): the
        # tensor-bound stages (C >= 384), whose backward GEMM then multiplies by g' instead of bringing z in by TMA and
        # evaluating GELU' (0.27 against 0.38 ms per stage-2 launch).  At stages 0-1 the launches are bound by their
        # epilogues' issue slots and the per-row FP16 accesses cost more than the bytes they save (HEALSWIN_MLP_COMPACT=all).
        compact = (_MLP_COMPACT != "0" and w1.shape[0] % 32 == 0 and (_MLP_COMPACT == "all" or w1.shape[1] > 256))
        z, h = _gemm3(x2, split_weight(w1), w1.shape[0], _f32c(b1), None,
                      _lib.GEMM_GELU_C if compact else _lib.GEMM_GELU, drop, seed)
        ctx.drop = (float(drop), int(seed))
        ctx.xshape = x.shape
        ctx.ln = gamma is not None
        if ctx.ln:
            # the whole second half of a v2 block: x + norm2(fc2(h) + b2) -- bias, LayerNorm and the residual add in
            # fc2's epilogue (swin_hp_transformer.py:336-338)
            y, pre, mean, rstd = _gemm3_ln(h, split_weight(w2), w2.shape[0], _f32c(b2), _f32c(gamma), _f32c(beta),
                                           gamma.numel(), eps, x2, True)
            ctx.save_for_backward(x2, w1, b1, w2, z, h, gamma, pre, mean, rstd)
            ctx.has_b2 = b2 is not None
            return y.view(*x.shape[:-1], w2.shape[0])
        y = _gemm3(h, split_weight(w2), w2.shape[0])
        ctx.save_for_backward(x2, w1, b1, w2, z, h)
        y = y.view(*x.shape[:-1], w2.shape[0])
        return (y, x.view_as(x)) if fork else y

    @staticmethod
    def backward(ctx, dy, d_pass=None):
        if dy is None:
            return (d_pass,) + (None,) * 10
        db2 = dgamma = dbeta = None
        if ctx.ln:
            x2, w1, b1, w2, z, h, gamma, pre, mean, rstd = ctx.saved_tensors
            d_pass = dy  # the residual shortcut: added in fc1's input-gradient epilogue
            need_b2 = ctx.has_b2 and ctx.needs_input_grad[7]
            dy2, dgamma, dbeta, db2 = _ln_tail_bwd(_f32c(dy).reshape(pre.shape), pre, mean, rstd, _f32c(gamma),
                                                   ctx.needs_input_grad[8], ctx.needs_input_grad[9], need_b2)
            T, K = x2.shape
            J, Cout = w1.shape[0], w2.shape[0]
            dw2, db2b = _wgrad(dy2, h, need_b2 and db2 is None)
            db2 = db2 if db2 is not None else db2b
        else:
            x2, w1, b1, w2, z, h = ctx.saved_tensors
            T, K = x2.shape
            J, Cout = w1.shape[0], w2.shape[0]
            dy2 = _f32c(dy).reshape(T, Cout)
            dw2, _ = _wgrad(dy2, h, False)
        # stage 0 (HBM-bound): the dedicated TF32 kernel; from stage 1 on, where one TF32 MMA per product is what
        # _dgrad_prec picks anyway, the general GEMM with the GELU' epilogue is faster (C = 192: 0.54 against 0.65 ms,
        # scripts/mlp_dgrad_s1.py)
        compact = z.dtype == torch.float16
        if (_TF32_MLP_DGRAD and lib.hs_mlp_dgrad_gelu_supported(T, Cout, J)
                and not (_GEMM_PRECISION != "bf16" and _dgrad_tensor_bound(J, Cout))):
            dz = torch.empty((T, J), device=z.device, dtype=torch.float32)
            STATS.launch("mlp_dgrad_gelu", lib.hs_mlp_dgrad_gelu, ptr(dy2), ptr(w2), ptr(z), ptr(b1),
                         C.c_float(ctx.drop[0]), C.c_uint64(ctx.drop[1]), ptr(dz), T, Cout, J,
                         _tc_flags() | (_lib.MLP_GRAD16 if compact else 0), current_stream(),
                         tag=(T, Cout, J, int(compact)))
        elif compact:
            prec = _dgrad_prec(T, J, Cout)
            dz = _gemm3(dy2, split_weight(w2, transposed=True, prec=prec), J, None, z, _lib.GEMM_GELU_GRAD_C, prec=prec)
        else:
            prec = _dgrad_prec(T, J, Cout)
            dz = _gemm3(dy2, split_weight(w2, transposed=True, prec=prec), J, _f32c(b1), z, _lib.GEMM_GELU_GRAD, *ctx.drop,
                        prec=prec)
        dx = db1 = None
        if ctx.needs_input_grad[0]:
            if _wgrad_fuses_bias(dz, x2):
                dx = _dgrad(dz, w1, d_pass, ctx.xshape)
            else:  # fc1's bias gradient = column sums of dz, taken by the input-gradient GEMM in its pass over dz
                dx, db1 = _dgrad(dz, w1, d_pass, ctx.xshape, want_bias_grad=True)
        dw1, db1b = _wgrad(dz, x2, db1 is None)
        db1 = db1 if db1 is not None else db1b
        return dx, dw1, db1, dw2, None, None, None, db2, dgamma, dbeta, None


def mlp_supported(x, fc1, fc2):
    """Whether ``mlp_core`` covers this MLP (both linears inside the hand-written GEMM)."""
    if not (_FUSED_MLP and fc1.bias is not None and gemm3_ok(x, fc1.weight)):
        return False
    T = x.numel() // x.shape[-1]
    w2 = fc2.weight
    return bool(w2.is_contiguous() and w2.dtype == torch.float32
                and lib.hs_gemm3_supported(T, w2.shape[0], w2.shape[1])
                and lib.hs_gemm3_supported(T, w2.shape[1], w2.shape[0]))


def mlp_core(x, fc1, fc2, drop=0.0, seed=None, fork=False):
    """``F.linear(dropout(GELU(fc1(x))), fc2.weight)`` (no fc2 bias) through the fused node; check ``mlp_supported``.
    ``fork``: also return the input as the residual shortcut (see ``linear``)."""
    drop = float(drop)
    seed = _pick_seed(seed, drop)
    return _MlpFn.apply(x, fc1.weight, fc1.bias, fc2.weight, drop, int(seed or 0), bool(fork))


def mlp_ln_supported(x, fc1, fc2, norm) -> bool:
    """Whether ``mlp_ln`` covers ``x + norm(mlp(x))`` as one autograd node with the LayerNorm in fc2's epilogue."""
    if not (mlp_supported(x, fc1, fc2) and _fusable_ln(norm, fc2.weight.shape[0])):
        return False
    N, K = fc2.weight.shape
    return bool(norm.normalized_shape[0] == N == x.shape[-1]
                and lib.hs_gemm3_ln_supported(x.numel() // x.shape[-1], N, K, N))


def mlp_ln(x, fc1, fc2, norm, drop=0.0, seed=None):
    """``x + norm(fc2(dropout(GELU(fc1(x)))))`` -- the MLP half of a v2-placement block (swin_hp_transformer.py:336-338
    with drop_path = identity and no dropout after fc2) in two launches; check ``mlp_ln_supported``."""
    drop = float(drop)
    seed = _pick_seed(seed, drop)
    return _MlpFn.apply(x, fc1.weight, fc1.bias, fc2.weight, drop, int(seed or 0), False, fc2.bias, norm.weight, norm.bias,
                        float(norm.eps))


class _CatLinearFn(torch.autograd.Function):
    """``F.linear(torch.cat([x1, x2], -1), weight, bias)`` without the concatenation (UnetDecoder skip connections,
    swin_hp_transformer.py:772-775): two accumulating GEMMs over the two column blocks of the weight -- the second adds the
    first's result in its epilogue (GEMM_ADD) -- and in the backward two input-gradient GEMMs writing dx1 / dx2 directly
    (no split copies of a (T, 2C) gradient) and two weight-gradient launches."""

    @staticmethod
    def forward(ctx, x1, x2, weight, bias):
        ctx.set_materialize_grads(False)
        K1, K2 = x1.shape[-1], x2.shape[-1]
        N = weight.shape[0]
        a1, a2 = _f32c(x1).reshape(-1, K1), _f32c(x2).reshape(-1, K2)
        part = _gemm3(a1, split_weight(weight, cols=(0, K1)), N, _f32c(bias))
        y = _gemm3(a2, split_weight(weight, cols=(K1, K2)), N, None, part, _lib.GEMM_ADD)
        ctx.save_for_backward(a1, a2, weight)
        ctx.meta = (x1.shape, x2.shape, bias is not None)
        return y.view(*x1.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        a1, a2, weight = ctx.saved_tensors
        s1, s2, has_bias = ctx.meta
        if dy is None:
            return None, None, None, None
        N = weight.shape[0]
        K1, K2 = a1.shape[1], a2.shape[1]
        dy2 = _f32c(dy).reshape(-1, N)
        need_b = has_bias and ctx.needs_input_grad[3]
        db = zeros((N,), dy2.device) if need_b else None
        dx1 = dx2 = dw = None
        if ctx.needs_input_grad[0]:
            p1 = _dgrad_prec(dy2.shape[0], K1, N)
            dx1 = _gemm3(dy2, split_weight(weight, transposed=True, cols=(0, K1), prec=p1), K1, colsum=db, prec=p1).view(s1)
        if ctx.needs_input_grad[1]:
            p2 = _dgrad_prec(dy2.shape[0], K2, N)
            dx2 = _gemm3(dy2, split_weight(weight, transposed=True, cols=(K1, K2), prec=p2), K2,
                         colsum=None if ctx.needs_input_grad[0] else db, prec=p2).view(s2)
        if need_b and not (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]):
            db = dy2.sum(0)
        if ctx.needs_input_grad[2]:
            dw = torch.cat([_wgrad(dy2, a1, False)[0], _wgrad(dy2, a2, False)[0]], dim=1)
        return dx1, dx2, dw, db


def cat_linear(x1, x2, weight, bias=None):
    """``F.linear(torch.cat([x1, x2], -1), weight, bias)``; the concatenated tensor is never built where the hand-written
    GEMM covers the shapes."""
    K1, K2 = x1.shape[-1], x2.shape[-1]
    T = x1.numel() // K1
    if (_GEMM_MODE == "bf16x3" and _on_device(x1) and x1.dtype == torch.float32 and x2.dtype == torch.float32
            and weight.dim() == 2 and weight.is_contiguous() and weight.shape[1] == K1 + K2 and K1 % 4 == 0
            and x1.shape[:-1] == x2.shape[:-1]
            and lib.hs_gemm3_supported(T, weight.shape[0], K1) and lib.hs_gemm3_supported(T, weight.shape[0], K2)
            and lib.hs_gemm3_supported(T, K1, weight.shape[0]) and lib.hs_gemm3_supported(T, K2, weight.shape[0])):
        return _CatLinearFn.apply(x1, x2, weight, bias)
    return linear(torch.cat([x1, x2], -1), weight, bias)


def linear(x, weight, bias=None, fork=False):
    """``F.linear`` on the hand-written GEMMs where the shape is covered (feature dimensions multiples of 4), else
    through torch.  ``fork=True`` returns ``(y, shortcut)`` where ``shortcut`` is ``x`` for the residual path of a block
    whose branch starts with this linear: its gradient is added inside the input-gradient GEMM."""
    if gemm3_ok(x, weight):
        return _LinearFn.apply(x, weight, bias, bool(fork))
    y = torch.nn.functional.linear(x, weight, bias)
    return (y, x) if fork else y
