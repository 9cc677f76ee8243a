"""Drop-in mirror of heal_swin/models_torch/swin_hp_transformer.py on the B200 engine.

Same class names, constructor signatures, parameter / buffer names and shapes as the reference
(SURVEY.md 8b), so Lightning wrappers, optimizers, DDP bucketing and checkpoints keep working.
What differs is what ``forward`` launches: the shift -> window-partition -> attention ->
window-reverse -> shift-back chain is one kernel with the HEALPix permutation folded into its
loads/stores (csrc/), fed by compact int32 / uint8 tables instead of the (nW, ws, ws) fp32 mask.
"""
import math
from dataclasses import dataclass, field
from typing import List, Literal, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.utils.checkpoint as checkpoint

from .. import hp_index, ops
from ..data_spec import DataSpec
from . import hp_shifting
from .hp_windowing import window_partition, window_reverse, get_nest_win_idcs  # noqa: F401 (re-export)


class DropPath(nn.Module):
    """Stochastic depth (timm 0.4.12 ``DropPath`` as used at swin_hp_transformer.py:261)."""

    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def sample_scale(self, x):
        """Per-sample factor (0 or 1 / keep) of one stochastic-depth draw, or None when inactive."""
        if not self.training or not self.drop_prob:
            return None
        keep = 1.0 - self.drop_prob
        return x.new_empty((x.shape[0],)).bernoulli_(keep) / keep

    def forward(self, x):
        scale = self.sample_scale(x)
        return x if scale is None else x * scale.view((x.shape[0],) + (1,) * (x.ndim - 1))


class Mlp(nn.Module):
    """swin_hp_transformer.py:21-44"""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def _fusable(self):
        return (isinstance(self.act, nn.GELU) and getattr(self.act, "approximate", "none") == "none"
                and self.fc1.bias is not None)

    def _drop_p(self):
        return self.drop.p if self.training else 0.0

    def forward_split(self, x):
        """(fc2 output WITHOUT its bias and WITHOUT the trailing dropout, that bias or None, that dropout's p, the input
        as residual shortcut): the caller applies bias and dropout (fused into the LayerNorm that follows in the v2
        placement) and adds the shortcut, whose gradient is folded into fc1's input-gradient GEMM.  fc1's bias add, the
        GELU, the first dropout and (backward) fc1's bias gradient are one kernel."""
        if not self._fusable() or self.fc2.bias is None:
            return self.forward(x), None, 0.0, x
        y, shortcut = self._core(x, fork=True)
        return y, self.fc2.bias, self._drop_p(), shortcut

    def ln_fusable(self, x, norm):
        """Whether ``x + norm(self(x))`` runs with the LayerNorm and the residual add in fc2's epilogue."""
        return self._fusable() and self._drop_p() == 0.0 and ops.mlp_ln_supported(x, self.fc1, self.fc2, norm)

    def forward_ln(self, x, norm):
        """``x + norm(self(x))``: fc1 + GELU, then fc2 + bias + LayerNorm + residual, two launches (``ln_fusable``)."""
        return ops.mlp_ln(x, self.fc1, self.fc2, norm)

    def _core(self, x, fork=False):
        """fc2(drop(GELU(fc1(x)))) without fc2's bias: one fused autograd node when the shapes allow it."""
        if ops.mlp_supported(x, self.fc1, self.fc2):
            return ops.mlp_core(x, self.fc1, self.fc2, drop=self._drop_p(), fork=fork)
        z = ops.linear(x, self.fc1.weight, fork=fork)
        z, shortcut = z if fork else (z, None)
        y = ops.linear(ops.bias_gelu(z, self.fc1.bias, drop=self._drop_p()), self.fc2.weight)
        return (y, shortcut) if fork else y

    def forward(self, x):
        if not self._fusable():
            h = self.drop(self.act(ops.linear(x, self.fc1.weight, self.fc1.bias)))
            return self.drop(ops.linear(h, self.fc2.weight, self.fc2.bias))
        y = self._core(x)
        return self.drop(y if self.fc2.bias is None else y + self.fc2.bias)


class WindowAttention(nn.Module):
    """Window multi-head self attention with relative position bias   [swin_hp_transformer.py:47-174]

    ``window_size`` is the flat number of tokens per window (a power of 2; a power of 4 with
    ``rel_pos_bias="flat"``).
    """

    def __init__(self, dim, window_size, num_heads, rel_pos_bias=None, qkv_bias=True, qk_scale=None,
                 attn_drop=0.0, proj_drop=0.0, use_cos_attn=False):
        super().__init__()
        self.dim = dim
        self.window_size = window_size
        self.num_heads = num_heads
        self.use_cos_attn = use_cos_attn
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        self.rel_pos_bias = rel_pos_bias

        if use_cos_attn:  # :84-87
            self.logit_scale = nn.Parameter(torch.log(10 * torch.ones((num_heads, 1, 1))), requires_grad=True)
        if rel_pos_bias == "flat":  # :89-114; zero-initialised like the reference (:121 is commented out)
            side = 2 * window_size**0.5 - 1
            self.relative_position_bias_table = nn.Parameter(torch.zeros((int(side * side), num_heads)))
            index = hp_index.rel_pos_index(window_size)
            self.register_buffer("relative_position_index", index)
            self.register_buffer("_hs_rel_index", index.to(torch.int32).reshape(-1).contiguous(), persistent=False)

        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.softmax = nn.Softmax(dim=-1)

    def _core(self, qkv, window_size, src, groups, dense_mask):
        # nn.Dropout(attn_drop) on the softmax output (:167-169) runs inside the kernel (counter-based mask)
        drop_p = self.attn_drop.p if self.training else 0.0
        table = self.relative_position_bias_table if self.rel_pos_bias is not None else None
        rel_index = self._hs_rel_index if table is not None else None
        logit_scale = self.logit_scale if self.use_cos_attn else None
        return ops.window_attention_core(qkv, table, logit_scale, src, groups, dense_mask, rel_index,
                                         self.scale, self.num_heads, window_size, self.use_cos_attn, attn_drop=drop_p)

    def forward_tokens(self, x, window_size, src=None, groups=None):
        """Fused path used by SwinTransformerBlock: x is the (B, N, C) token tensor in its natural
        (unshifted) order; ``src``/``groups`` are the block's shift tables (None = no shift)."""
        out = self._core(ops.linear(x, self.qkv.weight, self.qkv.bias), window_size, src, groups, None)
        return self.proj_drop(ops.linear(out, self.proj.weight, self.proj.bias))

    def forward_tokens_split(self, x, window_size, src=None, groups=None, defer_proj=False):
        """As forward_tokens, but returns (proj output WITHOUT its bias and WITHOUT proj_drop, that bias or None, the
        dropout probability still to be applied, the input as residual shortcut) so that the caller can fuse bias and
        dropout into the LayerNorm that follows (v2 norm placement); the shortcut's gradient is folded into the qkv
        input-gradient GEMM.  With ``defer_proj`` the first item is the attention output BEFORE ``self.proj``: the caller
        runs the projection itself, with the LayerNorm and the residual add in its epilogue (``_residual_tail``)."""
        if self.proj.bias is None:
            return self.forward_tokens(x, window_size, src, groups), None, 0.0, x
        qkv, shortcut = ops.linear(x, self.qkv.weight, self.qkv.bias, fork=True)
        out = self._core(qkv, window_size, src, groups, None)
        if not defer_proj:
            out = ops.linear(out, self.proj.weight)
        return out, self.proj.bias, (self.proj_drop.p if self.training else 0.0), shortcut

    def forward(self, x, mask=None):
        """x: (num_windows*B, N, C); mask: (num_windows, N, N) additive or None   [:124-174]"""
        B_, n, C = x.shape
        qkv = ops.linear(x, self.qkv.weight, self.qkv.bias)
        if mask is not None:
            nW = mask.shape[0]
            assert B_ % nW == 0
            qkv = qkv.reshape(B_ // nW, nW * n, 3 * C)
        out = self._core(qkv, n, None, None, mask)
        return self.proj_drop(ops.linear(out.reshape(B_, n, C), self.proj.weight, self.proj.bias))

    def extra_repr(self) -> str:
        return f"dim={self.dim}, window_size={self.window_size}, num_heads={self.num_heads}"


class SwinTransformerBlock(nn.Module):
    """Swin block on the HEALPix grid   [swin_hp_transformer.py:193-340]"""

    def __init__(self, dim, input_resolution, base_pix, num_heads, window_size=4, shift_size=0,
                 shift_strategy="nest_roll", rel_pos_bias=None, mlp_ratio=4.0, qkv_bias=True, qk_scale=None,
                 drop=0.0, attn_drop=0.0, drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm,
                 use_v2_norm_placement=False, use_cos_attn=False):
        super().__init__()
        self.dim = dim
        self.input_resolution = input_resolution
        self.num_heads = num_heads
        self.window_size = window_size
        self.shift_size = shift_size
        self.mlp_ratio = mlp_ratio
        self.use_v2_norm_placement = use_v2_norm_placement
        if self.input_resolution <= self.window_size:  # :243-246
            self.shift_size = 0
            self.window_size = self.input_resolution

        self.norm1 = norm_layer(dim)
        # NB (:249-251) the attention module keeps the configured window size for its bias table
        self.attn = WindowAttention(dim, window_size=window_size, num_heads=num_heads, rel_pos_bias=rel_pos_bias,
                                    qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop,
                                    use_cos_attn=use_cos_attn)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

        nside = hp_index.nside_of(input_resolution, base_pix)  # :271-274
        if self.shift_size > 0:  # :277-304
            if shift_strategy == "nest_roll":
                self.shifter = hp_shifting.NestRollShift(self.shift_size, self.input_resolution, self.window_size)
            elif shift_strategy == "nest_grid_shift":
                self.shifter = hp_shifting.NestGridShift(nside, base_pix, self.window_size)
            elif shift_strategy == "ring_shift":
                self.shifter = hp_shifting.RingShift(nside, base_pix, self.window_size, self.shift_size)
            else:
                raise KeyError(shift_strategy)
        else:
            self.shifter = hp_shifting.NoShift()

        # checkpoint-compatible buffer (:306-308); the kernels read the compact tables below instead
        self.register_buffer("attn_mask", self.shifter.get_mask())
        if self.shifter.shift_idcs is not None:
            self.register_buffer("_hs_src", self.shifter.shift_idcs.to(torch.int32).contiguous(), persistent=False)
            self.register_buffer("_hs_groups", self.shifter.groups.to(torch.uint8).contiguous(), persistent=False)
        else:
            self._hs_src = None
            self._hs_groups = None

    def forward(self, x):
        B, N, C = x.shape
        assert N == self.input_resolution, f"got {N} tokens, block was built for {self.input_resolution}"
        shortcut = x
        if not self.use_v2_norm_placement:
            x = ops.layer_norm(x, self.norm1)
        # shift + partition + W-MSA/SW-MSA + reverse + shift back, one kernel chain   [:319-330]
        if self.use_v2_norm_placement:
            defer = self.attn.proj.bias is not None
            x, pre_bias, pdrop, shortcut = self.attn.forward_tokens_split(x, self.window_size, self._hs_src,
                                                                          self._hs_groups, defer_proj=defer)
            return _residual_tail(self, shortcut, x, pre_bias, pdrop, proj=self.attn.proj if defer else None)
        x, pre_bias, pdrop = self.attn.forward_tokens(x, self.window_size, self._hs_src, self._hs_groups), None, 0.0
        return _residual_tail(self, shortcut, x, pre_bias, pdrop)

    def extra_repr(self) -> str:
        return (f"dim={self.dim}, input_resolution={self.input_resolution}, num_heads={self.num_heads},"
                f" window_size={self.window_size}, shift_size={self.shift_size}, mlp_ratio={self.mlp_ratio}")


def _residual_tail(blk, shortcut, x, pre_bias=None, pre_drop=0.0, proj=None):
    """The two residual branches after the attention   [swin_hp_transformer.py:333-338 / swin_transformer.py:394-401].
    ``x`` is the attention branch, ``pre_bias`` / ``pre_drop`` the not-yet-applied bias and dropout of its output
    projection (``proj``: that projection itself, when the caller deferred it).  In the v2 placement every
    ``shortcut + drop_path(norm(drop(branch + bias)))`` is ONE fused launch (bias, dropout, LayerNorm, per-sample
    stochastic-depth scale, residual add) -- and without dropout / stochastic depth, where the GEMM covers the width
    (LayerNorm over <= 192 channels), it is not a launch at all but the epilogue of the proj / fc2 GEMM."""
    dp = blk.drop_path
    scale_of = (lambda t: dp.sample_scale(t)) if isinstance(dp, DropPath) else (lambda t: None)
    if blk.use_v2_norm_placement:
        plain = not (isinstance(dp, DropPath) and dp.training and dp.drop_prob)  # no stochastic-depth scale to apply
        if proj is not None and plain and pre_drop == 0.0 and ops.linear_ln_supported(x, proj.weight, blk.norm1):
            x = ops.linear_ln(x, proj.weight, pre_bias, blk.norm1, residual=shortcut)
        else:
            if proj is not None:
                x = ops.linear(x, proj.weight)
            x = ops.layer_norm(x, blk.norm1, residual=shortcut, pre_bias=pre_bias, row_scale=scale_of(x),
                               in_drop=pre_drop)
        if plain and blk.mlp.ln_fusable(x, blk.norm2):
            return blk.mlp.forward_ln(x, blk.norm2)
        h, hb, hdrop, x = blk.mlp.forward_split(x)
        return ops.layer_norm(h, blk.norm2, residual=x, pre_bias=hb, row_scale=scale_of(x), in_drop=hdrop)
    x = shortcut + dp(x)
    return x + dp(blk.mlp(ops.layer_norm(x, blk.norm2)))


SwinHPTransformerBlock = SwinTransformerBlock  # the name BASELINE.json uses


class PatchMerging(nn.Module):
    """4 nested siblings -> 1 token: view (B, N/4, 4C) -> LayerNorm -> Linear(4C -> 2C)   [:364-395]"""

    def __init__(self, dim, dim_scale=2, norm_layer=nn.LayerNorm):
        super().__init__()
        self.dim = dim
        self.reduction = nn.Linear(4 * dim, dim_scale * dim, bias=False)
        self.norm = norm_layer(4 * dim)

    def forward(self, x):
        B, N, C = x.shape
        assert N % 4 == 0, f"x size {N} is not divisible by 4 as necessary for patching."
        # cat(x[:,0::4], ..., x[:,3::4]) on the channel axis is a plain view in nested order
        # (fused gather + norm + linear: the LayerNorm is folded into the reduction GEMM, ops.ln_linear)
        x = x.contiguous().view(B, N // 4, 4 * C)
        return ops.ln_linear(x, self.norm, self.reduction.weight, self.reduction.bias)

    def extra_repr(self) -> str:
        return f"dim={self.dim}"


PatchMerge = PatchMerging


class PatchExpand(nn.Module):
    """1 token -> 4 nested children: Linear(C -> 2C) -> view (B, 4N, C/2) -> LayerNorm   [:407-430]"""

    def __init__(self, dim, dim_scale=2, norm_layer=nn.LayerNorm):
        super().__init__()
        self.dim = dim
        self.expand = nn.Linear(dim, dim_scale * dim, bias=False) if dim_scale != 1 else nn.Identity()
        self.norm = norm_layer(dim * dim_scale // 4)

    def forward(self, x):
        if isinstance(self.expand, nn.Linear):
            w = self.expand.weight
            if w.shape[0] % 4 == 0 and ops.linear_ln_supported(x, w, self.norm) \
                    and self.norm.normalized_shape[0] == w.shape[0] // 4:
                # fused gather + linear + norm: the four children of a token are the four C/2-wide column groups of its
                # expanded row, so the LayerNorm of the (B, 4N, C/2) view runs in the GEMM's epilogue
                B, N, _ = x.shape
                return ops.linear_ln(x, w, self.expand.bias, self.norm).view(B, 4 * N, w.shape[0] // 4)
            x = ops.linear(x, w, self.expand.bias)
        B, N, C = x.shape
        return ops.layer_norm(x.contiguous().view(B, 4 * N, C // 4), self.norm)


PatchExpanding = PatchExpand  # the name BASELINE.json uses


class FinalPatchExpand_X4(nn.Module):
    """[:433-452]"""

    def __init__(self, patch_size, dim, norm_layer=nn.LayerNorm):
        super().__init__()
        self.dim = dim
        self.patch_size = patch_size
        self.expand = nn.Linear(dim, patch_size * dim, bias=False)
        self.output_dim = dim
        self.norm = norm_layer(self.output_dim)

    def forward_pre_norm(self, x):
        """The expanded tokens before ``self.norm`` (the decoder fuses that LayerNorm with the output projection)."""
        x = ops.linear(x, self.expand.weight, self.expand.bias)
        B, N, C = x.shape
        return x.contiguous().view(B, N * self.patch_size, C // self.patch_size)

    def forward(self, x):
        return ops.layer_norm(self.forward_pre_norm(x), self.norm)


def _make_blocks(dim, input_resolution, depth, num_heads, window_size, base_pix, shift_size, shift_strategy,
                 rel_pos_bias, mlp_ratio, qkv_bias, qk_scale, drop, attn_drop, drop_path, norm_layer,
                 use_v2_norm_placement, use_cos_attn):
    # even blocks W-MSA, odd blocks SW-MSA   [:508-531, :614-637]
    return nn.ModuleList([
        SwinTransformerBlock(
            dim=dim, input_resolution=input_resolution, num_heads=num_heads, window_size=window_size,
            base_pix=base_pix, shift_size=0 if (i % 2 == 0) else shift_size, shift_strategy=shift_strategy,
            rel_pos_bias=rel_pos_bias, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop,
            attn_drop=attn_drop, drop_path=drop_path[i] if isinstance(drop_path, list) else drop_path,
            norm_layer=norm_layer, use_v2_norm_placement=use_v2_norm_placement, use_cos_attn=use_cos_attn)
        for i in range(depth)])


class BasicLayer(nn.Module):
    """One encoder stage: ``depth`` blocks (+ PatchMerging)   [:455-547]"""

    def __init__(self, dim, input_resolution, depth, num_heads, window_size, base_pix, shift_size, shift_strategy,
                 rel_pos_bias, mlp_ratio=4.0, qkv_bias=True, qk_scale=None, drop=0.0, attn_drop=0.0, drop_path=0.0,
                 norm_layer=nn.LayerNorm, downsample=None, use_checkpoint=False, use_v2_norm_placement=False,
                 use_cos_attn=False):
        super().__init__()
        self.dim = dim
        self.input_resolution = input_resolution
        self.depth = depth
        self.use_checkpoint = use_checkpoint
        self.blocks = _make_blocks(dim, input_resolution, depth, num_heads, window_size, base_pix, shift_size,
                                   shift_strategy, rel_pos_bias, mlp_ratio, qkv_bias, qk_scale, drop, attn_drop,
                                   drop_path, norm_layer, use_v2_norm_placement, use_cos_attn)
        self.downsample = downsample(dim=dim, norm_layer=norm_layer) if downsample is not None else None

    def forward(self, x):
        for blk in self.blocks:
            x = checkpoint.checkpoint(blk, x) if self.use_checkpoint else blk(x)
        if self.downsample is not None:
            x = self.downsample(x)
        return x

    def extra_repr(self) -> str:
        return f"dim={self.dim}, input_resolution={self.input_resolution}, depth={self.depth}"


class BasicLayer_up(nn.Module):
    """One decoder stage: ``depth`` blocks (+ PatchExpand)   [:561-653]"""

    def __init__(self, dim, input_resolution, depth, num_heads, window_size, base_pix, shift_size, shift_strategy,
                 rel_pos_bias, mlp_ratio=4.0, qkv_bias=True, qk_scale=None, drop=0.0, attn_drop=0.0, drop_path=0.0,
                 norm_layer=nn.LayerNorm, upsample=None, use_checkpoint=False, use_v2_norm_placement=False,
                 use_cos_attn=False):
        super().__init__()
        self.dim = dim
        self.input_resolution = input_resolution
        self.depth = depth
        self.use_checkpoint = use_checkpoint
        self.blocks = _make_blocks(dim, input_resolution, depth, num_heads, window_size, base_pix, shift_size,
                                   shift_strategy, rel_pos_bias, mlp_ratio, qkv_bias, qk_scale, drop, attn_drop,
                                   drop_path, norm_layer, use_v2_norm_placement, use_cos_attn)
        self.upsample = PatchExpand(dim=dim, dim_scale=2, norm_layer=norm_layer) if upsample is not None else None

    def forward(self, x):
        for blk in self.blocks:
            x = checkpoint.checkpoint(blk, x) if self.use_checkpoint else blk(x)
        if self.upsample is not None:
            x = self.upsample(x)
        return x


class PatchEmbed(nn.Module):
    """Conv1d(k = s = patch_size) over the nested pixel axis   [:656-694]"""

    def __init__(self, config, data_spec):
        super().__init__()
        assert config.patch_size % 4 == 0, "required for valid nside in deeper layers"
        self.config = config
        self.data_spec = data_spec
        self.num_patches = data_spec.dim_in // config.patch_size
        self.proj = nn.Conv1d(data_spec.f_in, config.embed_dim, kernel_size=config.patch_size,
                              stride=config.patch_size)
        self.norm = config.patch_embed_norm_layer if config.patch_embed_norm_layer is not None else None

    def forward(self, x):
        B, C, N = x.shape
        assert N == self.data_spec.dim_in, f"Input image size ({N}) doesn't match model ({self.data_spec.dim_in})."
        # Conv1d(k = s = patch) == Linear over the (f_in x patch) values of every patch; written that way the result is
        # already the contiguous (B, N/patch, C) token tensor (the conv gives (B, C, N/patch) and a strided view)
        p = self.config.patch_size
        patches = x.reshape(B, C, N // p, p).permute(0, 2, 1, 3).reshape(B, N // p, C * p)
        x = ops.linear(patches, self.proj.weight.reshape(self.proj.weight.shape[0], C * p), self.proj.bias)
        if self.norm is not None:
            x = self.norm(x)
        return x


class UnetDecoder(nn.Module):
    """[:704-791]"""

    def __init__(self, config, data_spec, dpr):
        super().__init__()
        self.config = config
        self.num_layers = len(config.depths)
        self.num_features = int(config.embed_dim * 2 ** (self.num_layers - 1))
        num_patches = data_spec.dim_in // config.patch_size
        self.layers_up = nn.ModuleList()
        self.concat_back_dim = nn.ModuleList()
        for i_layer in range(self.num_layers):
            down_idx = self.num_layers - 1 - i_layer
            width = int(config.embed_dim * 2**down_idx)
            if i_layer == 0:
                self.concat_back_dim.append(nn.Identity())
                self.layers_up.append(PatchExpand(dim=width, dim_scale=2, norm_layer=config.norm_layer))
                continue
            self.concat_back_dim.append(nn.Linear(2 * width, width))
            lo, hi = sum(config.depths[:down_idx]), sum(config.depths[: down_idx + 1])
            self.layers_up.append(BasicLayer_up(
                dim=width, input_resolution=num_patches // (4**down_idx), depth=config.depths[down_idx],
                num_heads=config.num_heads[down_idx], window_size=config.window_size, base_pix=data_spec.base_pix,
                shift_size=config.shift_size, shift_strategy=config.shift_strategy, rel_pos_bias=config.rel_pos_bias,
                mlp_ratio=config.mlp_ratio, qkv_bias=config.qkv_bias, qk_scale=config.qk_scale,
                use_cos_attn=config.use_cos_attn, drop=config.drop_rate, attn_drop=config.attn_drop_rate,
                drop_path=dpr[lo:hi], norm_layer=config.norm_layer,
                use_v2_norm_placement=config.use_v2_norm_placement,
                upsample=PatchExpand if down_idx > 0 else None, use_checkpoint=config.use_checkpoint))
        self.up = FinalPatchExpand_X4(patch_size=config.patch_size, dim=config.embed_dim)
        self.output = nn.Conv1d(in_channels=config.embed_dim, out_channels=data_spec.f_out, kernel_size=1, bias=False)
        self.norm_up = config.norm_layer(config.embed_dim)

    def forward(self, x, x_downsample):
        for inx, layer_up in enumerate(self.layers_up):
            if inx > 0:
                # cat + Linear(2C -> C) (:772-775) as two accumulating GEMMs: the concatenated tensor is never written
                x = ops.cat_linear(x, x_downsample[self.num_layers - 1 - inx], self.concat_back_dim[inx].weight,
                                   self.concat_back_dim[inx].bias)
            x = layer_up(x)
        x = self.up.forward_pre_norm(ops.layer_norm(x, self.norm_up))
        # up.norm + Conv1d(kernel 1) over channels in one pass: the normalised (B, N_pix, C) activation is never written,
        # the f_out-channel result comes out directly as (B, f_out, N_pix)
        return ops.ln_head(x, self.up.norm, self.output.weight[:, :, 0], self.output.bias)


@dataclass
class SwinHPTransformerConfig:
    """[:794-818] -- same fields, same defaults"""

    patch_size: int = 4
    window_size: int = 4
    shift_size: int = 2
    shift_strategy: Literal["nest_roll", "nest_grid_shift", "ring_shift"] = "nest_roll"
    rel_pos_bias: Optional[Literal["flat"]] = None
    embed_dim: int = 96
    patch_embed_norm_layer: Optional[Literal[nn.LayerNorm]] = None
    depths: List[int] = field(default_factory=lambda: [2, 2, 2, 2])
    num_heads: List[int] = field(default_factory=lambda: [3, 6, 12, 24])
    mlp_ratio: float = 4.0
    qkv_bias: bool = True
    qk_scale: Optional[float] = None
    use_cos_attn: bool = False
    drop_rate: float = 0.0
    attn_drop_rate: float = 0.0
    drop_path_rate: float = 0.1
    norm_layer: Literal[nn.LayerNorm] = nn.LayerNorm
    use_v2_norm_placement: bool = False
    ape: bool = False
    patch_norm: bool = True
    use_checkpoint: bool = False
    dev_mode: bool = False
    decoder_class: Literal[UnetDecoder] = UnetDecoder


class SwinHPTransformerSys(nn.Module):
    """HEAL-SWIN-UNet   [:821-955]: forward(x: (B, f_in, N_pix)) -> (B, f_out, N_pix)"""

    def __init__(self, config: SwinHPTransformerConfig, data_spec: DataSpec, **kwargs):
        super().__init__()
        self.config = config
        self.data_spec = data_spec
        self.num_layers = len(config.depths)
        self.num_features = int(config.embed_dim * 2 ** (self.num_layers - 1))
        self.num_features_up = int(config.embed_dim * 2)

        self.patch_embed = PatchEmbed(config, data_spec=data_spec)
        num_patches = self.patch_embed.num_patches
        if config.ape:
            self.absolute_pos_embed = nn.Parameter(torch.zeros(1, num_patches, config.embed_dim))
            nn.init.trunc_normal_(self.absolute_pos_embed, std=0.02)
        self.pos_drop = nn.Dropout(p=config.drop_rate)

        dpr = [v.item() for v in torch.linspace(0, config.drop_path_rate, sum(config.depths))]  # :871-873
        self.layers = nn.ModuleList()
        for i_layer in range(self.num_layers):
            lo, hi = sum(config.depths[:i_layer]), sum(config.depths[: i_layer + 1])
            self.layers.append(BasicLayer(
                dim=int(config.embed_dim * 2**i_layer), input_resolution=num_patches // (4**i_layer),
                depth=config.depths[i_layer], num_heads=config.num_heads[i_layer], window_size=config.window_size,
                base_pix=data_spec.base_pix, shift_size=config.shift_size, shift_strategy=config.shift_strategy,
                rel_pos_bias=config.rel_pos_bias, mlp_ratio=config.mlp_ratio, qkv_bias=config.qkv_bias,
                qk_scale=config.qk_scale, use_cos_attn=config.use_cos_attn, drop=config.drop_rate,
                attn_drop=config.attn_drop_rate, drop_path=dpr[lo:hi], norm_layer=config.norm_layer,
                use_v2_norm_placement=config.use_v2_norm_placement,
                downsample=PatchMerging if (i_layer < self.num_layers - 1) else None,
                use_checkpoint=config.use_checkpoint))
        self.decoder = config.decoder_class(config, data_spec, dpr)
        out_channels = self.num_features * (1 if config.decoder_class == UnetDecoder else 2)
        self.norm = config.norm_layer(out_channels)
        self.apply(self._init_weights)

    def _init_weights(self, m):  # :912-919
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {"absolute_pos_embed"}

    @torch.jit.ignore
    def no_weight_decay_keywords(self):
        return {"relative_position_bias_table"}

    def forward_features(self, x):  # :930-946
        x = self.patch_embed(x)
        if self.config.ape:
            x = x + self.absolute_pos_embed
        x = self.pos_drop(x)
        x_downsample = []
        for layer in self.layers:
            x_downsample.append(x)
            x = layer(x)
        return ops.layer_norm(x, self.norm), x_downsample

    def forward(self, x):  # :948-955
        x, x_downsample = self.forward_features(x)
        return self.decoder(x, x_downsample)
