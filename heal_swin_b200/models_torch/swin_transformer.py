"""Drop-in mirror of heal_swin/models_torch/swin_transformer.py (the flat SWIN-UNet baseline) on the B200 engine.

Same class names, constructor signatures, parameter / buffer names as the reference so checkpoints and
the Lightning wrappers keep working.  The flat model shares the windowed-attention kernels with the
HEALPix model: a 2-D window of ``wh x ww`` pixels is just another gather table, so

    roll(-s) -> window_partition (6-D permute copy) -> attention -> window_reverse (copy) -> roll(+s)

(swin_transformer.py:366-389) becomes one kernel launch that reads q, k, v rows through the table
``src[slot] = ((h + s) % H) * W + (w + s) % W`` and writes each output row back to the token it was read
from; the nine-region SW-MSA mask (:312-352) is one byte per pixel instead of an (nW, ws, ws) fp32 tensor.
The 2x2 strided patch merge (:454-459) and the ``(p1 p2 c)`` pixel-shuffle of the expand layers
(:494-499, :523-531) are single row-gather launches.
"""
from dataclasses import dataclass, field
from typing import List, Literal, Optional, Tuple, Union

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.utils.checkpoint as checkpoint

from .. import hp_index, ops
from ..data_spec import DataSpec
from .swin_hp_transformer import DropPath, Mlp, _residual_tail  # identical in both reference files


def _pair(v):
    if isinstance(v, int):
        return [v, v]
    v = list(v)
    return [v[0], v[0]] if len(v) == 1 else v


# ----------------------------------------------------------------------------- index tables (host, init time)
def window_slot_table(H, W, wh, ww, roll_h=0, roll_w=0):
    """int64 (H*W,): token read by every window slot after ``torch.roll(x, (-roll_h, -roll_w), (1, 2))`` and the
    2-D window partition (swin_transformer.py:44-56, 366-368).  Slot order: window row, window column, then the
    pixel's row and column inside the window."""
    wi, wj, a, b = np.meshgrid(np.arange(H // wh), np.arange(W // ww), np.arange(wh), np.arange(ww), indexing="ij")
    hs, ws_ = wi * wh + a, wj * ww + b
    return torch.from_numpy((((hs + roll_h) % H) * W + (ws_ + roll_w) % W).reshape(-1).astype(np.int64))


def shift_region_ids(H, W, wh, ww, sh, sw):
    """int8 (H*W,) in window-slot order: the 0..8 region id of swin_transformer.py:312-336 (literal slice
    semantics, including the degenerate ``slice(-0, None)`` when one of the two shifts is zero)."""
    img = np.zeros((H, W), dtype=np.int8)
    cnt = 0
    for hsl in (slice(0, -wh), slice(-wh, -sh), slice(-sh, None)):
        for wsl in (slice(0, -ww), slice(-ww, -sw), slice(-sw, None)):
            img[hsl, wsl] = cnt
            cnt += 1
    slots = window_slot_table(H, W, wh, ww).numpy()
    return torch.from_numpy(img.reshape(-1)[slots].copy())


def flat_rel_pos_index(wh, ww):
    """(wh*ww, wh*ww) int64 relative position index (swin_transformer.py:125-135)."""
    r, c = np.divmod(np.arange(wh * ww), ww)
    idx = (r[:, None] - r[None, :] + wh - 1) * (2 * ww - 1) + (c[:, None] - c[None, :] + ww - 1)
    return torch.from_numpy(idx.astype(np.int64))


class _RowTable(nn.Module):
    """A fixed row permutation / gather applied with the hs_gather_rows kernel (int32 tables as buffers)."""

    def __init__(self, idx: torch.Tensor):
        super().__init__()
        self.register_buffer("_hs_idx", idx.to(torch.int32).contiguous(), persistent=False)
        inv = torch.empty_like(idx)
        inv[idx] = torch.arange(idx.numel(), dtype=idx.dtype)
        self.register_buffer("_hs_inv", inv.to(torch.int32).contiguous(), persistent=False)

    def forward(self, x):  # x: (B, rows, C) -> (B, rows, C)
        return ops.GatherRows.apply(x, self._hs_idx, self._hs_inv)


def window_partition(x, window_size):
    """(B, H, W, C) -> (num_windows*B, wh, ww, C)   [swin_transformer.py:44-56]"""
    B, H, W, C = x.shape
    wh, ww = window_size
    idx = window_slot_table(H, W, wh, ww)
    inv = torch.empty_like(idx)
    inv[idx] = torch.arange(idx.numel())
    out = ops.GatherRows.apply(x.reshape(B, H * W, C), idx.to(device=x.device, dtype=torch.int32),
                               inv.to(device=x.device, dtype=torch.int32))
    return out.view(-1, wh, ww, C)


def window_reverse(windows, window_size, H, W):
    """(num_windows*B, wh, ww, C) -> (B, H, W, C)   [swin_transformer.py:59-75]"""
    wh, ww = window_size
    B = int(windows.shape[0] / (H * W / wh / ww))
    idx = window_slot_table(H, W, wh, ww)
    inv = torch.empty_like(idx)
    inv[idx] = torch.arange(idx.numel())
    out = ops.GatherRows.apply(windows.reshape(B, H * W, -1), inv.to(device=windows.device, dtype=torch.int32),
                               idx.to(device=windows.device, dtype=torch.int32))
    return out.view(B, H, W, -1)


class WindowAttention(nn.Module):
    """Window multi-head self attention with relative position bias   [swin_transformer.py:78-217]"""

    def __init__(self, dim, window_size, num_heads, qkv_bias=True, qk_scale=None, attn_drop=0.0, proj_drop=0.0,
                 use_cos_attn=False, use_rel_pos_bias=True):
        super().__init__()
        self.use_rel_pos_bias = use_rel_pos_bias
        self.dim = dim
        self.window_size = window_size
        self.num_heads = num_heads
        self.use_cos_attn = use_cos_attn
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        if use_cos_attn:  # :113-116
            self.logit_scale = nn.Parameter(torch.log(10 * torch.ones((num_heads, 1, 1))), requires_grad=True)
        wh, ww = window_size[0], window_size[1]
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * wh - 1) * (2 * ww - 1), num_heads))
        index = flat_rel_pos_index(wh, ww)
        self.register_buffer("relative_position_index", index)
        self.register_buffer("_hs_rel_index", index.to(torch.int32).reshape(-1).contiguous(), persistent=False)
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)  # :143
        self.softmax = nn.Softmax(dim=-1)

    def _core(self, qkv, n_tokens, src, groups, dense_mask):
        # nn.Dropout(attn_drop) on the softmax output (:167-169) runs inside the kernel (counter-based mask)
        drop_p = self.attn_drop.p if self.training else 0.0
        table = self.relative_position_bias_table if self.use_rel_pos_bias else None
        rel_index = self._hs_rel_index if table is not None else None
        logit_scale = self.logit_scale if self.use_cos_attn else None
        return ops.window_attention_core(qkv, table, logit_scale, src, groups, dense_mask, rel_index,
                                         self.scale, self.num_heads, n_tokens, self.use_cos_attn, attn_drop=drop_p)

    def forward_tokens(self, x, n_tokens, src=None, groups=None):
        """x: (B, H*W, C) in raster order; ``src`` / ``groups``: the block's window-slot tables."""
        out = self._core(ops.linear(x, self.qkv.weight, self.qkv.bias), n_tokens, src, groups, None)
        return self.proj_drop(ops.linear(out, self.proj.weight, self.proj.bias))

    def forward_tokens_split(self, x, n_tokens, src=None, groups=None):
        """(proj output WITHOUT its bias and proj_drop, that bias or None, the dropout p still to apply, the input as
        residual shortcut): bias and dropout are fused into the following LayerNorm, the shortcut's gradient into the qkv
        input-gradient GEMM."""
        if self.proj.bias is None:
            return self.forward_tokens(x, n_tokens, src, groups), None, 0.0, x
        qkv, shortcut = ops.linear(x, self.qkv.weight, self.qkv.bias, fork=True)
        out = self._core(qkv, n_tokens, src, groups, None)
        return ops.linear(out, self.proj.weight), self.proj.bias, (self.proj_drop.p if self.training else 0.0), shortcut

    def forward(self, x, mask=None):
        """x: (num_windows*B, N, C); mask: (num_windows, N, N) additive or None   [:148-202]"""
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
    """[swin_transformer.py:220-423]"""

    def __init__(self, dim, input_resolution, num_heads, window_size=[4, 4], shift_size=-1, mlp_ratio=4.0,
                 qkv_bias=True, qk_scale=None, drop=0.0, attn_drop=0.0, drop_path=0.0, act_layer=nn.GELU,
                 norm_layer=nn.LayerNorm, use_masking=True, use_cos_attn=False, use_v2_norm_placement=False,
                 use_rel_pos_bias=True):
        super().__init__()
        self.dim = dim
        self.input_resolution = input_resolution
        self.num_heads = num_heads
        self.window_size = window_size
        self.shift_size = tuple([window_size[0] // 2, window_size[1] // 2]) if shift_size == -1 else shift_size
        self.mlp_ratio = mlp_ratio
        self.use_v2_norm_placement = use_v2_norm_placement
        if self.input_resolution[0] <= self.window_size[0] or self.input_resolution[1] <= self.window_size[1]:
            self.shift_size = [0, 0]  # :274-280
            self.window_size = self.input_resolution
        msg = "Shift size and window size must satisfy 0 <= shift_size[{i}] < window_size[{i}], got shift_size[{i}]={ss} and window_size[{i}]={ws}"
        for i in (0, 1):
            assert 0 <= self.shift_size[i] < self.window_size[i], msg.format(i=i, ss=self.shift_size[i], ws=self.window_size[i])

        self.norm1 = norm_layer(dim)
        self.attn = WindowAttention(dim, window_size=window_size, num_heads=num_heads, qkv_bias=qkv_bias,
                                    qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop,
                                    use_rel_pos_bias=use_rel_pos_bias, use_cos_attn=use_cos_attn)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

        H, W = self.input_resolution
        wh, ww = self.window_size
        sh, sw = self.shift_size
        shifted = sh > 0 or sw > 0
        # the forward roll uses shift_size[0] on both axes (:366-368) -- reproduced, not fixed
        self.register_buffer("_hs_src", window_slot_table(H, W, wh, ww, sh, sh if shifted else 0).to(torch.int32),
                             persistent=False)
        if use_masking and shifted:
            groups = shift_region_ids(H, W, wh, ww, sh, sw)
            attn_mask = hp_index.attn_mask_from_groups(groups, wh * ww)  # (nW, ws, ws) in {0, -100}   :338-352
            self.register_buffer("_hs_groups", groups.to(torch.uint8).contiguous(), persistent=False)
        else:
            attn_mask = None
            self._hs_groups = None
        self.register_buffer("attn_mask", attn_mask)
        # the reverse roll uses (shift_size[0], shift_size[1]) (:389): for sh != sw the output lands sw - sh columns
        # away from where it was read; that residual column roll is one extra row gather
        self.fixup = None
        if shifted and sh != sw:
            p = np.arange(H * W)
            self.fixup = _RowTable(torch.from_numpy(((p // W) * W + (p % W - (sw - sh)) % W).astype(np.int64)))

    def forward(self, x):
        H, W = self.input_resolution
        B, L, C = x.shape
        assert L == H * W, "input feature has wrong size"
        shortcut = x
        if not self.use_v2_norm_placement:
            x = ops.layer_norm(x, self.norm1)
        n_tok = self.window_size[0] * self.window_size[1]
        if self.use_v2_norm_placement:
            x, pre_bias, pdrop, shortcut = self.attn.forward_tokens_split(x, n_tok, self._hs_src, self._hs_groups)
        else:
            x, pre_bias, pdrop = self.attn.forward_tokens(x, n_tok, self._hs_src, self._hs_groups), None, 0.0
        if self.fixup is not None:
            x = self.fixup(x)
        return _residual_tail(self, shortcut, x, pre_bias, pdrop)

    def extra_repr(self) -> str:
        return (f"dim={self.dim}, input_resolution={self.input_resolution}, num_heads={self.num_heads},"
                f" window_size={self.window_size}, shift_size={self.shift_size}, mlp_ratio={self.mlp_ratio}")


class PatchMerging(nn.Module):
    """2x2 neighbourhood -> 1 token: gather (x[0::2,0::2], x[1::2,0::2], x[0::2,1::2], x[1::2,1::2]) on the channel
    axis, LayerNorm(4C), Linear(4C -> 2C)   [swin_transformer.py:426-472]"""

    def __init__(self, input_resolution, dim, norm_layer=nn.LayerNorm):
        super().__init__()
        self.input_resolution = input_resolution
        self.dim = dim
        self.patch_size = 4
        self.reduction = nn.Linear(self.patch_size * dim, 2 * dim, bias=False)
        self.norm = norm_layer(self.patch_size * dim)
        H, W = input_resolution
        if H % 2 == 0 and W % 2 == 0:
            i, j, q = np.meshgrid(np.arange(H // 2), np.arange(W // 2), np.arange(4), indexing="ij")
            rows = (2 * i + q % 2) * W + 2 * j + q // 2
            self.gather = _RowTable(torch.from_numpy(rows.reshape(-1).astype(np.int64)))

    def forward(self, x):
        H, W = self.input_resolution
        B, L, C = x.shape
        assert L == H * W, "input feature has wrong size"
        assert H % 2 == 0 and W % 2 == 0, f"x size ({H}*{W}) are not even."
        x = self.gather(x).view(B, L // 4, 4 * C)
        return ops.linear(ops.layer_norm(x, self.norm), self.reduction.weight, self.reduction.bias)

    def extra_repr(self) -> str:
        return f"input_resolution={self.input_resolution}, dim={self.dim}"


def _pixel_shuffle_table(H, W, p1, p2):
    """row table of einops ``b h w (p1 p2 c) -> b (h p1) (w p2) c`` viewed as rows of c channels"""
    hh, ww = np.meshgrid(np.arange(H * p1), np.arange(W * p2), indexing="ij")
    src = ((hh // p1) * W + ww // p2) * (p1 * p2) + (hh % p1) * p2 + ww % p2
    return torch.from_numpy(src.reshape(-1).astype(np.int64))


class PatchExpand(nn.Module):
    """[swin_transformer.py:476-501]"""

    def __init__(self, input_resolution, dim, dim_scale=2, norm_layer=nn.LayerNorm):
        super().__init__()
        self.input_resolution = input_resolution
        self.dim = dim
        self.expand = nn.Linear(dim, 2 * dim, bias=False) if dim_scale == 2 else nn.Identity()
        self.norm = norm_layer(dim // dim_scale)
        self.dim_scale = 4
        self.shuffle = _RowTable(_pixel_shuffle_table(input_resolution[0], input_resolution[1], 2, 2))

    def forward(self, x):
        H, W = self.input_resolution
        if isinstance(self.expand, nn.Linear):
            x = ops.linear(x, self.expand.weight, self.expand.bias)
        B, L, C = x.shape
        assert L == H * W, "input feature has wrong size"
        x = self.shuffle(x.contiguous().view(B, L * 4, C // self.dim_scale))
        return ops.layer_norm(x, self.norm)


class FinalPatchExpand_X4(nn.Module):
    """[swin_transformer.py:504-535]"""

    def __init__(self, input_resolution, patch_size, dim, norm_layer=nn.LayerNorm):
        super().__init__()
        self.input_resolution = input_resolution
        self.dim = dim
        self.patch_size = patch_size
        self.L = input_resolution[0] * input_resolution[1]
        self.expand = nn.Linear(dim, (patch_size[0] * patch_size[1]) * dim, bias=False)
        self.output_dim = dim
        self.norm = norm_layer(self.output_dim)
        self.shuffle = _RowTable(_pixel_shuffle_table(input_resolution[0], input_resolution[1], patch_size[0], patch_size[1]))

    def forward_pre_norm(self, x):
        """The pixel-shuffled tokens before ``self.norm`` (the decoder fuses that LayerNorm with the output projection)."""
        H, W = self.input_resolution
        x = ops.linear(x, self.expand.weight, self.expand.bias)
        B, L, C = x.shape
        assert L == H * W, "input feature has wrong size"
        pp = self.patch_size[0] * self.patch_size[1]
        return self.shuffle(x.contiguous().view(B, L * pp, C // pp))

    def forward(self, x):
        return ops.layer_norm(self.forward_pre_norm(x), self.norm)


def _make_blocks(dim, input_resolution, depth, num_heads, window_size, shift_size, mlp_ratio, qkv_bias, qk_scale, drop,
                 attn_drop, drop_path, norm_layer, use_masking, use_cos_attn, use_v2_norm_placement, use_rel_pos_bias):
    return nn.ModuleList([
        SwinTransformerBlock(
            dim=dim, input_resolution=input_resolution, num_heads=num_heads, window_size=window_size,
            shift_size=[0, 0] if (i % 2 == 0) else shift_size, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
            qk_scale=qk_scale, drop=drop, attn_drop=attn_drop,
            drop_path=drop_path[i] if isinstance(drop_path, list) else drop_path, norm_layer=norm_layer,
            use_masking=use_masking, use_rel_pos_bias=use_rel_pos_bias, use_v2_norm_placement=use_v2_norm_placement,
            use_cos_attn=use_cos_attn)
        for i in range(depth)])


class BasicLayer(nn.Module):
    """[swin_transformer.py:538-640]"""

    def __init__(self, dim, input_resolution, depth, num_heads, window_size, shift_size, mlp_ratio=4.0, qkv_bias=True,
                 qk_scale=None, drop=0.0, attn_drop=0.0, drop_path=0.0, norm_layer=nn.LayerNorm, downsample=None,
                 use_checkpoint=False, use_masking=True, use_cos_attn=False, use_v2_norm_placement=False,
                 use_rel_pos_bias=True):
        super().__init__()
        self.dim = dim
        self.input_resolution = input_resolution
        self.depth = depth
        self.use_checkpoint = use_checkpoint
        self.blocks = _make_blocks(dim, input_resolution, depth, num_heads, window_size, shift_size, mlp_ratio, qkv_bias,
                                   qk_scale, drop, attn_drop, drop_path, norm_layer, use_masking, use_cos_attn,
                                   use_v2_norm_placement, use_rel_pos_bias)
        self.downsample = downsample(input_resolution, dim=dim, norm_layer=norm_layer) if downsample is not None else None

    def forward(self, x):
        for blk in self.blocks:
            x = checkpoint.checkpoint(blk, x) if self.use_checkpoint else blk(x)
        if self.downsample is not None:
            x = self.downsample(x)
        return x

    def extra_repr(self) -> str:
        return f"dim={self.dim}, input_resolution={self.input_resolution}, depth={self.depth}"


class BasicLayer_up(nn.Module):
    """[swin_transformer.py:643-741]"""

    def __init__(self, dim, input_resolution, depth, num_heads, window_size, shift_size, mlp_ratio=4.0, qkv_bias=True,
                 qk_scale=None, drop=0.0, attn_drop=0.0, drop_path=0.0, norm_layer=nn.LayerNorm, upsample=None,
                 use_checkpoint=False, use_masking=True, use_cos_attn=False, use_v2_norm_placement=False,
                 use_rel_pos_bias=True):
        super().__init__()
        self.dim = dim
        self.input_resolution = input_resolution
        self.depth = depth
        self.use_checkpoint = use_checkpoint
        self.blocks = _make_blocks(dim, input_resolution, depth, num_heads, window_size, shift_size, mlp_ratio, qkv_bias,
                                   qk_scale, drop, attn_drop, drop_path, norm_layer, use_masking, use_cos_attn,
                                   use_v2_norm_placement, use_rel_pos_bias)
        self.upsample = (PatchExpand(input_resolution, dim=dim, dim_scale=2, norm_layer=norm_layer)
                         if upsample is not None else None)

    def forward(self, x):
        for blk in self.blocks:
            x = checkpoint.checkpoint(blk, x) if self.use_checkpoint else blk(x)
        if self.upsample is not None:
            x = self.upsample(x)
        return x


class PatchEmbed(nn.Module):
    """Conv2d(k = s = patch_size)   [swin_transformer.py:744-790]"""

    def __init__(self, config, data_spec):
        super().__init__()
        self.config = config
        self.data_spec = data_spec
        self.patches_resolution = [data_spec.dim_in[0] // config.patch_size[0], data_spec.dim_in[1] // config.patch_size[1]]
        self.num_patches = self.patches_resolution[0] * self.patches_resolution[1]
        self.proj = nn.Conv2d(data_spec.f_in, config.embed_dim, kernel_size=config.patch_size, stride=config.patch_size)
        self.norm = config.patch_embed_norm_layer if config.patch_embed_norm_layer is not None else None

    def forward(self, x):
        B, C, H, W = x.shape
        assert H == self.data_spec.dim_in[0] and W == self.data_spec.dim_in[1], (
            f"Input image size {H}*{W} doesn't match model ({self.data_spec.dim_in[0]}*{self.data_spec.dim_in[1]}).")
        # Conv2d(k = s = patch) == Linear over the (f_in x p1 x p2) values of every patch -> contiguous (B, L, C) tokens
        p1, p2 = self.config.patch_size[0], self.config.patch_size[1]
        patches = x.reshape(B, C, H // p1, p1, W // p2, p2).permute(0, 2, 4, 1, 3, 5).reshape(B, -1, C * p1 * p2)
        x = ops.linear(patches, self.proj.weight.reshape(self.proj.weight.shape[0], -1), self.proj.bias)
        if self.norm is not None:
            x = self.norm(x)
        return x


@dataclass
class SwinTransformerConfig:
    """[swin_transformer.py:793-818] -- same fields, same defaults"""

    patch_size: Union[int, Tuple[int, int]] = (4, 4)
    window_size: Union[int, Tuple[int, int]] = (4, 4)
    shift_size: Union[int, Tuple[int, int]] = -1
    embed_dim: int = 96
    patch_embed_norm_layer: Optional[str] = None
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
    final_upsample: Literal["expand_first"] = "expand_first"
    use_masking: bool = True
    use_rel_pos_bias: bool = True
    dev_mode: bool = False


class SwinTransformerSys(nn.Module):
    """Flat SWIN-UNet   [swin_transformer.py:823-1130]: forward(x: (B, f_in, H, W)) -> (B, f_out, H, W)"""

    def __init__(self, config: SwinTransformerConfig, data_spec: DataSpec, **kwargs):
        super().__init__()
        self.config = config
        self.data_spec = data_spec
        self.num_layers = len(config.depths)
        self.num_features = int(config.embed_dim * 2 ** (self.num_layers - 1))
        self.num_features_up = int(config.embed_dim * 2)
        H, W = data_spec.dim_in[0], data_spec.dim_in[1]
        config.patch_size = _pair(config.patch_size)
        config.window_size = _pair(config.window_size)
        merge = 2 ** (self.num_layers - 1)
        for name, size, p, w in (("H", H, config.patch_size[0], config.window_size[0]),
                                 ("W", W, config.patch_size[1], config.window_size[1])):
            assert (size / (merge * p * w)) % 1 == 0, (
                f"{name} must be divisible by merge_factor*patch*window = {merge}*{p}*{w} = {merge * p * w}, got {size}")
        assert (H * W / (merge**2 * config.patch_size[0] * config.patch_size[1])) % 1 == 0
        if config.shift_size == -1:
            self.shift_size = tuple([config.window_size[0] // 2, config.window_size[1] // 2])
        else:
            if isinstance(config.shift_size, int):
                config.shift_size = [config.shift_size, config.shift_size]
            self.shift_size = config.shift_size

        self.patch_embed = PatchEmbed(config, data_spec=data_spec)
        num_patches = self.patch_embed.num_patches
        res = self.patch_embed.patches_resolution
        self.patches_resolution = res
        if config.ape:
            self.absolute_pos_embed = nn.Parameter(torch.zeros(1, num_patches, config.embed_dim))
            nn.init.trunc_normal_(self.absolute_pos_embed, std=0.02)
        self.pos_drop = nn.Dropout(p=config.drop_rate)
        dpr = [v.item() for v in torch.linspace(0, config.drop_path_rate, sum(config.depths))]

        common = dict(window_size=config.window_size, shift_size=self.shift_size, mlp_ratio=config.mlp_ratio,
                      qkv_bias=config.qkv_bias, qk_scale=config.qk_scale, use_cos_attn=config.use_cos_attn,
                      drop=config.drop_rate, attn_drop=config.attn_drop_rate, norm_layer=config.norm_layer,
                      use_v2_norm_placement=config.use_v2_norm_placement, use_checkpoint=config.use_checkpoint,
                      use_masking=config.use_masking, use_rel_pos_bias=config.use_rel_pos_bias)
        self.layers = nn.ModuleList()
        for i in range(self.num_layers):
            self.layers.append(BasicLayer(
                dim=int(config.embed_dim * 2**i), input_resolution=(res[0] // 2**i, res[1] // 2**i),
                depth=config.depths[i], num_heads=config.num_heads[i],
                drop_path=dpr[sum(config.depths[:i]): sum(config.depths[: i + 1])],
                downsample=PatchMerging if (i < self.num_layers - 1) else None, **common))
        self.layers_up = nn.ModuleList()
        self.concat_back_dim = nn.ModuleList()
        for i in range(self.num_layers):
            d = self.num_layers - 1 - i
            width = int(config.embed_dim * 2**d)
            rs = (res[0] // 2**d, res[1] // 2**d)
            self.concat_back_dim.append(nn.Linear(2 * width, width) if i > 0 else nn.Identity())
            if i == 0:
                self.layers_up.append(PatchExpand(input_resolution=rs, dim=width, dim_scale=2, norm_layer=config.norm_layer))
            else:
                self.layers_up.append(BasicLayer_up(
                    dim=width, input_resolution=rs, depth=config.depths[d], num_heads=config.num_heads[d],
                    drop_path=dpr[sum(config.depths[:d]): sum(config.depths[: d + 1])],
                    upsample=PatchExpand if (i < self.num_layers - 1) else None, **common))
        self.norm = config.norm_layer(self.num_features)
        self.norm_up = config.norm_layer(config.embed_dim)
        if config.final_upsample == "expand_first":
            self.up = FinalPatchExpand_X4(input_resolution=(H // config.patch_size[0], W // config.patch_size[1]),
                                          patch_size=config.patch_size, dim=config.embed_dim)
            self.output = nn.Conv2d(in_channels=config.embed_dim, out_channels=data_spec.f_out, kernel_size=1, bias=False)
        self.apply(self._init_weights)

    def _init_weights(self, m):  # :1024-1031
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

    def forward_features(self, x):
        x = self.patch_embed(x)
        if self.config.ape:
            x = x + self.absolute_pos_embed
        x = self.pos_drop(x)
        x_downsample = []
        for layer in self.layers:
            x_downsample.append(x)
            x = layer(x)
        return ops.layer_norm(x, self.norm), x_downsample

    def forward_up_features(self, x, x_downsample):
        for inx, layer_up in enumerate(self.layers_up):
            if inx > 0:
                x = torch.cat([x, x_downsample[self.num_layers - 1 - inx]], -1)
                x = ops.linear(x, self.concat_back_dim[inx].weight, self.concat_back_dim[inx].bias)
            x = layer_up(x)
        return ops.layer_norm(x, self.norm_up)

    def up_x4(self, x):
        H, W = self.patches_resolution
        B, L, C = x.shape
        assert L == H * W, "input features has wrong size"
        if self.config.final_upsample == "expand_first":
            # up.norm + 1x1 Conv2d over channels in one pass; the result comes out channel-first, rows in raster order
            y = ops.ln_head(self.up.forward_pre_norm(x), self.up.norm, self.output.weight[:, :, 0, 0], self.output.bias)
            x = y.view(B, -1, self.config.patch_size[0] * H, self.config.patch_size[1] * W)
        return x

    def forward(self, x):
        x, x_downsample = self.forward_features(x)
        x = self.forward_up_features(x, x_downsample)
        return self.up_x4(x)
