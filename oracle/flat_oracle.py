"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the flat SWIN-UNet twin (heal_swin/models_torch/swin_transformer.py).

Functional fp32 restatement over a reference state_dict, every function citing the reference lines it follows.
Pinned against the reference itself through tests/golden/flat_*.npz (generated in the build container by
oracle/make_golden.py, which imports /root/reference read-only).  Only tests/, __graft_entry__.smoke() and
bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from .hp_oracle import LOGIT_SCALE_MAX, _ln


@dataclass
class FlatConfig:
    """SwinTransformerConfig (swin_transformer.py:793-818) + the DataSpec fields the model reads."""

    patch_size: Sequence[int] = (4, 4)
    window_size: Sequence[int] = (4, 4)
    shift_size: Sequence[int] = (2, 2)
    embed_dim: int = 96
    depths: Sequence[int] = field(default_factory=lambda: [2, 2, 2, 2])
    num_heads: Sequence[int] = field(default_factory=lambda: [3, 6, 12, 24])
    mlp_ratio: float = 4.0
    qkv_bias: bool = True
    qk_scale: Optional[float] = None
    use_cos_attn: bool = False
    use_v2_norm_placement: bool = False
    use_masking: bool = True
    use_rel_pos_bias: bool = True
    dim_in: Sequence[int] = (0, 0)
    f_in: int = 3
    f_out: int = 10


def window_partition(x, ws):
    """swin_transformer.py:44-56"""
    B, H, W, C = x.shape
    x = x.view(B, H // ws[0], ws[0], W // ws[1], ws[1], C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws[0], ws[1], C)


def window_reverse(windows, ws, H, W):
    """swin_transformer.py:59-75"""
    B = int(windows.shape[0] / (H * W / ws[0] / ws[1]))
    x = windows.view(B, H // ws[0], W // ws[1], ws[0], ws[1], -1)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(B, H, W, -1)


def relative_position_index(ws):
    """swin_transformer.py:125-135"""
    coords = torch.stack(torch.meshgrid([torch.arange(ws[0]), torch.arange(ws[1])], indexing="ij")).flatten(1)
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws[0] - 1
    rel[:, :, 1] += ws[1] - 1
    rel[:, :, 0] *= 2 * ws[1] - 1
    return rel.sum(-1)


def shifted_window_mask(H, W, ws, shift):
    """swin_transformer.py:312-352 -> (nW, wh*ww, wh*ww) in {0, -100}"""
    img = torch.zeros((1, H, W, 1))
    cnt = 0
    for h in (slice(0, -ws[0]), slice(-ws[0], -shift[0]), slice(-shift[0], None)):
        for w in (slice(0, -ws[1]), slice(-ws[1], -shift[1]), slice(-shift[1], None)):
            img[:, h, w, :] = cnt
            cnt += 1
    mw = window_partition(img, ws).view(-1, ws[0] * ws[1])
    m = mw.unsqueeze(1) - mw.unsqueeze(2)
    return m.masked_fill(m != 0, -100.0).masked_fill(m == 0, 0.0)


def window_attention(xw, sd, pre, num_heads, cfg: FlatConfig, ws, mask):
    """swin_transformer.py:148-202 (dropout 0)"""
    B_, n, C = xw.shape
    d = C // num_heads
    qkv = F.linear(xw, sd[pre + "qkv.weight"], sd.get(pre + "qkv.bias")).reshape(B_, n, 3, num_heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    if cfg.use_cos_attn:
        attn = F.normalize(q, dim=-1) @ F.normalize(k, dim=-1).transpose(-2, -1)
        attn = attn * torch.clamp(sd[pre + "logit_scale"], max=LOGIT_SCALE_MAX).exp()
    else:
        attn = (q * (cfg.qk_scale or d**-0.5)) @ k.transpose(-2, -1)
    if cfg.use_rel_pos_bias:
        bias = sd[pre + "relative_position_bias_table"][relative_position_index(ws).view(-1)].view(n, n, -1)
        attn = attn + bias.permute(2, 0, 1).unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = (attn.view(B_ // nW, nW, num_heads, n, n) + mask.unsqueeze(1).unsqueeze(0)).view(-1, num_heads, n, n)
    attn = attn.softmax(dim=-1)
    out = (attn @ v).transpose(1, 2).reshape(B_, n, C)
    return F.linear(out, sd[pre + "proj.weight"], sd[pre + "proj.bias"])


def swin_block(x, sd, pre, res, num_heads, cfg: FlatConfig, shift):
    """swin_transformer.py:236-403"""
    H, W = res
    B, L, C = x.shape
    ws = list(cfg.window_size)
    shift = list(shift)
    if H <= ws[0] or W <= ws[1]:
        shift, ws = [0, 0], [H, W]
    shifted = shift[0] > 0 or shift[1] > 0
    mask = shifted_window_mask(H, W, ws, shift) if (cfg.use_masking and shifted) else None
    shortcut = x
    if not cfg.use_v2_norm_placement:
        x = _ln(x, sd, pre + "norm1.")
    x = x.view(B, H, W, C)
    if shifted:  # NB both axes rolled by shift[0] on the way in (:366-368) ...
        x = torch.roll(x, shifts=(-shift[0], -shift[0]), dims=(1, 2))
    xw = window_partition(x, ws).view(-1, ws[0] * ws[1], C)
    aw = window_attention(xw, sd, pre + "attn.", num_heads, cfg, list(cfg.window_size), mask)
    x = window_reverse(aw.view(-1, ws[0], ws[1], C), ws, H, W)
    if shifted:  # ... and by (shift[0], shift[1]) on the way out (:389)
        x = torch.roll(x, shifts=(shift[0], shift[1]), dims=(1, 2))
    x = x.view(B, H * W, C)

    def mlp(t):
        t = F.gelu(F.linear(t, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"]))
        return F.linear(t, sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"])

    if cfg.use_v2_norm_placement:
        x = shortcut + _ln(x, sd, pre + "norm1.")
        x = x + _ln(mlp(x), sd, pre + "norm2.")
    else:
        x = shortcut + x
        x = x + mlp(_ln(x, sd, pre + "norm2."))
    return x


def patch_merging(x, sd, pre, res):
    """swin_transformer.py:443-466"""
    H, W = res
    B, L, C = x.shape
    x = x.view(B, H, W, C)
    x = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1).view(B, -1, 4 * C)
    return F.linear(_ln(x, sd, pre + "norm."), sd[pre + "reduction.weight"])


def patch_expand(x, sd, pre, res, p1=2, p2=2):
    """swin_transformer.py:485-501 and :515-535: Linear, 'b h w (p1 p2 c) -> b (h p1) (w p2) c', LayerNorm"""
    H, W = res
    x = F.linear(x, sd[pre + "expand.weight"])
    B, L, C = x.shape
    c = C // (p1 * p2)
    x = x.view(B, H, W, p1, p2, c).permute(0, 1, 3, 2, 4, 5).reshape(B, H * p1 * W * p2, c)
    return _ln(x, sd, pre + "norm.")


def flat_unet_forward(x, sd: Dict[str, torch.Tensor], cfg: FlatConfig):
    """SwinTransformerSys.forward, swin_transformer.py:1041-1117"""
    B = x.shape[0]
    ps = list(cfg.patch_size)
    L = len(cfg.depths)
    x = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=ps)
    R = (x.shape[2], x.shape[3])
    x = x.flatten(2).transpose(1, 2)
    res = lambda i: (R[0] // 2**i, R[1] // 2**i)  # noqa: E731

    def layer(x, pre, i):
        for b in range(cfg.depths[i]):
            x = swin_block(x, sd, f"{pre}blocks.{b}.", res(i), cfg.num_heads[i], cfg, [0, 0] if b % 2 == 0 else cfg.shift_size)
        return x

    skips = []
    for i in range(L):
        skips.append(x)
        x = layer(x, f"layers.{i}.", i)
        if i < L - 1:
            x = patch_merging(x, sd, f"layers.{i}.downsample.", res(i))
    x = _ln(x, sd, "norm.")
    for j in range(L):
        d = L - 1 - j
        if j == 0:
            x = patch_expand(x, sd, "layers_up.0.", res(d))
            continue
        x = torch.cat([x, skips[d]], -1)
        x = F.linear(x, sd[f"concat_back_dim.{j}.weight"], sd[f"concat_back_dim.{j}.bias"])
        x = layer(x, f"layers_up.{j}.", d)
        if d > 0:
            x = patch_expand(x, sd, f"layers_up.{j}.upsample.", res(d))
    x = _ln(x, sd, "norm_up.")
    x = patch_expand(x, sd, "up.", R, ps[0], ps[1])
    x = x.view(B, ps[0] * R[0], ps[1] * R[1], -1).permute(0, 3, 1, 2)
    return F.conv2d(x, sd["output.weight"])


def synth_state_dict(cfg: FlatConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Random weights in the reference state-dict layout of SwinTransformerSys (all affine terms randomised)."""
    g = torch.Generator().manual_seed(seed)

    def w(*shape, std=0.02):
        return torch.randn(*shape, generator=g) * std

    sd: Dict[str, torch.Tensor] = {}
    C0, L = cfg.embed_dim, len(cfg.depths)
    ws, ps = list(cfg.window_size), list(cfg.patch_size)

    def block(pre, C, h):
        sd[pre + "norm1.weight"] = 1 + w(C, std=0.1)
        sd[pre + "norm1.bias"] = w(C, std=0.1)
        sd[pre + "attn.qkv.weight"] = w(3 * C, C, std=0.1)
        if cfg.qkv_bias:
            sd[pre + "attn.qkv.bias"] = w(3 * C, std=0.1)
        sd[pre + "attn.proj.weight"] = w(C, C, std=0.1)
        sd[pre + "attn.proj.bias"] = w(C, std=0.1)
        if cfg.use_cos_attn:
            sd[pre + "attn.logit_scale"] = math.log(10.0) + w(h, 1, 1, std=0.3)
        sd[pre + "attn.relative_position_bias_table"] = w((2 * ws[0] - 1) * (2 * ws[1] - 1), h, std=0.5)
        sd[pre + "norm2.weight"] = 1 + w(C, std=0.1)
        sd[pre + "norm2.bias"] = w(C, std=0.1)
        Hd = int(C * cfg.mlp_ratio)
        sd[pre + "mlp.fc1.weight"] = w(Hd, C, std=0.1)
        sd[pre + "mlp.fc1.bias"] = w(Hd, std=0.1)
        sd[pre + "mlp.fc2.weight"] = w(C, Hd, std=0.05)
        sd[pre + "mlp.fc2.bias"] = w(C, std=0.1)

    sd["patch_embed.proj.weight"] = w(C0, cfg.f_in, ps[0], ps[1], std=0.3)
    sd["patch_embed.proj.bias"] = w(C0, std=0.1)
    for i in range(L):
        C = C0 * 2**i
        for b in range(cfg.depths[i]):
            block(f"layers.{i}.blocks.{b}.", C, cfg.num_heads[i])
        if i < L - 1:
            sd[f"layers.{i}.downsample.reduction.weight"] = w(2 * C, 4 * C, std=0.05)
            sd[f"layers.{i}.downsample.norm.weight"] = 1 + w(4 * C, std=0.1)
            sd[f"layers.{i}.downsample.norm.bias"] = w(4 * C, std=0.1)
    Ct = C0 * 2 ** (L - 1)
    sd["norm.weight"] = 1 + w(Ct, std=0.1)
    sd["norm.bias"] = w(Ct, std=0.1)
    sd["layers_up.0.expand.weight"] = w(2 * Ct, Ct, std=0.05)
    sd["layers_up.0.norm.weight"] = 1 + w(Ct // 2, std=0.1)
    sd["layers_up.0.norm.bias"] = w(Ct // 2, std=0.1)
    for j in range(1, L):
        d = L - 1 - j
        C = C0 * 2**d
        sd[f"concat_back_dim.{j}.weight"] = w(C, 2 * C, std=0.05)
        sd[f"concat_back_dim.{j}.bias"] = w(C, std=0.1)
        for b in range(cfg.depths[d]):
            block(f"layers_up.{j}.blocks.{b}.", C, cfg.num_heads[d])
        if d > 0:
            sd[f"layers_up.{j}.upsample.expand.weight"] = w(2 * C, C, std=0.05)
            sd[f"layers_up.{j}.upsample.norm.weight"] = 1 + w(C // 2, std=0.1)
            sd[f"layers_up.{j}.upsample.norm.bias"] = w(C // 2, std=0.1)
    sd["norm_up.weight"] = 1 + w(C0, std=0.1)
    sd["norm_up.bias"] = w(C0, std=0.1)
    sd["up.expand.weight"] = w(ps[0] * ps[1] * C0, C0, std=0.05)
    sd["up.norm.weight"] = 1 + w(C0, std=0.1)
    sd["up.norm.bias"] = w(C0, std=0.1)
    sd["output.weight"] = w(cfg.f_out, C0, 1, 1, std=0.1)
    return sd
