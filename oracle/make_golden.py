"""TEST INFRASTRUCTURE ONLY -- regenerate tests/golden/*.npz from the REAL reference.

Run in the build container (needs /root/reference):

    python -m oracle.make_golden

Everything written here is produced by the reference's own classes
(``heal_swin.models_torch``; imported read-only via oracle/ref_import.py), never by
the oracle restatement and never by the CUDA path.  Weights come from
``oracle.hp_oracle.synth_state_dict`` (deterministic torch CPU generator); a
checksum of them is stored so that a drifting RNG is detected instead of being
misread as a parity failure.
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import hp_oracle as O  # noqa: E402
from oracle.ref_import import import_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# name -> (HPConfig kwargs, batch)
MODEL_CASES = {
    # BASELINE.json configs[0] shape family, shrunk to nside=16 so the fixture stays small
    "roll_v1_ws16": (dict(patch_size=4, window_size=16, shift_size=4, shift_strategy="nest_roll",
                          rel_pos_bias="flat", embed_dim=48, depths=[2, 2], num_heads=[3, 6],
                          dim_in=12 * 16 * 16, f_in=3, f_out=10, base_pix=12), 2),
    "grid_cos_v2_ws16": (dict(patch_size=4, window_size=16, shift_size=4, shift_strategy="nest_grid_shift",
                              rel_pos_bias="flat", embed_dim=32, depths=[2, 2], num_heads=[2, 4],
                              use_cos_attn=True, use_v2_norm_placement=True,
                              dim_in=8 * 32 * 32, f_in=3, f_out=5, base_pix=8), 2),
    "ring_cos_v2_ws16": (dict(patch_size=4, window_size=16, shift_size=4, shift_strategy="ring_shift",
                              rel_pos_bias="flat", embed_dim=32, depths=[2, 2], num_heads=[2, 4],
                              use_cos_attn=True, use_v2_norm_placement=True,
                              dim_in=8 * 32 * 32, f_in=3, f_out=5, base_pix=8), 2),
    # production tile shape (ws=64, head_dim=32) at a small sphere: exercises the tcgen05 path
    "ring_cos_v2_ws64": (dict(patch_size=4, window_size=64, shift_size=4, shift_strategy="ring_shift",
                              rel_pos_bias="flat", embed_dim=96, depths=[2, 2], num_heads=[3, 6],
                              use_cos_attn=True, use_v2_norm_placement=True,
                              dim_in=8 * 32 * 32, f_in=3, f_out=4, base_pix=8), 1),
    "roll_v1_ws64": (dict(patch_size=4, window_size=64, shift_size=4, shift_strategy="nest_roll",
                          rel_pos_bias="flat", embed_dim=96, depths=[2, 2], num_heads=[3, 6],
                          dim_in=12 * 32 * 32, f_in=3, f_out=4, base_pix=12), 1),
    # no relative-position bias, window larger than the deepest stage (block shrinks its window)
    "roll_nobias_shrink": (dict(patch_size=4, window_size=64, shift_size=8, shift_strategy="nest_roll",
                                rel_pos_bias=None, embed_dim=32, depths=[2, 2, 2], num_heads=[1, 2, 4],
                                dim_in=8 * 16 * 16, f_in=1, f_out=3, base_pix=8), 1),
}

GRAD_KEYS = (
    "patch_embed.proj.weight",
    "layers.0.blocks.1.attn.qkv.weight",
    "layers.0.blocks.1.attn.qkv.bias",
    "layers.0.blocks.1.attn.proj.weight",
    "layers.0.blocks.1.attn.relative_position_bias_table",
    "layers.0.blocks.1.attn.logit_scale",
    "layers.0.blocks.1.norm1.weight",
    "layers.0.blocks.1.mlp.fc1.weight",
    "layers.0.downsample.reduction.weight",
    "layers.0.downsample.norm.weight",
    "decoder.layers_up.0.expand.weight",
    "decoder.layers_up.0.norm.bias",
    "decoder.up.expand.weight",
    "decoder.output.weight",
)

INDEX_CASES = [
    # (strategy, nside_tokens, base_pix, ws, shift)
    ("nest_roll", 8, 12, 16, 4),
    ("nest_roll", 64, 12, 64, 4),
    ("nest_grid_shift", 8, 8, 4, 0),
    ("nest_grid_shift", 16, 8, 16, 0),
    ("nest_grid_shift", 64, 8, 64, 0),
    ("nest_grid_shift", 128, 8, 64, 0),
    ("ring_shift", 8, 8, 4, 2),
    ("ring_shift", 16, 8, 16, 4),
    ("ring_shift", 64, 8, 64, 4),
    ("ring_shift", 128, 8, 64, 4),
    ("ring_shift", 32, 8, 16, 7),
]


def sha(a: np.ndarray) -> str:
    return hashlib.sha1(np.ascontiguousarray(a.astype("<i8")).tobytes()).hexdigest()


def weights_checksum(sd) -> float:
    return float(sum(v.double().abs().sum() for v in sd.values()))


def build_reference_model(hp_t, DataSpec, kw):
    cfgkw = {k: v for k, v in kw.items() if k not in ("dim_in", "f_in", "f_out", "base_pix")}
    cfg = hp_t.SwinHPTransformerConfig(**cfgkw, drop_path_rate=0.0)
    spec = DataSpec(dim_in=kw["dim_in"], f_in=kw["f_in"], f_out=kw["f_out"], base_pix=kw["base_pix"],
                    class_names=[str(i) for i in range(kw["f_out"])])
    return hp_t.SwinHPTransformerSys(cfg, data_spec=spec)


def main():
    os.makedirs(OUT, exist_ok=True)
    hp_t, hp_s, hp_w, flat, DataSpec = import_reference()
    torch.set_num_threads(8)

    # ---------------- index tables (bit-exact) ----------------
    idx = {}
    for strat, nside, bp, ws, sh in INDEX_CASES:
        N = bp * nside * nside
        if strat == "nest_roll":
            r = hp_s.NestRollShift(sh, N, ws)
            probe = torch.arange(N, dtype=torch.float64)[None, :, None]
            fwd = r.shift(probe)[0, :, 0].long().numpy()
            bwd = r.shift_back(probe)[0, :, 0].long().numpy()
            mask = r.get_mask().numpy()
        elif strat == "nest_grid_shift":
            r = hp_s.NestGridShift(nside, bp, ws)
            r._test_get_offset_dir1()  # the reference's own known answers (hp_shifting.py:148-160)
            r._test_shifted_idcs_dir1()
            r._test_shifted_idcs_dir2()
            fwd, bwd = r.shift_idcs.numpy(), r.back_shift_idcs.numpy()
            mask = r.get_mask().numpy()
        else:
            r = hp_s.RingShift(nside, bp, ws, sh)
            fwd, bwd = r.shift_idcs.numpy(), r.back_shift_idcs.numpy()
            mask = r.get_mask().numpy()
        key = f"{strat}_{nside}_{bp}_{ws}_{sh}"
        idx[key + "_fwd_sha"] = np.array(sha(fwd))
        idx[key + "_bwd_sha"] = np.array(sha(bwd))
        idx[key + "_mask_sha"] = np.array(hashlib.sha1(np.ascontiguousarray(mask.astype("<f4")).tobytes()).hexdigest())
        idx[key + "_fwd_head"] = fwd[:64].astype(np.int64)
        idx[key + "_nmasked"] = np.array(int((np.abs(mask).reshape(mask.shape[0], -1).max(1) > 0).sum()))
        if N <= 4096:
            idx[key + "_fwd"] = fwd.astype(np.int64)
            idx[key + "_bwd"] = bwd.astype(np.int64)
    for ws in (4, 16, 64, 256):
        idx[f"nest_win_idcs_{ws}"] = hp_w.get_nest_win_idcs(ws).numpy()
        wa = hp_t.WindowAttention(8, ws, 2, rel_pos_bias="flat")
        idx[f"rel_pos_index_{ws}"] = wa.relative_position_index.numpy().astype(np.int16)
    np.savez_compressed(os.path.join(OUT, "index_tables.npz"), **idx)
    print("index_tables.npz:", len(idx), "entries")

    # ---------------- model-level fwd/bwd ----------------
    for name, (kw, B) in MODEL_CASES.items():
        cfg = O.HPConfig(**kw)
        sd = O.synth_state_dict(cfg, seed=1234)
        model = build_reference_model(hp_t, DataSpec, kw)
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not unexpected, unexpected
        assert all(("attn_mask" in m) or ("relative_position_index" in m) for m in missing), missing
        model.train()  # all drop rates are 0 -> deterministic; train-mode graph like the benchmark
        g = torch.Generator().manual_seed(99)
        x = torch.randn(B, kw["f_in"], kw["dim_in"], generator=g)
        y = model(x)
        wgt = torch.randn(y.shape, generator=g)
        (y * wgt).sum().backward()
        out = {
            "x": x.numpy(), "y": y.detach().numpy(), "wgt": wgt.numpy(),
            "weights_checksum": np.array(weights_checksum(sd)),
        }
        params = dict(model.named_parameters())
        for k in GRAD_KEYS:
            if k in params and params[k].grad is not None:
                out["grad:" + k] = params[k].grad.numpy()
        # oracle self-check while we are here
        with torch.no_grad():
            yo = O.hp_unet_forward(x, sd, cfg)
        rel = float((yo - y.detach()).norm() / y.detach().norm())
        print(f"{name}: y {tuple(y.shape)} oracle-vs-reference rel {rel:.2e}")
        assert rel < 1e-5
        np.savez_compressed(os.path.join(OUT, f"model_{name}.npz"), **out)

    # ---------------- op-level: WindowAttention / PatchMerging / PatchExpand ----------------
    g = torch.Generator().manual_seed(7)
    ops = {}
    for tag, (C, h, ws, cos, nW, B) in {
        "attn_scaled_ws64": (96, 3, 64, False, 4, 2),
        "attn_cos_ws64": (96, 3, 64, True, 4, 2),
        "attn_cos_ws16_d16": (48, 3, 16, True, 8, 1),
    }.items():
        wa = hp_t.WindowAttention(C, ws, h, rel_pos_bias="flat", use_cos_attn=cos)
        with torch.no_grad():
            for p_ in wa.parameters():
                p_.copy_(torch.randn(p_.shape, generator=g) * 0.2)
            if cos:
                wa.logit_scale.copy_(np.log(10.0) + 0.3 * torch.randn(h, 1, 1, generator=g))
        groups = torch.randint(0, 3, (nW * ws,), generator=g)
        mask = torch.from_numpy(O.attn_mask_from_groups(groups.numpy(), ws))
        x = torch.randn(B * nW, ws, C, generator=g, requires_grad=True)
        y = wa(x, mask=mask)
        wgt = torch.randn(y.shape, generator=g)
        (y * wgt).sum().backward()
        ops[tag + ":x"] = x.detach().numpy()
        ops[tag + ":groups"] = groups.numpy().astype(np.int8)
        ops[tag + ":y"] = y.detach().numpy()
        ops[tag + ":wgt"] = wgt.numpy()
        ops[tag + ":dx"] = x.grad.numpy()
        for n_, p_ in wa.named_parameters():
            ops[tag + ":p:" + n_] = p_.detach().numpy()
            ops[tag + ":g:" + n_] = p_.grad.numpy()
    for tag, (C, N, B) in {"merge_c96": (96, 256, 2), "merge_c32": (32, 64, 1)}.items():
        m = hp_t.PatchMerging(C)
        with torch.no_grad():
            for p_ in m.parameters():
                p_.copy_(torch.randn(p_.shape, generator=g) * 0.2 + (1.0 if p_.ndim == 1 else 0.0))
        x = torch.randn(B, N, C, generator=g, requires_grad=True)
        y = m(x)
        wgt = torch.randn(y.shape, generator=g)
        (y * wgt).sum().backward()
        ops[tag + ":x"], ops[tag + ":y"], ops[tag + ":wgt"], ops[tag + ":dx"] = (
            x.detach().numpy(), y.detach().numpy(), wgt.numpy(), x.grad.numpy())
        for n_, p_ in m.named_parameters():
            ops[tag + ":p:" + n_] = p_.detach().numpy()
            ops[tag + ":g:" + n_] = p_.grad.numpy()
    for tag, (C, N, B) in {"expand_c192": (192, 64, 2), "expand_c64": (64, 16, 1)}.items():
        m = hp_t.PatchExpand(C)
        with torch.no_grad():
            for p_ in m.parameters():
                p_.copy_(torch.randn(p_.shape, generator=g) * 0.2 + (1.0 if p_.ndim == 1 else 0.0))
        x = torch.randn(B, N, C, generator=g, requires_grad=True)
        y = m(x)
        wgt = torch.randn(y.shape, generator=g)
        (y * wgt).sum().backward()
        ops[tag + ":x"], ops[tag + ":y"], ops[tag + ":wgt"], ops[tag + ":dx"] = (
            x.detach().numpy(), y.detach().numpy(), wgt.numpy(), x.grad.numpy())
        for n_, p_ in m.named_parameters():
            ops[tag + ":p:" + n_] = p_.detach().numpy()
            ops[tag + ":g:" + n_] = p_.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "ops.npz"), **ops)
    print("ops.npz:", len(ops), "entries")


if __name__ == "__main__":
    main()
