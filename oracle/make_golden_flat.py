"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/flat_*.npz from the REAL reference flat SWIN-UNet
(heal_swin/models_torch/swin_transformer.py, imported read-only via oracle/ref_import.py).

    python -m oracle.make_golden_flat        # build container only (needs /root/reference)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import flat_oracle as FO  # noqa: E402
from oracle.ref_import import import_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# name -> (FlatConfig kwargs, batch)
FLAT_CASES = {
    # BASELINE.json configs[4] tile shape (8x8 window, head_dim 32, shift 2, patch 2) on a small image
    "cos_v2_ws8": (dict(patch_size=[2, 2], window_size=[8, 8], shift_size=[2, 2], embed_dim=32, depths=[2, 2],
                        num_heads=[1, 2], use_cos_attn=True, use_v2_norm_placement=True, dim_in=(64, 64),
                        f_in=3, f_out=5), 2),
    # asymmetric shift: the reference rolls by (s0, s0) on the way in and (s0, s1) on the way out
    "v1_ws4_asym": (dict(patch_size=[2, 2], window_size=[4, 4], shift_size=[2, 1], embed_dim=16, depths=[2, 2],
                         num_heads=[2, 4], dim_in=(32, 48), f_in=1, f_out=3), 1),
    # no mask, no relative position bias, default shift (-1 -> half window), deepest stage == one window
    "nomask_norel": (dict(patch_size=[2, 2], window_size=[8, 8], shift_size=[4, 4], embed_dim=32, depths=[2, 2],
                          num_heads=[1, 2], use_masking=False, use_rel_pos_bias=False, dim_in=(32, 32),
                          f_in=2, f_out=2), 2),
    # non-power-of-two window (4 x 6 = 24 tokens) and a deepest stage that collapses to a single window (no shift)
    "v1_ws4x6": (dict(patch_size=[2, 2], window_size=[4, 6], shift_size=[2, 3], embed_dim=16, depths=[2, 2, 2],
                      num_heads=[1, 2, 4], dim_in=(32, 48), f_in=2, f_out=4), 2),
}

FLAT_GRAD_KEYS = (
    "patch_embed.proj.weight",
    "layers.0.blocks.1.attn.qkv.weight",
    "layers.0.blocks.1.attn.relative_position_bias_table",
    "layers.0.blocks.1.attn.logit_scale",
    "layers.0.blocks.1.norm1.weight",
    "layers.0.downsample.reduction.weight",
    "layers_up.0.expand.weight",
    "layers_up.1.blocks.1.attn.proj.weight",
    "up.expand.weight",
    "output.weight",
)


def weights_checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values()))


def build_reference_flat(flat, DataSpec, kw):
    cfgkw = {k: v for k, v in kw.items() if k not in ("dim_in", "f_in", "f_out")}
    cfg = flat.SwinTransformerConfig(**cfgkw, drop_path_rate=0.0)
    spec = DataSpec(dim_in=tuple(kw["dim_in"]), f_in=kw["f_in"], f_out=kw["f_out"], base_pix=None,
                    class_names=[str(i) for i in range(kw["f_out"])])
    return flat.SwinTransformerSys(cfg, data_spec=spec)


def main():
    hp_t, hp_s, hp_w, flat, DataSpec = import_reference()
    torch.set_num_threads(8)
    for name, (kw, B) in FLAT_CASES.items():
        cfg = FO.FlatConfig(**kw)
        sd = FO.synth_state_dict(cfg, seed=4321)
        model = build_reference_flat(flat, DataSpec, kw)
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not unexpected, unexpected
        assert all(("attn_mask" in m) or ("relative_position_index" in m) for m in missing), missing
        model.train()
        g = torch.Generator().manual_seed(77)
        x = torch.randn(B, kw["f_in"], *kw["dim_in"], generator=g)
        y = model(x)
        wgt = torch.randn(y.shape, generator=g)
        (y * wgt).sum().backward()
        out = {"x": x.numpy(), "y": y.detach().numpy(), "wgt": wgt.numpy(),
               "weights_checksum": np.array(weights_checksum(sd))}
        params = dict(model.named_parameters())
        for k in FLAT_GRAD_KEYS:
            if k in params and params[k].grad is not None:
                out["grad:" + k] = params[k].grad.numpy()
        # masks / index buffers of the first shifted block, for the bit-exact index checks
        blk = model.layers[0].blocks[1]
        if blk.attn_mask is not None:
            out["attn_mask_l0b1"] = blk.attn_mask.numpy().astype(np.float32)
        out["rel_pos_index"] = blk.attn.relative_position_index.numpy().astype(np.int16)
        with torch.no_grad():
            yo = FO.flat_unet_forward(x, sd, cfg)
        rel = float((yo - y.detach()).norm() / y.detach().norm())
        print(f"flat_{name}: y {tuple(y.shape)} oracle-vs-reference rel {rel:.2e}")
        assert rel < 1e-5
        np.savez_compressed(os.path.join(OUT, f"flat_{name}.npz"), **out)


if __name__ == "__main__":
    main()
