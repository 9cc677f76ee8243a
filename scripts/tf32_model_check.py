#!/usr/bin/env python
"""Whole-network forward error of the product model vs the CPU oracle with the library GEMMs in fp32 and in TF32
(torch.backends.cuda.matmul.allow_tf32), on the committed fixture models and on a deeper random model."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import hp_oracle as O  # noqa: E402
from tests.util import build_product_model, load_model_case, rel_err  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    cases = []
    for name in ("ring_cos_v2_ws64", "roll_v1_ws64"):
        kw, cfg, sd, gold = load_model_case(name)
        cases.append((name, kw, cfg, sd, torch.from_numpy(gold["x"])))
    # deeper: the BASELINE depth profile at a small sphere
    kw = dict(patch_size=4, window_size=64, shift_size=4, shift_strategy="nest_roll", rel_pos_bias="flat", embed_dim=96,
              depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], use_cos_attn=True, use_v2_norm_placement=True,
              dim_in=12 * 64 * 64, f_in=3, f_out=10, base_pix=12)
    cfg = O.HPConfig(**kw)
    sd = O.synth_state_dict(cfg, seed=7)
    cases.append(("deep_2262_nside64", kw, cfg, sd, torch.randn(1, 3, kw["dim_in"], generator=torch.Generator().manual_seed(3))))
    for name, kw, cfg, sd, x in cases:
        with torch.no_grad():
            want = O.hp_unet_forward(x, sd, cfg)
        model = build_product_model(kw, sd, dev).eval()
        for tf32 in (False, True):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            with torch.no_grad():
                got = model(x.to(dev)).cpu()
            print(f"{name}: library GEMMs {'tf32' if tf32 else 'fp32'}: forward rel err vs oracle {rel_err(got, want):.3e}", flush=True)


if __name__ == "__main__":
    main()
