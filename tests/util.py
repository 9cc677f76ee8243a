"""Shared helpers for the parity tests (test infrastructure only)."""
import os

import numpy as np
import torch

from oracle import hp_oracle as O
from oracle.make_golden import MODEL_CASES

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_err(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def load_model_case(name):
    kw, B = MODEL_CASES[name]
    cfg = O.HPConfig(**kw)
    sd = O.synth_state_dict(cfg, seed=1234)
    gold = np.load(os.path.join(GOLDEN, f"model_{name}.npz"))
    chk = float(sum(v.double().abs().sum() for v in sd.values()))
    assert abs(chk - float(gold["weights_checksum"])) <= 1e-6 * chk, "synthetic weight RNG drifted from the fixture"
    return kw, cfg, sd, gold


def build_product_model(kw, sd=None, device="cpu"):
    from heal_swin_b200.data_spec import DataSpec
    from heal_swin_b200.models_torch import swin_hp_transformer as M

    cfgkw = {k: v for k, v in kw.items() if k not in ("dim_in", "f_in", "f_out", "base_pix")}
    cfgkw.setdefault("drop_path_rate", 0.0)
    cfg = M.SwinHPTransformerConfig(**cfgkw)
    spec = DataSpec(dim_in=kw["dim_in"], f_in=kw["f_in"], f_out=kw["f_out"], base_pix=kw["base_pix"])
    model = M.SwinHPTransformerSys(cfg, data_spec=spec)
    if sd is not None:
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not unexpected, unexpected
        assert all(("attn_mask" in m) or ("relative_position_index" in m) for m in missing), missing
    return model.to(device)
