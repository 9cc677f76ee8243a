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
    from heal_swin_b200.factory import build_hp_model

    return build_hp_model(kw, sd, device)
