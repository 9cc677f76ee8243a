"""TEST INFRASTRUCTURE ONLY -- import the *real* reference from /root/reference.

Only usable in the build container (the GPU box has no /root/reference).  Used by
``oracle/make_golden.py`` to generate ``tests/golden/*.npz`` and by the
``-m "not gpu"`` tests that pin ``oracle/hp_oracle.py`` against the reference
when it happens to be present.  Nothing in the product package imports this.

The reference needs two third-party modules that are not in this image
(SURVEY.md section 8c):

* ``timm.models.layers`` (timm 0.4.12: ``DropPath``, ``trunc_normal_``,
  ``to_2tuple``) -- imported at swin_hp_transformer.py:14, swin_transformer.py:13
* ``healpy`` (1.15.2) -- imported at hp_shifting.py:5; only
  ``hp.pixelfunc.ring2nest`` / ``nest2ring`` are touched (hp_shifting.py:329,333)

Both are shimmed here.  The healpy shim is backed by the oracle's own integer
HEALPix index maps (``oracle.hp_oracle.nest2ring`` / ``ring2nest``), which are
pinned against healpy's documented known answers in tests/test_oracle_index.py.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("HEALSWIN_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "heal_swin", "models_torch"))


def _install_shims():
    import torch
    import torch.nn as nn

    if "timm" not in sys.modules:
        timm = types.ModuleType("timm")
        models = types.ModuleType("timm.models")
        layers = types.ModuleType("timm.models.layers")

        class DropPath(nn.Module):
            """Stochastic depth per sample (timm 0.4.12 semantics)."""

            def __init__(self, drop_prob=None):
                super().__init__()
                self.drop_prob = drop_prob

            def forward(self, x):
                if not self.drop_prob or not self.training:
                    return x
                keep = 1.0 - self.drop_prob
                shape = (x.shape[0],) + (1,) * (x.ndim - 1)
                rnd = keep + torch.rand(shape, dtype=x.dtype, device=x.device)
                return x.div(keep) * rnd.floor_()

        def to_2tuple(v):
            return tuple(v) if isinstance(v, (tuple, list)) else (v, v)

        layers.DropPath = DropPath
        layers.trunc_normal_ = torch.nn.init.trunc_normal_
        layers.to_2tuple = to_2tuple
        timm.models = models
        models.layers = layers
        sys.modules["timm"] = timm
        sys.modules["timm.models"] = models
        sys.modules["timm.models.layers"] = layers

    if "healpy" not in sys.modules:
        from oracle import hp_oracle

        hp = types.ModuleType("healpy")
        pixelfunc = types.ModuleType("healpy.pixelfunc")
        pixelfunc.ring2nest = hp_oracle.ring2nest
        pixelfunc.nest2ring = hp_oracle.nest2ring
        hp.pixelfunc = pixelfunc
        hp.ring2nest = hp_oracle.ring2nest
        hp.nest2ring = hp_oracle.nest2ring
        sys.modules["healpy"] = hp
        sys.modules["healpy.pixelfunc"] = pixelfunc


def import_reference():
    """Returns the reference's (swin_hp_transformer, hp_shifting, hp_windowing,
    swin_transformer, DataSpec) modules/classes, imported read-only."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    sys.dont_write_bytecode = True  # /root/reference is read-only
    import importlib

    hp_t = importlib.import_module("heal_swin.models_torch.swin_hp_transformer")
    hp_s = importlib.import_module("heal_swin.models_torch.hp_shifting")
    hp_w = importlib.import_module("heal_swin.models_torch.hp_windowing")
    flat = importlib.import_module("heal_swin.models_torch.swin_transformer")
    spec = importlib.import_module("heal_swin.data.segmentation.data_spec")
    return hp_t, hp_s, hp_w, flat, spec.DataSpec
