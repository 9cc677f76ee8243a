"""Construction helper shared by bench.py, smoke() and the tests: a SwinHPTransformerSys from one flat keyword dict
(the SwinHPTransformerConfig fields plus the four DataSpec fields), optionally loaded with a reference-layout state
dict (swin_hp_transformer.py:821-969; data/segmentation/data_spec.py:5-11)."""
from .data_spec import DataSpec

_SPEC_KEYS = ("dim_in", "f_in", "f_out", "base_pix")


def build_hp_model(kw, state_dict=None, device="cpu"):
    from .models_torch import swin_hp_transformer as M

    cfgkw = {k: v for k, v in kw.items() if k not in _SPEC_KEYS}
    cfgkw.setdefault("drop_path_rate", 0.0)
    cfg = M.SwinHPTransformerConfig(**cfgkw)
    spec = DataSpec(**{k: kw[k] for k in _SPEC_KEYS})
    model = M.SwinHPTransformerSys(cfg, data_spec=spec)
    if state_dict is not None:
        missing, unexpected = model.load_state_dict(state_dict, strict=False)
        assert not unexpected, unexpected
        # the mask / index buffers are rebuilt from the index layer; everything else must be present
        assert all(("attn_mask" in m) or ("relative_position_index" in m) for m in missing), missing
    return model.to(device)
