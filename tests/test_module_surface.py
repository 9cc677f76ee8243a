"""CPU: the drop-in module surface (names, signatures, state-dict layout, error behaviour)."""
import inspect

import pytest
import torch

from heal_swin_b200.data_spec import DataSpec
from heal_swin_b200.models_torch import hp_shifting, hp_windowing, swin_hp_transformer as M
from oracle import hp_oracle as O
from oracle.make_golden import MODEL_CASES
from oracle.ref_import import reference_available
from tests.util import build_product_model


def test_public_names_exist():
    for n in ("Mlp", "WindowAttention", "SwinTransformerBlock", "SwinHPTransformerBlock", "PatchMerging",
              "PatchExpand", "PatchExpanding", "FinalPatchExpand_X4", "BasicLayer", "BasicLayer_up", "PatchEmbed",
              "UnetDecoder", "SwinHPTransformerConfig", "SwinHPTransformerSys"):
        assert hasattr(M, n), n
    for n in ("window_partition", "window_reverse", "get_nest_win_idcs"):
        assert hasattr(hp_windowing, n) and hasattr(M, n)
    for n in ("get_attn_mask_from_mask", "NoShift", "NestRollShift", "NestGridShift", "RingShift"):
        assert hasattr(hp_shifting, n)


def test_constructor_signatures_follow_the_reference():
    # SURVEY.md 8b
    sig = list(inspect.signature(M.WindowAttention.__init__).parameters)
    assert sig == ["self", "dim", "window_size", "num_heads", "rel_pos_bias", "qkv_bias", "qk_scale", "attn_drop",
                   "proj_drop", "use_cos_attn"]
    sig = list(inspect.signature(M.SwinTransformerBlock.__init__).parameters)
    assert sig == ["self", "dim", "input_resolution", "base_pix", "num_heads", "window_size", "shift_size",
                   "shift_strategy", "rel_pos_bias", "mlp_ratio", "qkv_bias", "qk_scale", "drop", "attn_drop",
                   "drop_path", "act_layer", "norm_layer", "use_v2_norm_placement", "use_cos_attn"]
    sig = list(inspect.signature(M.BasicLayer.__init__).parameters)
    assert sig[:10] == ["self", "dim", "input_resolution", "depth", "num_heads", "window_size", "base_pix",
                        "shift_size", "shift_strategy", "rel_pos_bias"]
    assert list(inspect.signature(M.PatchMerging.__init__).parameters) == ["self", "dim", "dim_scale", "norm_layer"]
    assert list(inspect.signature(M.PatchExpand.__init__).parameters) == ["self", "dim", "dim_scale", "norm_layer"]
    cfg = M.SwinHPTransformerConfig()
    assert (cfg.patch_size, cfg.window_size, cfg.shift_size, cfg.shift_strategy, cfg.embed_dim) == (4, 4, 2, "nest_roll", 96)
    assert cfg.drop_path_rate == 0.1 and cfg.decoder_class is M.UnetDecoder


@pytest.mark.parametrize("name", list(MODEL_CASES))
def test_state_dict_layout(name):
    kw, _ = MODEL_CASES[name]
    model = build_product_model(kw)
    sd = model.state_dict()
    want = O.synth_state_dict(O.HPConfig(**kw))
    params = {k: v for k, v in sd.items() if "attn_mask" not in k and "relative_position_index" not in k}
    assert set(params) == set(want)
    for k in want:
        assert tuple(params[k].shape) == tuple(want[k].shape), k
    assert not any("_hs_" in k for k in sd), "private kernel tables must not leak into checkpoints"
    ws = kw["window_size"]
    for k, v in sd.items():
        if k.endswith("attn_mask"):
            assert v.shape[1:] == (ws, ws) or v.shape[1] == v.shape[2]
        if k.endswith("relative_position_index"):
            assert v.dtype == torch.int64 and tuple(v.shape) == (ws, ws)


def test_reference_quirks_are_kept():
    m = build_product_model(MODEL_CASES["grid_cos_v2_ws16"][0])
    blk = m.layers[0].blocks[1]
    assert float(blk.attn.relative_position_bias_table.abs().max()) == 0.0  # zero-init, :92-96,:121
    assert torch.allclose(blk.attn.logit_scale, torch.log(torch.tensor(10.0)))
    assert set(blk.attn_mask.unique().tolist()) == {0.0, -100.0}  # -100, not -inf (hp_shifting.py:25)
    assert m.layers[0].blocks[0].attn_mask is None
    shr = build_product_model(MODEL_CASES["roll_nobias_shrink"][0])
    last = shr.layers[2].blocks[1]
    assert last.window_size == 32 and last.shift_size == 0 and last.attn.window_size == 64  # :243-251


def test_cpu_forward_fails_loudly():
    m = build_product_model(MODEL_CASES["roll_v1_ws16"][0])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 12 * 16 * 16))


def test_views():
    x = torch.arange(2 * 32 * 3, dtype=torch.float32).reshape(2, 32, 3)
    w = hp_windowing.window_partition(x, 8)
    assert w.shape == (8, 8, 3) and w.data_ptr() == x.data_ptr()
    assert torch.equal(hp_windowing.window_reverse(w, 8, 32), x)
    with pytest.raises(AssertionError):
        hp_windowing.window_partition(x, 6)
    with pytest.raises(AssertionError):
        M.PatchMerging(3)(torch.zeros(1, 6, 3))


@pytest.mark.skipif(not reference_available(), reason="/root/reference not mounted (GPU box)")
def test_state_dict_keys_equal_live_reference():
    from oracle.ref_import import import_reference

    hp_t, _, _, _, RefSpec = import_reference()
    for name in ("ring_cos_v2_ws16", "grid_cos_v2_ws16", "roll_v1_ws16"):
        kw, _ = MODEL_CASES[name]
        cfgkw = {k: v for k, v in kw.items() if k not in ("dim_in", "f_in", "f_out", "base_pix")}
        ref = hp_t.SwinHPTransformerSys(hp_t.SwinHPTransformerConfig(**cfgkw),
                                        data_spec=RefSpec(kw["dim_in"], kw["f_in"], kw["f_out"], kw["base_pix"], []))
        mine = build_product_model(kw)
        a, b = ref.state_dict(), mine.state_dict()
        assert list(a) == list(b)
        for k in a:
            assert a[k].shape == b[k].shape and a[k].dtype == b[k].dtype, k
            if "attn_mask" in k or "relative_position_index" in k:
                assert torch.equal(a[k], b[k]), k
        mine.load_state_dict(a, strict=True)
