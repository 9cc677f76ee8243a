"""CPU: the product's C-ABI index layer (csrc/hs_index.cpp) is bit-exact against the oracle and
the reference-generated fixtures."""
import hashlib
import os

import numpy as np
import pytest

from heal_swin_b200 import _lib, hp_index
from oracle import hp_oracle as O
from oracle.make_golden import INDEX_CASES

CODES = {"nest_roll": _lib.SHIFT_NEST_ROLL, "nest_grid_shift": _lib.SHIFT_NEST_GRID, "ring_shift": _lib.SHIFT_RING}


def sha_i64(a):
    return hashlib.sha1(np.ascontiguousarray(np.asarray(a).astype("<i8")).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def idx(golden_dir):
    return np.load(os.path.join(golden_dir, "index_tables.npz"))


@pytest.mark.parametrize("ws", [4, 16, 64, 256])
def test_window_tables(idx, ws):
    assert np.array_equal(hp_index.nest_win_idcs(ws).numpy(), idx[f"nest_win_idcs_{ws}"])
    assert np.array_equal(hp_index.rel_pos_index(ws).numpy(), idx[f"rel_pos_index_{ws}"].astype(np.int64))


def test_rel_pos_index_rejects_non_power_of_4():
    with pytest.raises(AssertionError):
        hp_index.rel_pos_index(32)


@pytest.mark.parametrize("nside", [1, 2, 8, 64, 256])
def test_nest_ring(nside):
    p = np.arange(12 * nside * nside)
    assert np.array_equal(hp_index.nest2ring(nside, p), O.nest2ring(nside, p))
    assert np.array_equal(hp_index.ring2nest(nside, p), O.ring2nest(nside, p))
    with pytest.raises(AssertionError):
        hp_index.nest2ring(nside, [12 * nside * nside])


@pytest.mark.parametrize("case", INDEX_CASES, ids=lambda c: "_".join(map(str, c)))
def test_shift_tables_vs_fixture_and_oracle(idx, case):
    strat, nside, bp, ws, sh = case
    fwd, back, grp = hp_index.shift_tables(CODES[strat], nside, bp, ws, sh)
    key = f"{strat}_{nside}_{bp}_{ws}_{sh}"
    assert sha_i64(fwd.numpy()) == str(idx[key + "_fwd_sha"])
    assert sha_i64(back.numpy()) == str(idx[key + "_bwd_sha"])
    mask = hp_index.attn_mask_from_groups(grp, ws).numpy()
    assert hashlib.sha1(mask.astype("<f4").tobytes()).hexdigest() == str(idx[key + "_mask_sha"])
    N = bp * nside * nside
    o = (O.nest_roll_tables(sh, N, ws) if strat == "nest_roll" else
         O.nest_grid_tables(nside, bp, ws) if strat == "nest_grid_shift" else
         O.ring_shift_tables(nside, bp, ws, sh))
    assert np.array_equal(grp.numpy().astype(np.int64), o.groups)


def test_full_size_tables_are_permutations():
    # BASELINE configs[1] stage-0 size: nside 128 tokens (N_side=256 pixels, patch 4), window 64
    for code, bp in ((_lib.SHIFT_NEST_GRID, 8), (_lib.SHIFT_RING, 8), (_lib.SHIFT_NEST_ROLL, 12)):
        fwd, back, grp = hp_index.shift_tables(code, 128, bp, 64, 4)
        N = bp * 128 * 128
        assert np.array_equal(np.sort(fwd.numpy()), np.arange(N))
        assert np.array_equal(fwd.numpy()[back.numpy()], np.arange(N))
        assert grp.numpy().max() <= 12


def test_error_convention():
    with pytest.raises(AssertionError, match="8 base pixels"):
        hp_index.shift_tables(_lib.SHIFT_NEST_GRID, 8, 12, 16, 0)
    with pytest.raises(AssertionError, match="8 base pixels"):
        hp_index.shift_tables(_lib.SHIFT_RING, 8, 6, 16, 4)
    with pytest.raises(AssertionError, match="power of 2"):
        hp_index.shift_tables(_lib.SHIFT_NEST_ROLL, 8, 12, 24, 4)


def test_module_shifters_match_oracle():
    from heal_swin_b200.models_torch import hp_shifting as S

    r = S.NestRollShift(4, 768, 16)
    o = O.nest_roll_tables(4, 768, 16)
    assert np.array_equal(r.shift_idcs.numpy(), o.shift_idcs)
    assert np.array_equal(r.back_shift_idcs.numpy(), o.back_idcs)
    assert np.array_equal(r.get_mask().numpy(), O.attn_mask_from_groups(o.groups, 16))
    g = S.NestGridShift(16, 8, 16)
    o = O.nest_grid_tables(16, 8, 16)
    assert np.array_equal(g.shift_idcs.numpy(), o.shift_idcs)
    assert np.array_equal(g.get_mask(False).numpy().astype(np.int64), o.groups)
    q = S.RingShift(16, 8, 16, 4)
    o = O.ring_shift_tables(16, 8, 16, 4)
    assert np.array_equal(q.shift_idcs.numpy(), o.shift_idcs)
    assert np.array_equal(q.get_mask(False).numpy(), o.groups)
    assert S.NoShift().get_mask() is None


def test_shift_tables_sweep_is_bit_exact_and_rejects_what_the_oracle_rejects():
    """Every strategy over a grid of (nside, base_pix, window, shift), including single-window spheres, 1-pixel shifts and
    windows larger than a base pixel: the C++ tables equal the oracle's bit for bit, and a configuration is refused by
    one exactly when it is refused by the other (the reference's asserts, SURVEY 8b error convention)."""
    checked = rejected = 0
    for strat, code in CODES.items():
        for nside in (2, 4, 8, 16, 32):
            bps = (8,) if strat == "nest_grid_shift" else (1, 2, 4, 8) if strat == "ring_shift" else (1, 8, 12)
            for bp in bps:
                N = bp * nside * nside
                for ws in (4, 16, 64, 256):
                    if ws > N:
                        continue
                    for sh in ((ws // 2,) if strat == "nest_grid_shift" else (1, ws // 4, ws // 2, ws - 1)):
                        try:
                            o = (O.nest_roll_tables(sh, N, ws) if strat == "nest_roll" else
                                 O.nest_grid_tables(nside, bp, ws) if strat == "nest_grid_shift" else
                                 O.ring_shift_tables(nside, bp, ws, sh))
                        except (AssertionError, KeyError, IndexError, ValueError):
                            o = None
                        try:
                            got = hp_index.shift_tables(code, nside, bp, ws, sh)
                        except AssertionError:
                            got = None
                        assert (o is None) == (got is None), (strat, nside, bp, ws, sh)
                        if o is None:
                            rejected += 1
                            continue
                        fwd, back, grp = got
                        assert np.array_equal(fwd.numpy(), o.shift_idcs), (strat, nside, bp, ws, sh)
                        assert np.array_equal(back.numpy(), o.back_idcs), (strat, nside, bp, ws, sh)
                        assert np.array_equal(grp.numpy().astype(np.int64), o.groups), (strat, nside, bp, ws, sh)
                        checked += 1
    assert checked > 200 and rejected > 50
