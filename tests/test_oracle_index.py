"""CPU: the oracle's index layer against the committed reference-generated fixtures, healpy's
documented known answers, and -- when /root/reference is mounted -- the reference itself."""
import hashlib
import os

import numpy as np
import pytest

from oracle import hp_oracle as O
from oracle.make_golden import INDEX_CASES
from oracle.ref_import import reference_available


def sha_i64(a):
    return hashlib.sha1(np.ascontiguousarray(np.asarray(a).astype("<i8")).tobytes()).hexdigest()


def sha_f32(a):
    return hashlib.sha1(np.ascontiguousarray(np.asarray(a).astype("<f4")).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def idx(golden_dir):
    return np.load(os.path.join(golden_dir, "index_tables.npz"))


def test_healpy_known_answers():
    # healpy docstring values
    assert O.nest2ring(16, 1130) == 1504
    assert list(O.nest2ring(2, np.arange(10))) == [13, 5, 4, 0, 15, 7, 6, 1, 17, 9]
    assert list(O.ring2nest(2, np.arange(10))) == [3, 7, 11, 15, 2, 1, 6, 5, 10, 9]
    assert [O.nest2ring(n, 11) for n in (1, 2, 4, 8)] == [11, 2, 12, 211]
    assert [O.ring2nest(n, 11) for n in (1, 2, 4, 8)] == [11, 13, 61, 253]


@pytest.mark.parametrize("nside", [1, 2, 4, 8, 32, 256])
def test_nest_ring_are_inverse_permutations(nside):
    p = np.arange(12 * nside * nside)
    r = O.nest2ring(nside, p)
    assert np.array_equal(np.sort(r), p)
    assert np.array_equal(O.ring2nest(nside, r), p)


@pytest.mark.parametrize("ws", [4, 16, 64, 256])
def test_window_tables_match_reference_fixture(idx, ws):
    assert np.array_equal(O.get_nest_win_idcs(ws), idx[f"nest_win_idcs_{ws}"])
    assert np.array_equal(O.relative_position_index(ws), idx[f"rel_pos_index_{ws}"].astype(np.int64))


def _tables(strat, nside, bp, ws, sh):
    N = bp * nside * nside
    if strat == "nest_roll":
        return O.nest_roll_tables(sh, N, ws)
    if strat == "nest_grid_shift":
        return O.nest_grid_tables(nside, bp, ws)
    return O.ring_shift_tables(nside, bp, ws, sh)


@pytest.mark.parametrize("case", INDEX_CASES, ids=lambda c: "_".join(map(str, c)))
def test_shift_tables_match_reference_fixture(idx, case):
    strat, nside, bp, ws, sh = case
    t = _tables(*case)
    key = f"{strat}_{nside}_{bp}_{ws}_{sh}"
    assert sha_i64(t.shift_idcs) == str(idx[key + "_fwd_sha"])
    assert sha_i64(t.back_idcs) == str(idx[key + "_bwd_sha"])
    assert np.array_equal(t.shift_idcs[:64], idx[key + "_fwd_head"])
    mask = O.attn_mask_from_groups(t.groups, ws)
    assert sha_f32(mask) == str(idx[key + "_mask_sha"])
    assert int((np.abs(mask).reshape(mask.shape[0], -1).max(1) > 0).sum()) == int(idx[key + "_nmasked"])
    if key + "_fwd" in idx.files:
        assert np.array_equal(t.shift_idcs, idx[key + "_fwd"])
        assert np.array_equal(t.back_idcs, idx[key + "_bwd"])


def test_reference_known_answers_for_grid_offsets():
    # hp_shifting.py:148-160 (_test_get_offset_dir1), nside=128, ws=64 -> 256 windows per base pixel
    g = O._GridShift(128, 8, 64)
    for w, want in ((2, 1), (3, 1), (6, 1), (7, 1), (8, 5), (9, 5), (10, 1), (11, 1), (12, 5), (32, 21)):
        assert g.offset_dir1(w) == want
    assert g.offset_dir1(0) // g.wpb == 2
    # survey probe: first entries and sha of the nside=128 table
    t = O.nest_grid_tables(128, 8, 64)
    assert list(t.shift_idcs[:8]) == list(range(16368, 16376))
    assert sha_i64(t.shift_idcs)[:16] == "49749ddb28dc9d2e"
    assert int((O.attn_mask_from_groups(t.groups, 64).reshape(2048, -1).min(1) < 0).sum()) == 128


def test_ring_shift_rejects_what_the_reference_rejects():
    with pytest.raises((ValueError, IndexError, KeyError)):
        O.ring_shift_tables(8, 4, 16, 4)
    with pytest.raises((ValueError, IndexError, KeyError)):
        O.ring_shift_tables(8, 6, 16, 4)


@pytest.mark.skipif(not reference_available(), reason="/root/reference not mounted (GPU box)")
def test_oracle_against_live_reference():
    import torch
    from oracle.ref_import import import_reference

    hp_t, hp_s, hp_w, flat, _ = import_reference()
    for ws in (4, 16, 64):
        assert np.array_equal(hp_w.get_nest_win_idcs(ws).numpy(), O.get_nest_win_idcs(ws))
    r = hp_s.NestGridShift(32, 8, 16)
    o = O.nest_grid_tables(32, 8, 16)
    assert np.array_equal(r.shift_idcs.numpy(), o.shift_idcs)
    assert np.array_equal(r.get_mask(False).numpy().astype(np.int64), o.groups)
    r = hp_s.RingShift(32, 8, 16, 4)
    o = O.ring_shift_tables(32, 8, 16, 4)
    assert np.array_equal(r.shift_idcs.numpy(), o.shift_idcs)
    assert np.array_equal(r.back_shift_idcs.numpy(), o.back_idcs)
    assert np.array_equal(r.get_mask(False).numpy(), o.groups)
    r = hp_s.NestRollShift(4, 768, 16)
    o = O.nest_roll_tables(4, 768, 16)
    x = torch.arange(768)[None, :, None].double()
    assert np.array_equal(r.shift(x)[0, :, 0].long().numpy(), o.shift_idcs)
    assert np.array_equal(r.get_mask().numpy(), O.attn_mask_from_groups(o.groups, 16))
