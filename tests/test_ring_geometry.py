"""Independent pin of nest2ring / ring2nest (healpy 1.15.2 is absent; RingShift at hp_shifting.py:327-334 depends on them).

The C++ maps (csrc/hs_index.cpp) and the oracle (oracle/hp_oracle.py) both restate the integer algorithm of the HEALPix
library.  This file derives the RING index of every NESTED pixel WITHOUT that algorithm, from the geometric definition
(Gorski et al. 2005; planar form: Calabretta & Roukema 2007):

  * a NESTED index is (face f, ix, iy) with ix / iy the even / odd bits of the in-face index;
  * in the HEALPix plane the 12 base faces are unit diamonds: faces 0-3 centred at (X, Y) = (pi/4 + f pi/2, +pi/4), 4-7 at
    ((f-4) pi/2, 0), 8-11 at (pi/4 + (f-8) pi/2, -pi/4); the centre of pixel (ix, iy) is displaced by
    ((x - y), (x + y - 1)) * pi/4 with x = (ix + 1/2) / nside, y = (iy + 1/2) / nside;
  * latitude depends on Y only (z = 8 Y / (3 pi) for |Y| <= pi/4, z = +-(1 - sigma^2 / 3), sigma = 2 - 4|Y|/pi beyond), and
    in the polar caps the longitude is phi = X_c + (X - X_c) / sigma around the centre meridian X_c of the facet;
  * RING ordering numbers the pixels along iso-latitude rings from north to south, and by increasing phi inside a ring.

So the ring index is the RANK of the pixel centre under the key (-Y, phi).  Everything is done in exact integer / rational
arithmetic (units of pi / (8 nside)), so there are no ties to break.  Also here: the reference's in-code asserts of
RingShift (hp_shifting.py:361-363, 369-371, 385-388) as explicit tests at nside 128."""
from fractions import Fraction

import numpy as np
import pytest

from heal_swin_b200 import _lib, hp_index
from oracle import hp_oracle as O


def _compact(v):
    """every second bit of v, packed"""
    out = np.zeros_like(v)
    for b in range(16):
        out |= ((v >> (2 * b)) & 1) << b
    return out


def geometric_nest2ring(nside):
    n = nside
    p = np.arange(12 * n * n, dtype=np.int64)
    face, ipf = p // (n * n), p % (n * n)
    ix, iy = _compact(ipf), _compact(ipf >> 1)
    u, v = 2 * ix + 1, 2 * iy + 1                       # x, y in units of 1 / (2 nside)
    # plane coordinates in units of pi / (8 nside): face half-diagonal pi/4 = 2 nside units
    row = face // 4                                      # 0 north, 1 equatorial, 2 south
    xc = np.where(row == 1, 4 * n * (face - 4), 2 * n + 4 * n * (face % 4))
    yc = np.where(row == 0, 2 * n, np.where(row == 1, 0, -2 * n))
    X = (xc + (u - v)) % (16 * n)
    Y = yc + (u + v - 2 * n)
    keys = []
    for Xi, Yi in zip(X.tolist(), Y.tolist()):
        if abs(Yi) <= 2 * n:                             # equatorial zone: phi = X
            phi = Fraction(Xi)
        else:                                            # polar cap: phi = X_c + (X - X_c) / sigma, sigma = (4n - |Y|) / (2n)
            Xc = 2 * n + 4 * n * (Xi // (4 * n))
            phi = Xc + Fraction((Xi - Xc) * 2 * n, 4 * n - abs(Yi))
        keys.append((-Yi, phi))
    order = sorted(range(len(keys)), key=keys.__getitem__)   # order[r] = nested index of ring pixel r
    assert len({keys[i] for i in order}) == len(order), "two pixel centres coincide"
    ring = np.empty(len(order), dtype=np.int64)
    ring[np.asarray(order)] = np.arange(len(order))
    return ring


@pytest.mark.parametrize("nside", [1, 2, 4, 8, 16, 32, 64])
def test_nest2ring_equals_the_geometric_ring_rank(nside):
    want = geometric_nest2ring(nside)
    p = np.arange(12 * nside * nside)
    assert np.array_equal(hp_index.nest2ring(nside, p), want)        # product (C++)
    assert np.array_equal(O.nest2ring(nside, p), want)               # oracle
    inv = np.empty_like(want)
    inv[want] = p
    assert np.array_equal(hp_index.ring2nest(nside, p), inv)
    assert np.array_equal(O.ring2nest(nside, p), inv)


def test_ring_structure_properties_at_nside_128():
    """Size-independent properties at the N_side of the bench model's first stage (N_side 256 with patch size 4 -> 128):
    mutual inverses, the polar rings, healpy's documented answers."""
    nside = 128
    npix = 12 * nside * nside
    p = np.arange(npix)
    r = hp_index.nest2ring(nside, p)
    assert np.array_equal(np.sort(r), p)
    assert np.array_equal(hp_index.ring2nest(nside, r), p)
    # the first ring holds the 4 pixels touching the north pole: the last nested pixel of faces 0-3
    assert sorted(hp_index.ring2nest(nside, np.arange(4)).tolist()) == [(f + 1) * nside * nside - 1 for f in range(4)]
    # ... and the last ring the first nested pixel of faces 8-11
    assert sorted(hp_index.ring2nest(nside, np.arange(npix - 4, npix)).tolist()) == [f * nside * nside for f in range(8, 12)]
    assert hp_index.nest2ring(16, [1130])[0] == 1504                       # healpy docstring
    assert hp_index.nest2ring(2, np.arange(10)).tolist() == [13, 5, 4, 0, 15, 7, 6, 1, 17, 9]
    assert hp_index.ring2nest(2, np.arange(10)).tolist() == [3, 7, 11, 15, 2, 1, 6, 5, 10, 9]


@pytest.mark.parametrize("nside,base_pix,ws,shift", [(128, 8, 64, 4), (128, 8, 64, 32), (64, 8, 16, 8), (32, 8, 64, 2)])
def test_ring_shift_reference_asserts_hold(nside, base_pix, ws, shift):
    """hp_shifting.py:361-363 ("not enough source pixel"), :369-371 ("number of unused source pixels"), :385-388
    (shift_idcs is a permutation), recomputed from the product's own maps and checked on the product's table."""
    npix = base_pix * nside * nside
    size = nside * nside
    rolled = np.roll(np.arange(12 * nside * nside), shift)
    result = hp_index.ring2nest(nside, rolled)[hp_index.nest2ring(nside, np.arange(npix))]
    outside = result > npix - 1
    lost = [np.setdiff1d(np.arange(i * size, (i + 1) * size), result) for i in range(base_pix)]
    take_from = {4: 7, 5: 4, 6: 5, 7: 6}
    unused = 0
    for i in range(4, base_pix):
        need = int(outside[i * size:(i + 1) * size].sum())
        have = len(lost[take_from[i]])
        assert need <= have, f"for base pixel {i}, there were not enough source pixel"
        unused += have - need
    assert unused == int(outside[: 4 * size].sum()), "unused source pixels do not match the pixels to be filled"
    fwd, back, grp = hp_index.shift_tables(_lib.SHIFT_RING, nside, base_pix, ws, shift)
    fwd, back, grp = fwd.numpy(), back.numpy(), grp.numpy()
    assert np.array_equal(np.sort(fwd), np.arange(npix))                # :385-388
    assert np.array_equal(fwd[back], np.arange(npix))
    # pixels that stayed inside the domain keep the plain ring-roll source; masked ids are base pixel + 1 (:339-344)
    assert np.array_equal(fwd[~outside], result[~outside])
    want_grp = np.zeros(npix, dtype=np.int64)
    for i in range(base_pix):
        want_grp[i * size:(i + 1) * size][outside[i * size:(i + 1) * size]] = i + 1
    assert np.array_equal(grp.astype(np.int64), want_grp)
