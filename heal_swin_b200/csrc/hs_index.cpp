// Host-side HEALPix index layer: nested-window tables, shift permutations and SW-MSA mask
// group ids.  Integer work only; bit-exact against the reference (tests/test_index_layer.py).
//
// Replaces the init-time Python of
//   hp_windowing.get_nest_win_idcs            hp_windowing.py:43-62
//   WindowAttention.relative_position_index   swin_hp_transformer.py:98-114
//   NestRollShift / NestGridShift / RingShift hp_shifting.py:42-404
//   healpy.pixelfunc.nest2ring / ring2nest    (third-party, called at hp_shifting.py:329,333)
#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

#include "hs_common.h"

namespace {

inline int64_t floordiv(int64_t a, int64_t b) {
  int64_t q = a / b, r = a % b;
  return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q;
}
inline int64_t floormod(int64_t a, int64_t b) { return a - floordiv(a, b) * b; }

inline uint64_t compact_even(uint64_t v) {
  v &= 0x5555555555555555ull;
  v = (v | (v >> 1)) & 0x3333333333333333ull;
  v = (v | (v >> 2)) & 0x0F0F0F0F0F0F0F0Full;
  v = (v | (v >> 4)) & 0x00FF00FF00FF00FFull;
  v = (v | (v >> 8)) & 0x0000FFFF0000FFFFull;
  v = (v | (v >> 16)) & 0x00000000FFFFFFFFull;
  return v;
}
inline uint64_t spread(uint64_t v) {
  v &= 0x00000000FFFFFFFFull;
  v = (v | (v << 16)) & 0x0000FFFF0000FFFFull;
  v = (v | (v << 8)) & 0x00FF00FF00FF00FFull;
  v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0Full;
  v = (v | (v << 2)) & 0x3333333333333333ull;
  v = (v | (v << 1)) & 0x5555555555555555ull;
  return v;
}
inline int64_t isqrt64(int64_t v) {
  int64_t r = (int64_t)std::floor(std::sqrt((double)v));
  while (r * r > v) --r;
  while ((r + 1) * (r + 1) <= v) ++r;
  return r;
}

const int64_t kJrll[12] = {2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4};
const int64_t kJpll[12] = {1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7};

// HEALPix NESTED -> RING (Gorski et al. 2005; same maths as healpix_base::nest2ring)
int64_t nest2ring1(int64_t nside, int64_t p) {
  const int64_t npface = nside * nside, npix = 12 * npface;
  const int64_t ncap = 2 * nside * (nside - 1), nl4 = 4 * nside;
  const int64_t face = p / npface, ipf = p % npface;
  const int64_t ix = (int64_t)compact_even((uint64_t)ipf);
  const int64_t iy = (int64_t)compact_even((uint64_t)ipf >> 1);
  const int64_t jr = kJrll[face] * nside - ix - iy - 1;
  int64_t nr, n_before, kshift;
  if (jr < nside) {
    nr = jr; n_before = 2 * nr * (nr - 1); kshift = 0;
  } else if (jr > 3 * nside) {
    nr = nl4 - jr; n_before = npix - 2 * (nr + 1) * nr; kshift = 0;
  } else {
    nr = nside; n_before = ncap + (jr - nside) * nl4; kshift = (jr - nside) & 1;
  }
  int64_t jp = floordiv(kJpll[face] * nr + ix - iy + 1 + kshift, 2);
  if (jp > nl4) jp -= nl4;
  if (jp < 1) jp += nl4;
  return n_before + jp - 1;
}

// HEALPix RING -> NESTED (healpix_base::ring2xyf + xyf2nest)
int64_t ring2nest1(int64_t nside, int64_t p) {
  const int64_t npface = nside * nside, npix = 12 * npface;
  const int64_t ncap = 2 * nside * (nside - 1), nl2 = 2 * nside, nl4 = 4 * nside;
  int64_t iring, iphi, kshift, nr, face;
  if (p < ncap) {
    iring = (1 + isqrt64(1 + 2 * p)) >> 1;
    iphi = (p + 1) - 2 * iring * (iring - 1);
    kshift = 0; nr = iring;
    face = (iphi - 1) / nr;
  } else if (p < npix - ncap) {
    const int64_t ip = p - ncap, tmp = ip / nl4;
    iring = tmp + nside;
    iphi = ip - nl4 * tmp + 1;
    kshift = (iring + nside) & 1; nr = nside;
    const int64_t ire = iring - nside + 1, irm = nl2 + 2 - ire;
    const int64_t ifm = floordiv(iphi - ire / 2 + nside - 1, nside);
    const int64_t ifp = floordiv(iphi - irm / 2 + nside - 1, nside);
    face = (ifp == ifm) ? (ifp | 4) : ((ifp < ifm) ? ifp : (ifm + 8));
  } else {
    const int64_t ip = npix - p;
    const int64_t ir = (1 + isqrt64(2 * ip - 1)) >> 1;
    iphi = 4 * ir + 1 - (ip - 2 * ir * (ir - 1));
    kshift = 0; nr = ir; iring = 2 * nl2 - ir;
    face = 8 + (iphi - 1) / nr;
  }
  const int64_t irt = iring - kJrll[face] * nside + 1;
  int64_t ipt = 2 * iphi - kJpll[face] * nr - kshift - 1;
  if (ipt >= nl2) ipt -= 8 * nside;
  const int64_t ix = (ipt - irt) >> 1, iy = (-(ipt + irt)) >> 1;
  return face * npface + (int64_t)spread((uint64_t)ix) + ((int64_t)spread((uint64_t)iy) << 1);
}

inline bool is_pow2(int64_t v) { return v > 0 && (v & (v - 1)) == 0; }
inline bool is_pow4(int64_t v) { return is_pow2(v) && (__builtin_ctzll((unsigned long long)v) % 2 == 0); }

// in-window nested index t -> Cartesian (row, col); quadrant order of hp_windowing.py:54-58
inline void nest_rowcol(int t, int S, int* row, int* col) {
  *row = (int)compact_even((uint64_t)t >> 1);
  *col = (S - 1) - (int)compact_even((uint64_t)t);
}

// ---------------------------------------------------------------- NestGridShift, hp_shifting.py:76-306
struct GridShift {
  int64_t ws, npix, n_windows, wpb;  // wpb: windows per base pixel (a power of 4)
  int top;                           // log4(wpb)
  static const int kDir1Base[8];     // hp_shifting.py:126
  static const int kDir2Base[8];     // hp_shifting.py:196

  int scale(int64_t w) const {       // hp_shifting.py:104-115
    int64_t s = wpb; int lg = top;
    while (floormod(w, s) != 0) { s /= 4; --lg; }
    return lg;
  }
  static int64_t pow4(int e) { return (int64_t)1 << (2 * e); }
  int64_t offset_dir1(int64_t w) const {  // hp_shifting.py:117-146 (unit: windows)
    int sc;
    for (;;) {
      sc = scale(w);
      w -= pow4(sc);
      if (sc >= scale(w)) break;
    }
    int64_t off = 0;
    for (int q = 0; q <= sc; ++q) off += pow4(q);
    if (sc == top) {
      w += pow4(sc);
      off += (int64_t)(kDir1Base[floordiv(w, wpb)] - 1) * wpb;
    }
    return off;
  }
  int64_t offset_dir2(int64_t w) const {  // hp_shifting.py:189-212 (unit: windows)
    int sc = scale(w);
    while (floordiv(floormod(w, pow4(sc + 1)), pow4(sc)) == 2) {
      w -= 2 * pow4(sc);
      sc = scale(w);
    }
    int64_t off = 0;
    for (int q = 0; q < sc; ++q) off += 2 * pow4(q);
    if (sc == top) off += (int64_t)kDir2Base[floordiv(w, wpb)] * wpb;
    return off;
  }
};
const int GridShift::kDir1Base[8] = {2, 2, 2, 6, 3, 3, 3, 3};
const int GridShift::kDir2Base[8] = {3, 3, 3, 3, 3, 3, 3, 3};

int grid_tables(int64_t nside, int base_pix, int ws, std::vector<int64_t>* idcs, std::vector<int8_t>* groups) {
  HS_REQUIRE(base_pix == 8, "NestGridShift is currently only implemented for 8 base pixels");
  GridShift g;
  g.ws = ws;
  g.npix = (int64_t)base_pix * nside * nside;
  g.n_windows = g.npix / ws;
  g.wpb = (g.npix / base_pix) / ws;
  HS_REQUIRE(g.wpb >= 1 && is_pow4(g.wpb) && ws % 4 == 0,
             "nest_grid_shift needs nside^2/window_size to be a power of 4 (got nside=%lld ws=%d)",
             (long long)nside, ws);
  g.top = __builtin_ctzll((unsigned long long)g.wpb) / 2;
  const int64_t h = ws / 2, q = ws / 4, npix = g.npix;
  if (idcs) {
    std::vector<int64_t> d1(npix), d2(npix);
    for (int64_t w = 0; w < g.n_windows; ++w) {
      const int64_t first = w * ws;
      const int64_t e1 = first - g.offset_dir1(w) * ws;  // hp_shifting.py:170-178
      for (int64_t i = 0; i < h; ++i) {
        d1[first + i] = floormod(e1 - h + i, npix);
        d1[first + h + i] = first + i;
      }
      const int64_t e2 = first - g.offset_dir2(w) * ws;  // hp_shifting.py:233-247
      for (int64_t i = 0; i < q; ++i) {
        d2[first + i] = floormod(e2 - h - q + i, npix);
        d2[first + q + i] = first + i;
        d2[first + h + i] = floormod(e2 - q + i, npix);
        d2[first + h + q + i] = first + h + i;
      }
    }
    idcs->resize(npix);
    for (int64_t p = 0; p < npix; ++p) (*idcs)[p] = d1[d2[p]];  // hp_shifting.py:90-91
  }
  if (groups) {  // hp_shifting.py:261-300
    groups->assign(npix, 0);
    const int64_t span = g.wpb * ws;
    // left_mask_subset recursion == windows whose base-4 digits are all in {0,1};
    // right_mask_subset == digits all in {0,2}
    for (int b = 4; b < 8; ++b) {
      for (int64_t w = 0; w < g.wpb; ++w) {
        bool left = true, right = true;
        for (int64_t t = w, lvl = 0; lvl < g.top; ++lvl, t >>= 2) {
          const int dgt = (int)(t & 3);
          left = left && (dgt == 0 || dgt == 1);
          right = right && (dgt == 0 || dgt == 2);
        }
        const int64_t first = b * span + w * ws;
        if (left) for (int64_t i = 0; i < h; ++i) (*groups)[first + i] = (int8_t)(b + 1);
        if (right) {  // applied after the left marks, like the reference's call order
          for (int64_t i = 0; i < q; ++i) {
            (*groups)[first + i] = (int8_t)(b + 1 + 4);
            (*groups)[first + h + i] = (int8_t)(b + 1 + 4);
          }
        }
      }
      const int64_t first_co = (int64_t)(b - 4) * span;
      for (int64_t i = 0; i < q; ++i) (*groups)[first_co + i] = (int8_t)(b + 1);
    }
  }
  return HS_OK;
}

// ---------------------------------------------------------------- RingShift, hp_shifting.py:309-404
int ring_tables(int64_t nside, int base_pix, int shift, std::vector<int64_t>* idcs_out,
                std::vector<int8_t>* groups) {
  // The reference only survives base_pix == 8: np.concatenate([]) raises for <= 4, lost_pix[7]
  // is out of range for 5..7, GET_LOST_FROM has no key >= 8 (hp_shifting.py:354-366).
  HS_REQUIRE(base_pix == 8, "ring_shift is only defined for 8 base pixels (reference raises for %d)", base_pix);
  const int64_t face = nside * nside, npix = (int64_t)base_pix * face, full = 12 * face;
  std::vector<int64_t> res(npix);
  for (int64_t p = 0; p < npix; ++p)  // hp_shifting.py:327-334
    res[p] = ring2nest1(nside, floormod(nest2ring1(nside, p) - shift, full));
  const int64_t top = npix - 1;
  if (groups) groups->assign(npix, 0);
  std::vector<char> used(full, 0);
  for (int64_t p = 0; p < npix; ++p) {
    used[res[p]] = 1;
    if (groups && res[p] > top) (*groups)[p] = (int8_t)(p / face + 1);  // hp_shifting.py:339-344
  }
  if (!idcs_out) return HS_OK;
  std::vector<std::vector<int64_t>> lost(base_pix);  // hp_shifting.py:346-349 (setdiff1d: ascending)
  for (int b = 0; b < base_pix; ++b)
    for (int64_t p = b * face; p < (b + 1) * face; ++p)
      if (!used[p]) lost[b].push_back(p);
  static const int kLostFrom[8] = {-1, -1, -1, -1, 7, 4, 5, 6};  // hp_shifting.py:354
  std::vector<int64_t> spare;
  for (int b = 4; b < base_pix; ++b) {  // hp_shifting.py:356-366
    const std::vector<int64_t>& src = lost[kLostFrom[b]];
    size_t n = 0;
    for (int64_t p = b * face; p < (b + 1) * face; ++p)
      if (res[p] > top) {
        HS_REQUIRE(n < src.size(), "for base pixel %d, there were not enough source pixel", b);
        res[p] = src[n++];
      }
    spare.insert(spare.end(), src.begin() + n, src.end());
  }
  int64_t still = 0;
  for (int64_t p = 0; p < npix; ++p) still += res[p] > top;
  HS_REQUIRE((int64_t)spare.size() == still,
             "the number of unused source pixels does not match the number of pixels to be filled");
  size_t cur = 0;
  for (int64_t p = 0; p < 4 * face && p < npix; ++p)  // hp_shifting.py:372-378
    if (res[p] > top) res[p] = spare[cur++];
  idcs_out->swap(res);
  return HS_OK;
}

int check_permutation(const std::vector<int64_t>& idcs, int64_t nside, int ws) {
  std::vector<char> seen(idcs.size(), 0);
  for (int64_t v : idcs) {
    HS_REQUIRE(v >= 0 && v < (int64_t)idcs.size() && !seen[v],
               "shift validation failed for nside=%lld, window_size=%d", (long long)nside, ws);
    seen[v] = 1;
  }
  return HS_OK;
}

}  // namespace

extern "C" {

int hs_nest2ring(int64_t nside, const int64_t* in, int64_t* out, int64_t n) {
  HS_REQUIRE(nside >= 1 && in && out && n >= 0, "hs_nest2ring: bad arguments");
  for (int64_t i = 0; i < n; ++i) {
    HS_REQUIRE(in[i] >= 0 && in[i] < 12 * nside * nside, "hs_nest2ring: pixel %lld out of range", (long long)in[i]);
    out[i] = nest2ring1(nside, in[i]);
  }
  return HS_OK;
}

int hs_ring2nest(int64_t nside, const int64_t* in, int64_t* out, int64_t n) {
  HS_REQUIRE(nside >= 1 && in && out && n >= 0, "hs_ring2nest: bad arguments");
  for (int64_t i = 0; i < n; ++i) {
    HS_REQUIRE(in[i] >= 0 && in[i] < 12 * nside * nside, "hs_ring2nest: pixel %lld out of range", (long long)in[i]);
    out[i] = ring2nest1(nside, in[i]);
  }
  return HS_OK;
}

int hs_nest_win_idcs(int ws, int64_t* out) {
  HS_REQUIRE(ws >= 4 && out, "hs_nest_win_idcs: bad arguments");
  const int S = (int)std::sqrt((double)ws);
  HS_REQUIRE(is_pow2(S), "hs_nest_win_idcs: sqrt(window_size) must be a power of 2 (window_size=%d)", ws);
  for (int t = 0; t < S * S; ++t) {
    int r, c;
    nest_rowcol(t, S, &r, &c);
    out[r * S + c] = t;
  }
  return HS_OK;
}

int hs_rel_pos_index(int ws, int64_t* out) {
  HS_REQUIRE(ws >= 1 && out, "hs_rel_pos_index: bad arguments");
  const int S = (int)std::sqrt((double)ws);
  HS_REQUIRE(S * S == ws && is_pow2(S), "rel_pos_bias='flat' needs a power-of-4 window_size (got %d)", ws);
  std::vector<int> row(ws), col(ws);
  for (int t = 0; t < ws; ++t) nest_rowcol(t, S, &row[t], &col[t]);
  for (int i = 0; i < ws; ++i)
    for (int j = 0; j < ws; ++j)
      out[(int64_t)i * ws + j] = (int64_t)(row[i] - row[j] + S - 1) * (2 * S - 1) + (col[i] - col[j] + S - 1);
  return HS_OK;
}

int hs_shift_tables(int strategy, int64_t nside, int base_pix, int ws, int shift, int64_t* shift_idcs,
                    int64_t* back_idcs, int8_t* groups) {
  HS_REQUIRE(nside >= 1 && base_pix >= 1 && base_pix <= 12, "hs_shift_tables: bad nside/base_pix");
  HS_REQUIRE(is_pow2(ws), "window_size must be a power of 2 (got %d)", ws);
  const int64_t N = (int64_t)base_pix * nside * nside;
  HS_REQUIRE(N % ws == 0, "window_size %d does not divide the %lld pixels", ws, (long long)N);
  std::vector<int64_t> idcs;
  std::vector<int8_t> grp;
  switch (strategy) {
    case HS_SHIFT_NONE:
      idcs.resize(N);
      std::iota(idcs.begin(), idcs.end(), (int64_t)0);
      grp.assign(N, 0);
      break;
    case HS_SHIFT_NEST_ROLL: {  // hp_shifting.py:48-73
      HS_REQUIRE(shift > 0 && shift < N, "nest_roll: bad shift_size %d", shift);
      idcs.resize(N);
      for (int64_t p = 0; p < N; ++p) idcs[p] = (p + shift) % N;
      grp.assign(N, 0);
      for (int64_t p = std::max<int64_t>(N - ws, 0); p < N - shift; ++p) grp[p] = 1;
      for (int64_t p = N - shift; p < N; ++p) grp[p] = 2;
      break;
    }
    case HS_SHIFT_NEST_GRID: {
      int rc = grid_tables(nside, base_pix, ws, (shift_idcs || back_idcs) ? &idcs : nullptr, groups ? &grp : nullptr);
      if (rc) return rc;
      break;
    }
    case HS_SHIFT_RING: {
      int rc = ring_tables(nside, base_pix, shift, (shift_idcs || back_idcs) ? &idcs : nullptr, groups ? &grp : nullptr);
      if (rc) return rc;
      break;
    }
    default:
      return hs::fail(HS_ERR_ARG, "unknown shift strategy %d", strategy);
  }
  if (shift_idcs || back_idcs) {
    int rc = check_permutation(idcs, nside, ws);  // hp_shifting.py:96-99, 385-388
    if (rc) return rc;
  }
  if (shift_idcs) std::copy(idcs.begin(), idcs.end(), shift_idcs);
  if (back_idcs)
    for (int64_t p = 0; p < N; ++p) back_idcs[idcs[p]] = p;  // argsort of a permutation
  if (groups) std::copy(grp.begin(), grp.end(), groups);
  return HS_OK;
}

int hs_attn_mask_from_groups(const int8_t* groups, int64_t N, int ws, float* mask) {
  // any window size: the flat (lat-lon) twin takes rectangular windows such as 4 x 6; only the HEALPix tables above need 2^k
  HS_REQUIRE(groups && mask && ws > 0 && N % ws == 0, "hs_attn_mask_from_groups: bad arguments");
  for (int64_t w = 0; w < N / ws; ++w)
    for (int i = 0; i < ws; ++i)
      for (int j = 0; j < ws; ++j)
        mask[(w * ws + i) * ws + j] = (groups[w * ws + i] != groups[w * ws + j]) ? -100.0f : 0.0f;
  return HS_OK;
}

}  // extern "C"
