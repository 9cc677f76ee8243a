// Backward of the windowed attention core on the sm_100a tensor cores (tcgen05 + TMEM + TMA); window = 64
// tokens, head_dim = 32, fp32 in / fp32 out, TF32 operands with fp32 accumulation.  Adjoint of
// hs_attn_tc.cu (reference: autograd of swin_hp_transformer.py:136-171 + the shift / partition / reverse at
// :319-330).  S is recomputed from q, k; the softmax row statistics come from the forward pass (its saved
// log-sum-exp, and rowsum(P o dP) = dO . O from its output), so no statistics sweep over S is needed.
//
// One work unit = one (window, head).  Per unit, with S = q k^T, P = softmax(S*scale + bias + mask), dP = dO v^T,
// dS = P o (dP - rowsum(P o dP)):
//      dV = P^T dO      dQ = scale * dS k      dK = scale * dS^T q            (+ the F.normalize terms for cos attention)
// The TMEM operand of tcgen05.mma is the A matrix with M on the TMEM lanes, so dQ needs dS with the query index on
// the lanes while dV and dK need P^T / dS^T with the key index on the lanes.  Both orientations are produced by the
// tensor cores themselves, stacked into one M = 128 tile (lanes 0-63 "natural", lanes 64-127 "transposed"):
//      D1 = [Q;K] [Q;K]^T    (one N = 128 MMA per K step):   lanes 0-63 x cols 64-127 = S,   lanes 64-127 x cols 0-63 = S^T
//      D2 = [dO;V] [dO;V]^T                                   lanes 0-63 x cols 64-127 = dP,  lanes 64-127 x cols 0-63 = dP^T
// (the other half of each product is unused).  Threads 0-63 of a warpgroup own one query row each (softmax statistics,
// dS row, dbias row), threads 64-127 one key row each (P^T, dS^T from the row statistics published in shared memory).
// They write dS (over S), P^T (over S^T) and dS^T (over dP^T) back to TMEM as TF32, and three more MMAs with
// MN-major B tiles (K, dO, Q) produce dQ, dV, dK into dead column blocks (see the MMA issuer for the exact map).
//
// Every input tile is needed K-major (scores) and, except V, MN-major (outputs): they are fetched twice by TMA with the
// two swizzles (the second fetch hits L2).  HBM traffic per unit: q, k, v, dO in; dq, dk, dv out = 7 x 8 KB (+ the 8 KB
// of the forward output read by the statistics warps).
//
// Warp roles (512 threads, register file split with setmaxnreg 152 / 104 / 104):
//   warps 0-3 / 4-7   two elementwise warpgroups: unit n -> group n & 1, TMEM stage n & 1.  They only sweep: wait for the
//                     scores, write dS / P^T / dS^T back to TMEM, hand the row sums to the epilogue warps (S.part) and go
//                     on with their next unit.
//   warp 8            load producer (row table one unit ahead; TMA loads, or cp.async gathers for shifted windows)
//   warp 9            MMA issuer (one elected thread, polling)
//   warps 10-11       row statistics (-lse from the forward, -delta = -dO . O with dO read from the K-major tile)
//   warps 12-15       epilogue: one TMEM lane quadrant each; scale / F.normalize correction, staging in the unit's (dead)
//                     MN-major tiles, TMA store, slot release.
// Shared memory: a ring of 2 K-major entries (32 KB: [Q;K], [dO;V]; dead as soon as the score MMAs are done) and a ring of 4
// MN-major entries (24 KB: q, k, dO; B operands of the output MMAs, then output staging until the TMA store has read
// them).  With one ring of 3 x 56 KB the long-lived MN tiles limited the units in flight to 3 and the whole chain
// load -> scores -> sweep -> output MMAs -> epilogue -> store ran at one unit per 4.3 k cycles; see DESIGN.md section 3
// for the measurements that led here (profiles/r2z_attn_bwd_*).
#include <cfloat>

#include "hs_common.h"
#include "hs_kernels.h"
#include "hs_sm100.cuh"
#include "hs_tc_common.cuh"

namespace {

using namespace hs::sm100;
using namespace hs::tc;

constexpr int kKS = 2;  // ring of K-major input tiles (live from the load to the end of the score MMAs)
constexpr int kMS = 4;  // ring of MN-major input tiles (live until the TMA store of the outputs staged in them has been read)
constexpr int kStageCols = 256;  // D1 (128) + D2 (128)
constexpr int kTmemCols = 512;
constexpr int kThreads = 512;
constexpr int kEwThreads = 256;  // two elementwise warpgroups
// How the two elementwise warpgroups share the work.  true: both sweep every unit, 32 columns each.  false: warpgroup g
// sweeps all 64 columns of the units with (n & 1) == g.
#ifndef HS_BWD_COOP
#define HS_BWD_COOP 0
#endif
constexpr bool kCoop = HS_BWD_COOP != 0;
#ifndef HS_BWD_PREFETCH
#define HS_BWD_PREFETCH 0
#endif
constexpr bool kL2Prefetch = HS_BWD_PREFETCH != 0;  // L2 prefetch of the next unit's tiles by the producer
// Which half accumulates the dbias tile.  With cos attention the query-row sweep is the heavier one (row sums for the
// F.normalize / logit_scale terms), so the key-row threads take the tile (held transposed, dbt[j][i]); without, the
// query-row threads do (measured on one box, stage 0: cos 1.47 -> 1.40 ms, plain 1.01 -> 1.05 ms the other way round;
// splitting the tile between the halves by 32 x 32 quadrant was slower than either: 1.54 / 1.08 ms).
#ifndef HS_BWD_DBT_KEY_ROWS
#define HS_BWD_DBT_KEY_ROWS 2  // 0: query rows, 1: key rows, 2: key rows for cos attention only
#endif
__host__ __device__ constexpr bool dbt_on_key_rows(bool cos) { return HS_BWD_DBT_KEY_ROWS == 2 ? cos : HS_BWD_DBT_KEY_ROWS != 0; }
// Mirrored stacking on TMEM stage 1: its units are computed as [K;Q] [K;Q]^T and [V;dO] [V;dO]^T, so the key-row
// orientation sits on lanes 0-63 and the query rows on lanes 64-127 -- every TMEM column offset of the unmirrored map
// XOR 64.  A TMEM lane quadrant is tied to a warp (id % 4) and thereby to a scheduler; without the mirror the two
// heavier roles (key-row sweep, dK + dV epilogue) always run on schedulers 2-3.  Stand-alone (random data, stage 0) the
// mirror is 2-4 % faster; inside the training step (22 launches, four stage shapes, shifted blocks with gathered and
// masked windows) it measured 15.2-15.5 ms against 14.0-14.2 ms without, same box, alternating runs
// (profiles/r2z_attn_bwd_variants.log) -- so it is off.
#ifndef HS_BWD_MIRROR
#define HS_BWD_MIRROR 0
#endif
constexpr bool kMirror = !kCoop && HS_BWD_MIRROR != 0;
constexpr int kSweepCols = kCoop ? kWS / 2 : kWS;  // columns swept by one thread
constexpr int kSweepers = kCoop ? 2 : 1;           // threads per row
constexpr int kDbtPitch = 64;   // floats; 16-byte chunk c4 of row r is stored at chunk (c4 ^ (r & 15)): conflict-free float4 RMW
// Output path of the epilogue.  true: every thread writes its 128-byte output row(s) straight from registers (full
// cache lines); the slot is released immediately.  false: stage the tiles in the slot and TMA-store them (the slot then
// stays occupied until the TMA engine has read the staging tiles).  Measured at stage 0: direct 1.71 ms, TMA store 1.20 ms.
constexpr bool kDirectStore = false;
constexpr int kBiasPitch = 68;  // floats; 16-byte chunk index advances by 17 per row -> conflict-free LDS.128

struct SlotMeta {
  int rows[kWS];        // global row (b * N + token) of every slot of the unit
  uint8_t groups[kWS];  // mask group ids
  int flags;
  int pad[3];
};

struct KSlot {
  uint8_t qk[2 * kTile];   // [Q;K]  K-major, SWIZZLE_128B
  uint8_t dov[2 * kTile];  // [dO;V] K-major, SWIZZLE_128B
};
struct Slot {
  uint8_t q_mn[kTile];     // MN-major (SWIZZLE_128B_ATOM_32B); reused as dQ staging
  uint8_t k_mn[kTile];     //   "                                reused as dK staging
  uint8_t do_mn[kTile];    //   "                                reused as dV staging
};

struct Smem {
  KSlot kslot[kKS];
  Slot slot[kMS];
  float bias[kWS * kBiasPitch];     // bias[i][j] * log2(e); query rows read it row-wise (LDS.128), key rows column-wise
  SlotMeta meta[kMS];
  float inv[kMS][2 * kWS];  // per slot (statistics warps, from the forward): [0,64) 1/max(|q_i|,eps), [64,128) 1/max(|k_j|,eps)
  float nlse[kMS][kWS];    // per slot (written by the statistics warp): MINUS the log2-domain log-sum-exp of every query row
  float ndelta[kMS][kWS];  // per slot: MINUS rowsum(P o dP) = -dO_i . O_i
  float4 part[2][2][2 * kWS];  // per TMEM stage and warpgroup: every row's partial sums (ds_sweep), elementwise -> epilogue warps
  float dbt[2][kWS * kDbtPitch];  // per warpgroup: dbt[i][j] = sum over its units of dS[i][j] (query-row thread i owns row i)
  uint64_t kfull[kKS], kempty[kKS];
  uint64_t full[kMS], empty[kMS], meta_ready[kMS], stats_ready[kMS];
  uint64_t s_ready[2], dsn_ready[2], dst_ready[2], o_ready[2], stage_free[2];
  uint32_t tmem_base;
};

struct BwdArgs {
  const float* qkv;
  const float* out;  // forward output (B, N, C)
  const float* lse;  // forward statistics (P, H, B*N): plane 0 log2-domain log-sum-exp; planes 1, 2 (cos) 1/|q|, 1/|k|
  const float* dout;
  float* dqkv;
  const int32_t* src;
  const uint8_t* groups;
  const float* bias;         // (H, 64, 64) or null
  const float* logit_scale;  // (H) or null
  float* dbias;              // (H, 64, 64) or null, accumulated
  float* dlogit;             // (H) or null, accumulated
  float scale;
  float fix1, fix2;  // TF32 truncation compensation (hs_tc_common.cuh), 1.0 when disabled
  uint32_t drop_thresh;  // attention-probability dropout (0 = off), see hs_common.h
  float drop_scale;
  uint64_t seed;
  int B, nW, C, H, cos;
  long long N;
  int total;  // B * nW units per head
};

#ifdef HS_BWD_TRACE
// Diagnostics build only (tools/trace_bwd.cu): per-phase clock64 stamps of CTA (0, 0), role x unit x point.
constexpr int kTraceUnits = 48, kTracePoints = 8, kTraceRoles = 8;  // roles: wg0 nat, wg0 tr, wg1 nat, wg1 tr, mma, producer, epilogue nat, epilogue tr
__device__ long long g_trace[kTraceRoles * kTraceUnits * kTracePoints];
#define HS_TRACE(role, n, k)                                                               \
  do {                                                                                     \
    if (blockIdx.x == 0 && blockIdx.y == 0 && (n) < kTraceUnits)                           \
      g_trace[((role) * kTraceUnits + (n)) * kTracePoints + (k)] = clock64();              \
  } while (0)
#else
#define HS_TRACE(role, n, k) do {} while (0)
#endif


// ---------------------------------------------------------------------------------------------------------------
// Elementwise stage of one unit.  A ROLLED loop over 8-column chunks with everything recomputed from TMEM (no per-row
// register arrays): the fully unrolled first version (4000 instructions, 64 KB) spent half of its issue slots waiting
// for instruction fetches (profiles/r1d_attn_bwd_tc_*).  TMEM loads are double-buffered: the next chunk is in flight
// while the current one is used.  The sweep is specialised at compile time on the thread's orientation (query row /
// key row) and on the attention variant (cos, bias), and does its fp32 arithmetic two columns at a time with the
// packed FMUL2 / FFMA2 / FADD2 instructions of sm_100: the first version (run-time flags, scalar math) issued 29
// instructions per element and was bound by the issue slots of the two elementwise warps per scheduler
// (profiles/r1l_attn_bwd_full_summary.txt: 0.48 IPC, tensor pipe 19 %).
struct RowCtx {
  uint32_t s_src, dp_src;  // my 64-column block of D1 (S or S^T) and D2 (dP or dP^T), lane field included
  uint32_t p_dst, ds_dst;  // where P^T (key rows only) and dS / dS^T go (TMEM, lane field included)
  uint32_t oinv, brow;     // shared-memory addresses: normalisation of the other index (cos), my bias row / column
  uint32_t nlse_v, ndelta_v;  // key rows: shared addresses of the (negated) statistics vectors over the query index
  uint32_t groups;         // shared-memory address of the unit's 64 group ids
  uint32_t dbt_row;        // shared address of my row of the dbias tile (0 = none / the other half accumulates it)
  int dbt_xor;             //   16-byte chunk c4 of row r lives at chunk (c4 ^ (r & 15))
  float row_scale;         // log2(e) * scale (* my 1/|row| for cos) * truncation fix
  float nlse, ndelta;      // query rows: minus my log-sum-exp / minus my rowsum(P o dP)
  float fix2;
  int my_group;
  bool masked;
  // attention dropout: element (i, j) of the unit; my row is index `drop_r`, columns run over the other index
  uint32_t drop_key, drop_thresh;  // thresh 0 = off
  float drop_scale;
  int drop_r;
};

constexpr int kCW = 8;  // columns per chunk of the elementwise loops (16 was measured slower: 1.44 vs 1.27 ms)

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2u(uint32_t a, uint32_t b) { return make_float2(__uint_as_float(a), __uint_as_float(b)); }

// all threads: p = exp2(logit - lse), dS = p (dP - delta); dS (scaled by the other index' 1/norm for cos) goes back to
// TMEM as the TF32 A operand of dQ (query rows, over S) / dK (key rows, over dP^T); key rows also write P^T (over S^T)
// for dV; query rows accumulate dS into their warpgroup's dbias tile.  Returns sum_c dS_c * raw_c, centred (cos only).
template <bool kNat, bool kCos, bool kBias, bool kDrop>
__device__ __forceinline__ float4 ds_sweep(const RowCtx& R, const int cb) {
  // rs = sum_c dS_c * w_c with w = raw * (1/norm of the other index) (the cos logit up to my row's scale).  In exact
  // arithmetic sum_c dS_c = 0 along a query row, so any constant may be subtracted from w: the P-weighted mean of w is
  // subtracted (sds * pw) so that an error of the row's delta (it now comes from the forward output) is not amplified.
  float2 rs[2] = {f2(0.f, 0.f), f2(0.f, 0.f)}, sds = f2(0.f, 0.f), pw = f2(0.f, 0.f);
  const float2 rsc = f2(R.row_scale, R.row_scale), fx2 = f2(R.fix2, R.fix2);
  const float2 nl_own = f2(R.nlse, R.nlse), nd_own = f2(R.ndelta, R.ndelta);
  auto chunk = [&](const uint32_t (&sraw)[kCW], const uint32_t (&dpr)[kCW], int c0) {
    float2 raw[4], x[4], ov[4], nl[4], nd[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) raw[p] = f2u(sraw[2 * p], sraw[2 * p + 1]);
    if (kCos) {
      const float4 a4 = lds_f4(R.oinv + 4 * c0), b4 = lds_f4(R.oinv + 4 * c0 + 16);
      ov[0] = f2(a4.x, a4.y); ov[1] = f2(a4.z, a4.w); ov[2] = f2(b4.x, b4.y); ov[3] = f2(b4.z, b4.w);
    }
    if (kNat) {
#pragma unroll
      for (int p = 0; p < 4; ++p) { nl[p] = nl_own; nd[p] = nd_own; }
    } else {
      const float4 a4 = lds_f4(R.nlse_v + 4 * c0), b4 = lds_f4(R.nlse_v + 4 * c0 + 16);
      const float4 c4 = lds_f4(R.ndelta_v + 4 * c0), d4 = lds_f4(R.ndelta_v + 4 * c0 + 16);
      nl[0] = f2(a4.x, a4.y); nl[1] = f2(a4.z, a4.w); nl[2] = f2(b4.x, b4.y); nl[3] = f2(b4.z, b4.w);
      nd[0] = f2(c4.x, c4.y); nd[1] = f2(c4.z, c4.w); nd[2] = f2(d4.x, d4.y); nd[3] = f2(d4.z, d4.w);
    }
    // log2-domain logits
#pragma unroll
    for (int p = 0; p < 4; ++p) x[p] = __fmul2_rn(raw[p], rsc);
    if (kBias) {
      float2 bv[4];
      if (kNat) {
        const float4 a4 = lds_f4(R.brow + 4 * c0), b4 = lds_f4(R.brow + 4 * c0 + 16);
        bv[0] = f2(a4.x, a4.y); bv[1] = f2(a4.z, a4.w); bv[2] = f2(b4.x, b4.y); bv[3] = f2(b4.z, b4.w);
      } else {  // key rows: bias[c][r] walks down a column; consecutive lanes hit consecutive banks
        const uint32_t bp = R.brow + 4 * kBiasPitch * c0;
#pragma unroll
        for (int p = 0; p < 4; ++p)
          bv[p] = f2(lds_f1(bp + 4 * kBiasPitch * (2 * p)), lds_f1(bp + 4 * kBiasPitch * (2 * p + 1)));
      }
#pragma unroll
      for (int p = 0; p < 4; ++p) x[p] = kCos ? __ffma2_rn(x[p], ov[p], bv[p]) : __fadd2_rn(x[p], bv[p]);
    } else if (kCos) {
#pragma unroll
      for (int p = 0; p < 4; ++p) x[p] = __fmul2_rn(x[p], ov[p]);
    }
    if (R.masked) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        uint32_t g4;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(g4) : "r"(R.groups + c0 + 4 * q) : "memory");
        if ((int)(g4 & 0xff) != R.my_group) x[2 * q].x += kMaskFill * kLog2e;
        if ((int)((g4 >> 8) & 0xff) != R.my_group) x[2 * q].y += kMaskFill * kLog2e;
        if ((int)((g4 >> 16) & 0xff) != R.my_group) x[2 * q + 1].x += kMaskFill * kLog2e;
        if ((int)((g4 >> 24) & 0xff) != R.my_group) x[2 * q + 1].y += kMaskFill * kLog2e;
      }
    }
    float4 acc[2];
    if (kNat != dbt_on_key_rows(kCos) && R.dbt_row) {  // issue the tile loads early; the row is owned by this thread alone (plain read-modify-write)
#pragma unroll
      for (int q = 0; q < 2; ++q) acc[q] = lds_f4(R.dbt_row + 16 * (((c0 >> 2) + q) ^ R.dbt_xor));
    }
    uint32_t pa[kCW], ua[kCW];
    float2 ds[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const float2 arg = __fadd2_rn(x[p], nl[p]);
      const float2 pv = f2(ex2_approx(arg.x), ex2_approx(arg.y));
      float2 dpe = f2u(dpr[2 * p], dpr[2 * p + 1]);
      float2 pd = pv;
      if (kDrop) {  // O = dropout(P) V:  dP -> dP o m,  the P fed to dV is P o m   (m = 0 or 1 / (1 - p))
        const int c = c0 + 2 * p;
        const bool k0 = kNat ? hs::drop_keep(R.drop_key, R.drop_r, c, kWS, R.drop_thresh)
                             : hs::drop_keep(R.drop_key, c, R.drop_r, kWS, R.drop_thresh);
        const bool k1 = kNat ? hs::drop_keep(R.drop_key, R.drop_r, c + 1, kWS, R.drop_thresh)
                             : hs::drop_keep(R.drop_key, c + 1, R.drop_r, kWS, R.drop_thresh);
        const float2 mk = f2(k0 ? R.drop_scale : 0.f, k1 ? R.drop_scale : 0.f);
        dpe = __fmul2_rn(dpe, mk);
        pd = __fmul2_rn(pd, mk);
      }
      dpe = __ffma2_rn(dpe, fx2, nd[p]);
      ds[p] = __fmul2_rn(pv, dpe);
      const float2 u = kCos ? __fmul2_rn(ds[p], ov[p]) : ds[p];
      if (kCos) {
        rs[p & 1] = __ffma2_rn(u, raw[p], rs[p & 1]);
        if (kNat) {
          sds = __fadd2_rn(sds, ds[p]);
          pw = __ffma2_rn(__fmul2_rn(pv, ov[p]), raw[p], pw);
        }
      }
      if (!kNat) {
        pa[2 * p] = __float_as_uint(tf32_rna(pd.x));
        pa[2 * p + 1] = __float_as_uint(tf32_rna(pd.y));
      }
      ua[2 * p] = __float_as_uint(tf32_rna(u.x));
      ua[2 * p + 1] = __float_as_uint(tf32_rna(u.y));
    }
    // the chunk of S / dP at these columns has been consumed: overwrite in place
    if (!kNat) tmem_st8(R.p_dst + c0, pa);
    tmem_st8(R.ds_dst + c0, ua);
    if (kNat != dbt_on_key_rows(kCos) && R.dbt_row) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float2 lo = __fadd2_rn(f2(acc[q].x, acc[q].y), ds[2 * q]), hi = __fadd2_rn(f2(acc[q].z, acc[q].w), ds[2 * q + 1]);
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(R.dbt_row + 16 * (((c0 >> 2) + q) ^ R.dbt_xor)),
                     "f"(lo.x), "f"(lo.y), "f"(hi.x), "f"(hi.y)
                     : "memory");
      }
    }
  };
  uint32_t sa[kCW], da[kCW], sb[kCW], db[kCW];
  tmem_ld8(R.s_src + cb, sa);
  tmem_ld8(R.dp_src + cb, da);
#pragma unroll 1
  for (int c0 = cb; c0 < cb + kSweepCols; c0 += 2 * kCW) {  // my warpgroup's share of the columns
    tmem_wait_ld();
    tmem_ld8(R.s_src + c0 + kCW, sb);
    tmem_ld8(R.dp_src + c0 + kCW, db);
    chunk(sa, da, c0);
    tmem_wait_ld();
    if (c0 + 2 * kCW < cb + kSweepCols) {
      tmem_ld8(R.s_src + c0 + 2 * kCW, sa);
      tmem_ld8(R.dp_src + c0 + 2 * kCW, da);
    }
    chunk(sb, db, c0 + kCW);
  }
  tmem_wait_st();
  // partial sums over my columns: (sum dS w, sum dS, sum P w); the epilogue adds the two halves of the row and centres
  return make_float4((rs[0].x + rs[0].y) + (rs[1].x + rs[1].y), sds.x + sds.y, pw.x + pw.y, 0.f);
}

template <bool kDrop, bool kCos, bool kBias>
__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap map_qkv_k, const __grid_constant__ CUtensorMap map_qkv_mn,
                   const __grid_constant__ CUtensorMap map_do_k, const __grid_constant__ CUtensorMap map_do_mn,
                   const __grid_constant__ CUtensorMap map_dqkv, const __grid_constant__ CUtensorMap map_out,
                   const BwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y;
  constexpr bool has_bias = kBias;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kKS; ++i) {
      mbar_init(&S.kfull[i], 2);
      mbar_init(&S.kempty[i], 128 * kSweepers);  // the unit's elementwise threads: scores done, statistics done
    }
    for (int i = 0; i < kMS; ++i) {
      mbar_init(&S.full[i], 2);
      mbar_init(&S.empty[i], 128);  // the epilogue threads
      mbar_init(&S.meta_ready[i], 1);
      mbar_init(&S.stats_ready[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&S.s_ready[i], 1);
      mbar_init(&S.dsn_ready[i], 64 * kSweepers);  // dS of the 64 query rows is in TMEM   -> dQ
      mbar_init(&S.dst_ready[i], 64 * kSweepers);  // P^T, dS^T of the 64 key rows           -> dV, dK
      mbar_init(&S.o_ready[i], 1);
      mbar_init(&S.stage_free[i], 128);  // the epilogue threads, once the outputs are in their registers
    }
    mbar_fence_init();
  }
  if (warp == 9) tmem_alloc(&S.tmem_base, kTmemCols);
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&map_qkv_k);
    tma_prefetch_desc(&map_qkv_mn);
    tma_prefetch_desc(&map_do_k);
    tma_prefetch_desc(&map_do_mn);
    tma_prefetch_desc(&map_dqkv);
    tma_prefetch_desc(&map_out);
  }
  for (int idx = threadIdx.x; idx < 2 * kWS * kDbtPitch; idx += kThreads) (&S.dbt[0][0])[idx] = 0.f;
  if (has_bias) {
    const float* bp = a.bias + (long long)h * kWS * kWS;
    for (int idx = threadIdx.x; idx < kWS * kWS; idx += kThreads) {
      const int i = idx >> 6, j = idx & 63;
      const float v = __ldg(bp + idx) * kLog2e;
      S.bias[i * kBiasPitch + j] = v;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;

  // register file: 2 x 128 x 152 (elementwise) + 128 x 104 (producer / MMA / statistics) + 128 x 104 (epilogue) = 64 K
  if (warp >= 12) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
    // ================================================================= epilogue warps: one TMEM lane quadrant each (warps 12, 13:
    // query rows -> dQ; warps 14, 15: key rows -> dK, dV), every unit of both TMEM stages.  The elementwise warpgroups
    // never wait for the output MMAs: they hand the row sums over through S.rs and go on with their next unit.
    const int q4 = warp - 12;
    const int L = q4 * 32 + lane;  // TMEM lane
    const bool lower = L < kWS;
    const int r = L & 63;
    const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
    const float eff = kCos ? __expf(fminf(__ldg(a.logit_scale + h), kLogitScaleMax)) : a.scale;
    bool store_pending = false;
    float racc = 0.f;  // running sum of dS o (logits without bias / mask): d logit_scale
    int n = 0;
    for (int unit = blockIdx.x; unit < a.total; unit += gridDim.x, ++n) {
      const int slot = n % kMS, t = n & 1;
      const uint32_t ph = (uint32_t)(n >> 1) & 1;
      const bool mirror = kMirror && t == 1;
      const uint32_t sx = mirror ? 64u : 0u;  // XOR of the TMEM column offsets
      const bool nat = lower != mirror;       // query rows (-> dQ) or key rows (-> dK, dV)
      // named barrier of the 64 threads of this half: by LANE half, not by role -- with the mirror the roles alternate per
      // unit, and a barrier id per role would pair a warp of one unit with a warp of the next
      const int half_bar = lower ? 5 : 6;
      [[maybe_unused]] const int trole = nat ? 6 : 7;  // trace role (diagnostics build only)
      if (r == 0) HS_TRACE(trole, n, 0);
      mbar_wait(&S.full[slot], (uint32_t)(n / kMS) & 1);  // (long complete: makes the row table visible)
      const SlotMeta& M = S.meta[slot];
      Slot& T = S.slot[slot];
      const int flags = M.flags;
      const int my_row = M.rows[r];
      const uint32_t D1 = tmem + (uint32_t)t * kStageCols + lane_addr, D2 = D1 + 128;
      mbar_wait(nat ? &S.dsn_ready[t] : &S.dst_ready[t], ph);  // my half's sweep is done: S.rs is valid
      const float my_inv = S.inv[slot][nat ? r : kWS + r];
      float rs = 0.f;  // sum_c dS[r][c] * (eff * cos(r, c))   (cos attention only)
      if (kCos) {
        const float4 p0 = S.part[t][0][L], p1 = kCoop ? S.part[t][1][L] : make_float4(0.f, 0.f, 0.f, 0.f);
        // (the centring applies along a query row only: the column sums of dS seen by the key rows do not vanish)
        const float centre = nat ? (p0.y + p1.y) * (p0.z + p1.z) : 0.f;
        rs = ((p0.x + p1.x) - centre) * (eff * my_inv * a.fix2);
        if (nat) racc += rs;
      }
      mbar_wait(&S.o_ready[t], ph);
      if (r == 0) HS_TRACE(trole, n, 1);
      tc_fence_after();
      uint32_t acc0[kD], acc1[kD];
      if (nat) {
        tmem_ld32(D2 + (64u ^ sx), acc0);  // dQ
      } else {
        tmem_ld32(D1 + (64u ^ sx), acc0);  // dK
        tmem_ld32(D2 + (96u ^ sx), acc1);  // dV
      }
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(&S.stage_free[t]);
      if (r == 0) HS_TRACE(trole, n, 2);

      // through the scaling / F.normalize:  d row = g * acc - row * corr
      const float g = eff * my_inv * a.fix1;  // dS (rounded) x k / q (truncated)
      const bool clamped = my_inv >= 1.0f / kNormEps;
      const float corr = (kCos && !clamped) ? my_inv * my_inv * rs : 0.f;
      const int row0 = M.rows[0];
      const bool contig = !kDirectStore && (flags & kFlagContig) != 0;
      uint8_t* st0 = (nat ? T.q_mn : T.k_mn) + r * 128;
      uint8_t* st1 = T.do_mn + r * 128;
      float* g0 = a.dqkv + (long long)my_row * 3 * a.C + (nat ? 0 : a.C) + h * kD;
      const float2 g2 = f2(g, g), nc2 = f2(-corr, -corr), f12 = f2(a.fix1, a.fix1);
      if (kCos) {
        // my own row of q / k comes from the MN-major copy (the K-major tiles are long recycled); the staged output row
        // goes to the same 128 bytes in another chunk order: read the whole row before the first write
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 x4 = *reinterpret_cast<const float4*>(st0 - r * 128 + sw128b32_off(r, c));
          const float2 lo = __ffma2_rn(f2(x4.x, x4.y), nc2, __fmul2_rn(f2u(acc0[4 * c + 0], acc0[4 * c + 1]), g2));
          const float2 hi = __ffma2_rn(f2(x4.z, x4.w), nc2, __fmul2_rn(f2u(acc0[4 * c + 2], acc0[4 * c + 3]), g2));
          acc0[4 * c + 0] = __float_as_uint(lo.x); acc0[4 * c + 1] = __float_as_uint(lo.y);
          acc0[4 * c + 2] = __float_as_uint(hi.x); acc0[4 * c + 3] = __float_as_uint(hi.y);
        }
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float2 lo = __fmul2_rn(f2u(acc0[4 * c + 0], acc0[4 * c + 1]), g2);
          const float2 hi = __fmul2_rn(f2u(acc0[4 * c + 2], acc0[4 * c + 3]), g2);
          acc0[4 * c + 0] = __float_as_uint(lo.x); acc0[4 * c + 1] = __float_as_uint(lo.y);
          acc0[4 * c + 2] = __float_as_uint(hi.x); acc0[4 * c + 3] = __float_as_uint(hi.y);
        }
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 v4 = make_float4(__uint_as_float(acc0[4 * c + 0]), __uint_as_float(acc0[4 * c + 1]),
                                      __uint_as_float(acc0[4 * c + 2]), __uint_as_float(acc0[4 * c + 3]));
        if (contig)
          *reinterpret_cast<float4*>(st0 + ((c ^ (r & 7)) << 4)) = v4;
        else
          reinterpret_cast<float4*>(g0)[c] = v4;
      }
      if (!nat) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float2 lo = __fmul2_rn(f2u(acc1[4 * c + 0], acc1[4 * c + 1]), f12);  // P^T (rounded) x dO (truncated)
          const float2 hi = __fmul2_rn(f2u(acc1[4 * c + 2], acc1[4 * c + 3]), f12);
          const float4 v4 = make_float4(lo.x, lo.y, hi.x, hi.y);
          if (contig)
            *reinterpret_cast<float4*>(st1 + ((c ^ (r & 7)) << 4)) = v4;
          else
            reinterpret_cast<float4*>(g0 + a.C)[c] = v4;
        }
      }
      if (contig) {
        fence_proxy_async_smem();
        named_bar_sync(half_bar, 64);
        if (r == 0) {
          if (nat) {
            tma_store_2d(&map_dqkv, T.q_mn, h * kD, row0);
          } else {
            tma_store_2d(&map_dqkv, T.k_mn, a.C + h * kD, row0);
            tma_store_2d(&map_dqkv, T.do_mn, 2 * a.C + h * kD, row0);
          }
          tma_store_commit();
          store_pending = true;
          if (r == 0) HS_TRACE(trole, n, 3);
          tma_store_wait_read<0>();  // the staging tiles live in the slot: it is free once the TMA engine has read them
        }
      }
      mbar_arrive(&S.empty[slot]);
      if (r == 0) HS_TRACE(trole, n, 4);
    }
    if (store_pending) tma_store_wait<0>();
    if (kCos && a.dlogit) {  // (every epilogue thread owned query rows in some units)
      // d logit_scale = sum dS o logits_cos, zero where the clamp is active (torch.clamp backward)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) racc += __shfl_xor_sync(0xffffffffu, racc, o);
      if (lane == 0 && __ldg(a.logit_scale + h) <= kLogitScaleMax) atomicAdd(a.dlogit + h, racc);
    }
  } else if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
    if (warp == 8) {
      // ================================================================= load producer
      // The row table of a unit (two global loads per lane) is fetched one unit ahead, and as soon as it is known the
      // unit's tiles are prefetched into L2: the TMA loads proper are issued only when ring entries free up, two units
      // ahead of their use, which alone does not cover the HBM latency under load (profiles/r2z_attn_bwd_*).
      auto fetch_rows = [&](int u, int& r0, int& r1, int& g0, int& g1) {
        const int w = u % a.nW;
        const long long s0 = (long long)w * kWS;
        g0 = g1 = 0;
        if (a.src) {
          r0 = a.src[s0 + lane];
          r1 = a.src[s0 + 32 + lane];
        } else {
          r0 = (int)s0 + lane;
          r1 = r0 + 32;
        }
        if (a.groups) {
          g0 = a.groups[s0 + lane];
          g1 = a.groups[s0 + 32 + lane];
        }
      };
      int nr0 = 0, nr1 = 0, ng0 = 0, ng1 = 0;
      if ((int)blockIdx.x < a.total) fetch_rows(blockIdx.x, nr0, nr1, ng0, ng1);
      int n = 0;
      for (int unit = blockIdx.x; unit < a.total; unit += gridDim.x, ++n) {
        const int slot = n % kMS, ks = n % kKS;
        const int r0 = nr0, r1 = nr1, g0 = ng0, g1 = ng1;
        const int unit_next = unit + (int)gridDim.x;
        if (unit_next < a.total) {
          fetch_rows(unit_next, nr0, nr1, ng0, ng1);
          const int nb = __shfl_sync(0xffffffffu, nr0, 0);
          if (kL2Prefetch && __all_sync(0xffffffffu, (nr0 == nb + lane) && (nr1 == nb + 32 + lane)) && elect_one()) {
            const int row = (int)((long long)(unit_next / a.nW) * a.N) + nb;
            tma_prefetch_l2_2d(&map_qkv_k, h * kD, row);
            tma_prefetch_l2_2d(&map_qkv_k, a.C + h * kD, row);
            tma_prefetch_l2_2d(&map_qkv_k, 2 * a.C + h * kD, row);
            tma_prefetch_l2_2d(&map_do_k, h * kD, row);
            tma_prefetch_l2_2d(&map_out, h * kD, row);
          }
        }
        mbar_wait(&S.kempty[ks], ((uint32_t)(n / kKS) & 1) ^ 1);
        mbar_wait(&S.empty[slot], ((uint32_t)(n / kMS) & 1) ^ 1);
        if (lane == 0) HS_TRACE(5, n, 0);
        SlotMeta& M = S.meta[slot];
        Slot& T = S.slot[slot];
        KSlot& TK = S.kslot[ks];
        const int qoff = (kMirror && (n & 1)) ? kTile : 0, koff = kTile - qoff;  // [Q;K], [dO;V] or mirrored [K;Q], [V;dO]
        const int b = unit / a.nW;
        const int rbase = __shfl_sync(0xffffffffu, r0, 0);
        const int gbase = __shfl_sync(0xffffffffu, g0, 0);
        const bool contig = __all_sync(0xffffffffu, (r0 == rbase + lane) && (r1 == rbase + 32 + lane));
        const bool un = __all_sync(0xffffffffu, (g0 == gbase) && (g1 == gbase));
        const int goff = (int)((long long)b * a.N);
        M.rows[lane] = goff + r0;
        M.rows[lane + 32] = goff + r1;
        M.groups[lane] = (uint8_t)g0;
        M.groups[lane + 32] = (uint8_t)g1;
        if (lane == 0) M.flags = kFlagValid | (contig ? kFlagContig : 0) | (un ? kFlagUniform : 0);
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.meta_ready[slot]);  // the statistics warps can start (they read rows[] only)
        if (elect_one()) {
          mbar_arrive_expect_tx(&S.kfull[ks], contig ? 4u * kTile : 0u);
          mbar_arrive_expect_tx(&S.full[slot], contig ? 3u * kTile : 0u);
          if (contig) {
            const int row = goff + rbase;
            tma_load_2d(TK.qk + qoff, &map_qkv_k, &S.kfull[ks], h * kD, row);
            tma_load_2d(TK.qk + koff, &map_qkv_k, &S.kfull[ks], a.C + h * kD, row);
            tma_load_2d(TK.dov + qoff, &map_do_k, &S.kfull[ks], h * kD, row);
            tma_load_2d(TK.dov + koff, &map_qkv_k, &S.kfull[ks], 2 * a.C + h * kD, row);
            tma_load_2d(T.q_mn, &map_qkv_mn, &S.full[slot], h * kD, row);
            tma_load_2d(T.k_mn, &map_qkv_mn, &S.full[slot], a.C + h * kD, row);
            tma_load_2d(T.do_mn, &map_do_mn, &S.full[slot], h * kD, row);
          }
        }
        if (!contig) {
          // shifted window whose rows are not consecutive: 16 B cp.async gathers into the same swizzled layouts
          const int c16 = lane & 7;
#pragma unroll 2
          for (int it = 0; it < 16; ++it) {
            const int r = it * 4 + (lane >> 3);
            const long long grow = M.rows[r];
            const float* g = a.qkv + grow * 3 * a.C + h * kD + c16 * 4;
            const float* gd = a.dout + grow * a.C + h * kD + c16 * 4;
            const uint32_t ok = sw128_off(r, c16), om = sw128b32_off(r, c16);
            cp_async16(TK.qk + qoff + ok, g);
            cp_async16(TK.qk + koff + ok, g + a.C);
            cp_async16(TK.dov + qoff + ok, gd);
            cp_async16(TK.dov + koff + ok, g + 2 * a.C);
            cp_async16(T.q_mn + om, g);
            cp_async16(T.k_mn + om, g + a.C);
            cp_async16(T.do_mn + om, gd);
          }
          cp_async_wait_all();
          fence_proxy_async_smem();
        }
        __syncwarp();
        if (lane == 0) {  // arrival 2 of 2 (publishes the metadata too)
          mbar_arrive(&S.kfull[ks]);
          mbar_arrive(&S.full[slot]);
        }
        if (lane == 0) HS_TRACE(5, n, 1);
      }
    } else if (warp == 9 && elect_one()) {
      // ================================================================= MMA issuer (one elected thread: with elect.sync the
      // compiler keeps descriptors in uniform registers and emits back-to-back UTCHMMA)
      constexpr uint64_t kDescK = umma_smem_desc(16, 1024, kLayoutSw128);       // K-major, 8-row groups 1024 B apart
      constexpr uint64_t kDescMN = umma_smem_desc(1024, 512, kLayoutSw128B32);  // MN-major, 4-row k-atoms 512 B apart
      constexpr uint32_t kIdescS = umma_idesc_tf32(128, 128, 0, 0);
      constexpr uint32_t kIdescO = umma_idesc_tf32(128, 32, 0, 1);
      // One thread schedules the MMAs of both TMEM stages.  It POLLS the barriers (score MMAs of unit ns as soon as its
      // slot is full and its stage free; dQ of unit no once the query rows are done; dV, dK once the key rows are done),
      // so that the score MMAs of one warpgroup's next unit never queue behind the other warpgroup's unfinished sweep.
      int units = 0;
      for (int unit = blockIdx.x; unit < a.total; unit += gridDim.x) ++units;
      int ns = 0, no = 0, ophase = 0;
      long long idle0 = 0;
      while (no < units) {
        bool progressed = false;
        if (ns < units) {
          const int ks = ns % kKS, t = ns & 1;
          if (mbar_test_wait(&S.kfull[ks], (uint32_t)(ns / kKS) & 1) &&
              mbar_test_wait(&S.stage_free[t], ((uint32_t)(ns >> 1) & 1) ^ 1)) {
            HS_TRACE(4, ns, 0);
            tc_fence_after();
            const uint32_t D1 = tmem + (uint32_t)t * kStageCols, D2 = D1 + 128;
            const uint32_t qk = smem_u32(S.kslot[ks].qk), dov = smem_u32(S.kslot[ks].dov);
#pragma unroll
            for (int s = 0; s < 4; ++s)  // [Q;K] [Q;K]^T : lanes 0-63 x cols 64-127 = S, lanes 64-127 x cols 0-63 = S^T
              umma_tf32_ss(D1, umma_desc_at(kDescK, qk + s * 32), umma_desc_at(kDescK, qk + s * 32), kIdescS, s > 0);
#pragma unroll
            for (int s = 0; s < 4; ++s)  // [dO;V] [dO;V]^T : lanes 0-63 x cols 64-127 = dP, lanes 64-127 x cols 0-63 = dP^T
              umma_tf32_ss(D2, umma_desc_at(kDescK, dov + s * 32), umma_desc_at(kDescK, dov + s * 32), kIdescS, s > 0);
            umma_commit(&S.s_ready[t]);
            HS_TRACE(4, ns, 1);
            ++ns;
            progressed = true;
          }
        }
        if (no < ns) {
          const int t = no & 1, slot = no % kMS;
          const uint32_t ph = (uint32_t)(no >> 1) & 1;
          const uint32_t D1 = tmem + (uint32_t)t * kStageCols, D2 = D1 + 128;
          const Slot& T = S.slot[slot];
          const uint32_t sx = (kMirror && t == 1) ? 64u : 0u;  // mirrored stage: every column offset XOR 64
          if (ophase == 0 && mbar_test_wait(&S.dsn_ready[t], ph) && mbar_test_wait(&S.full[slot], (uint32_t)(no / kMS) & 1)) {
            HS_TRACE(4, no, 2);
            tc_fence_after();
            const uint32_t kb = smem_u32(T.k_mn);
#pragma unroll
            for (int s = 0; s < 8; ++s)  // dQ = dS k            A: D1[:, 64:128)  ->  D2[:, 64:96)
              umma_tf32_ts(D2 + (64u ^ sx), D1 + (64u ^ sx) + s * 8, umma_desc_at(kDescMN, kb + s * 1024), kIdescO, s > 0);
            ophase = 1;
            progressed = true;
          }
          if (ophase == 1 && mbar_test_wait(&S.dst_ready[t], ph)) {
            tc_fence_after();
            const uint32_t db = smem_u32(T.do_mn), qb = smem_u32(T.q_mn);
#pragma unroll
            for (int s = 0; s < 8; ++s)  // dV = P^T dO          A: D1[:, 0:64)    ->  D2[:, 96:128)
              umma_tf32_ts(D2 + (96u ^ sx), D1 + sx + s * 8, umma_desc_at(kDescMN, db + s * 1024), kIdescO, s > 0);
            HS_TRACE(4, no, 3);
            // dK overwrites the dS block that dQ has read: tcgen05.mma instructions of one thread execute in issue order
#pragma unroll
            for (int s = 0; s < 8; ++s)  // dK = dS^T q          A: D2[:, 0:64)    ->  D1[:, 64:96)
              umma_tf32_ts(D1 + (64u ^ sx), D2 + sx + s * 8, umma_desc_at(kDescMN, qb + s * 1024), kIdescO, s > 0);
            umma_commit(&S.o_ready[t]);
            HS_TRACE(4, no, 4);
            ophase = 0;
            ++no;
            progressed = true;
          }
        }
        if (progressed) {
          idle0 = 0;
        } else {  // nothing ready: a protocol bug must not hang the device (same ~2 s bound as mbar_wait)
          if (idle0 == 0) idle0 = clock64();
          else if (clock64() - idle0 > 4000000000ll) __trap();
          __nanosleep(20);
        }
      }
    } else if (warp >= 10) {
      // ================================================================= statistics warps (10: even units, 11: odd units).
      // The softmax row statistics come from the forward pass -- lse from its saved vector, delta_i = sum_j P_ij dP_ij
      // = dO_i . O_i from its output -- so the elementwise warpgroups need no statistics sweep over S.  They start as
      // soon as the producer has published the unit's row table, i.e. in parallel with the TMA loads of the slot.
      int n = 0;
      for (int unit = blockIdx.x; unit < a.total; unit += gridDim.x, ++n) {
        if ((n & 1) != (warp & 1)) continue;
        const int slot = n % kMS, ks = n % kKS;
        mbar_wait(&S.meta_ready[slot], (uint32_t)(n / kMS) & 1);
        const SlotMeta& M = S.meta[slot];
        float4 o[2][8];
        float lse[2], qi[2] = {1.f, 1.f}, ki[2] = {1.f, 1.f};
        const long long plane = (long long)a.H * ((long long)a.B * a.N);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const long long row = M.rows[lane + 32 * k];
          const float4* orow = reinterpret_cast<const float4*>(a.out + row * a.C + h * kD);
#pragma unroll
          for (int c = 0; c < 8; ++c) o[k][c] = __ldg(orow + c);
          lse[k] = __ldg(a.lse + (long long)h * ((long long)a.B * a.N) + row);
          if (kCos) {
            qi[k] = __ldg(a.lse + plane + (long long)h * ((long long)a.B * a.N) + row);
            ki[k] = __ldg(a.lse + 2 * plane + (long long)h * ((long long)a.B * a.N) + row);
          }
        }
        // dO comes from the slot's K-major tile (rows 0-63 of [dO;V]) once the loads have landed
        mbar_wait(&S.kfull[ks], (uint32_t)(n / kKS) & 1);
        const uint8_t* dtile = S.kslot[ks].dov + ((kMirror && (n & 1)) ? kTile : 0);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int rr = lane + 32 * k;
          float2 dot = f2(0.f, 0.f);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 d4 = *reinterpret_cast<const float4*>(dtile + sw128_off(rr, c));
            dot = __ffma2_rn(f2(o[k][c].x, o[k][c].y), f2(d4.x, d4.y), dot);
            dot = __ffma2_rn(f2(o[k][c].z, o[k][c].w), f2(d4.z, d4.w), dot);
          }
          S.nlse[slot][rr] = -lse[k];
          S.ndelta[slot][rr] = -(dot.x + dot.y);
          S.inv[slot][rr] = qi[k];
          S.inv[slot][kWS + rr] = ki[k];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.stats_ready[slot]);
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    // ================================================================= elementwise warpgroups
    const int wg = warp >> 2;              // see kCoop
    const int L = (warp & 3) * 32 + lane;  // TMEM lane
    const bool lower = L < kWS;
    const bool nat = lower != (kMirror && wg == 1);  // query rows or key rows (mirrored on stage 1 = warpgroup 1)
    const int r = L & 63;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const float eff = kCos ? __expf(fminf(__ldg(a.logit_scale + h), kLogitScaleMax)) : a.scale;
    int n = 0;
    for (int unit = blockIdx.x; unit < a.total; unit += gridDim.x, ++n) {
      if (!kCoop && (n & 1) != wg) continue;
      const int slot = n % kMS, ks = n % kKS, t = n & 1;
      const uint32_t it = (uint32_t)(n >> 1) & 1;
      const uint32_t D1 = tmem + (uint32_t)t * kStageCols + lane_addr, D2 = D1 + 128;
      mbar_wait(&S.kfull[ks], (uint32_t)(n / kKS) & 1);  // (makes the unit's metadata visible as well)
      [[maybe_unused]] const int trole = wg * 2 + (nat ? 0 : 1);  // trace role (diagnostics build only)
      if (r == 0) HS_TRACE(trole, n, 0);

      const SlotMeta& M = S.meta[slot];
      const int flags = M.flags;
      if (r == 0) HS_TRACE(trole, n, 6);
      mbar_wait(&S.stats_ready[slot], (uint32_t)(n / kMS) & 1);  // lse / delta / norms of this unit (statistics warps)
      if (r == 0) HS_TRACE(trole, n, 7);
      const float my_inv = S.inv[slot][nat ? r : kWS + r];  // 1/|q_r| (query rows) or 1/|k_r| (key rows); 1 without cos attention
      const float row_scale = eff * kLog2e * my_inv * a.fix2;  // S = q k^T has two truncated operands

      mbar_wait(&S.s_ready[t], it);
      mbar_arrive(&S.kempty[ks]);  // score MMAs and statistics are done with the K-major tiles
      if (r == 0) HS_TRACE(trole, n, 1);
      tc_fence_after();
      RowCtx R;
      R.s_src = D1 + (lower ? 64u : 0u);   // S (query rows) / S^T (key rows): lanes 0-63 x cols 64-127, lanes 64-127 x cols 0-63
      R.dp_src = D2 + (lower ? 64u : 0u);  // dP / dP^T
      R.p_dst = R.s_src;                   // P^T over S^T (key rows)
      R.ds_dst = nat ? R.s_src : R.dp_src; // dS over S / dS^T over dP^T
      R.oinv = smem_u32(S.inv[slot] + (nat ? kWS : 0));
      R.brow = smem_u32(S.bias + (nat ? r * kBiasPitch : r));
      R.nlse_v = smem_u32(S.nlse[slot]);
      R.ndelta_v = smem_u32(S.ndelta[slot]);
      R.nlse = S.nlse[slot][r];
      R.ndelta = S.ndelta[slot][r];
      R.groups = smem_u32(M.groups);
      R.dbt_row = (nat != dbt_on_key_rows(kCos) && a.dbias) ? smem_u32(S.dbt[wg] + r * kDbtPitch) : 0u;
      R.dbt_xor = r & 15;
      R.row_scale = row_scale;
      R.fix2 = a.fix2;
      R.my_group = M.groups[r];
      R.masked = !(flags & kFlagUniform);
      R.drop_thresh = a.drop_thresh;
      R.drop_scale = a.drop_scale;
      R.drop_key = a.drop_thresh ? hs::drop_unit_key(a.seed, unit, h, a.H) : 0u;
      R.drop_r = r;
      const int cb = kCoop ? 32 * wg : 0;
      const float4 part = nat ? ds_sweep<true, kCos, kBias, kDrop>(R, cb) : ds_sweep<false, kCos, kBias, kDrop>(R, cb);
      if (kCos) S.part[t][kCoop ? wg : 0][L] = part;  // for the epilogue warps (ordered by the arrive below)
      tc_fence_before();
      mbar_arrive(nat ? &S.dsn_ready[t] : &S.dst_ready[t]);
      if (r == 0) HS_TRACE(trole, n, 2);
    }

    if (a.dbias) {
      // both warpgroups accumulated into the same tile: wait for all 8 elementwise warps, then one atomic per entry
      named_bar_sync(9, kEwThreads);
      float* gb = a.dbias + (long long)h * kWS * kWS;
      for (int idx = threadIdx.x; idx < kWS * kWS; idx += kEwThreads) {
        const int i = idx >> 6, j = idx & 63;
        const int tr = dbt_on_key_rows(kCos) ? j : i, tc = dbt_on_key_rows(kCos) ? i : j;  // tile row (owner thread) / column
        const int pos = tr * kDbtPitch + 4 * ((tc >> 2) ^ (tr & 15)) + (tc & 3);
        atomicAdd(gb + idx, S.dbt[0][pos] + S.dbt[1][pos]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, kTmemCols);
}

}  // namespace

namespace hs {

int window_attn_bwd_tc(const float* qkv, const float* out, const float* lse, const float* dout, const int32_t* src,
                       const uint8_t* groups, const float* bias, const float* logit_scale, float scale, DropCfg drop,
                       float* dqkv, float* dbias, float* dlogit, int B, int64_t N, int C, int H, uint32_t flags,
                       cudaStream_t stream) {
  HS_REQUIRE(qkv && dout && dqkv && out && lse, "hs_window_attn_bwd: null qkv/out/lse/dout/dqkv");
  HS_REQUIRE(!(flags & HS_ATTN_COS) || logit_scale, "hs_window_attn_bwd: cos attention needs logit_scale");
  CUtensorMap map_qkv_k, map_qkv_mn, map_do_k, map_do_mn, map_dqkv, map_out;
  const long long rows = (long long)B * N;
  int rc;
  if ((rc = make_map(&map_qkv_k, qkv, rows, 3 * C, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_map(&map_qkv_mn, qkv, rows, 3 * C, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
  if ((rc = make_map(&map_do_k, dout, rows, C, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_map(&map_do_mn, dout, rows, C, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
  if ((rc = make_map(&map_dqkv, dqkv, rows, 3 * C, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_map(&map_out, out, rows, C, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;  // (L2 prefetch only)
  BwdArgs a{};
  a.qkv = qkv; a.out = out; a.lse = lse; a.dout = dout; a.dqkv = dqkv; a.src = src; a.groups = groups; a.bias = bias;
  a.logit_scale = logit_scale; a.dbias = dbias; a.dlogit = dlogit; a.scale = scale;
  a.B = B; a.nW = (int)(N / kWS); a.C = C; a.H = H; a.cos = (flags & HS_ATTN_COS) ? 1 : 0; a.N = N;
  a.total = B * a.nW;
  a.fix1 = (flags & HS_ATTN_NO_TRUNC_COMP) ? 1.0f : kTruncFix1;
  a.fix2 = (flags & HS_ATTN_NO_TRUNC_COMP) ? 1.0f : kTruncFix2;
  a.drop_thresh = drop.p > 0.f ? hs::drop_thresh(drop.p) : 0u;
  a.drop_scale = 1.0f / (1.0f - drop.p);
  a.seed = drop.seed;
  const size_t smem = sizeof(Smem) + 1024;
  int gx = sm_count() / H;
  if (gx < 1) gx = 1;
  if (gx > a.total) gx = a.total;
  const dim3 grid(gx, H);
  // one instantiation per (dropout, cos, bias): the variant flags are compile-time constants of the elementwise sweep
  using Kernel = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                          const CUtensorMap, const BwdArgs);
  static const Kernel table[8] = {
      attn_bwd_tc_kernel<false, false, false>, attn_bwd_tc_kernel<false, false, true>,
      attn_bwd_tc_kernel<false, true, false>,  attn_bwd_tc_kernel<false, true, true>,
      attn_bwd_tc_kernel<true, false, false>,  attn_bwd_tc_kernel<true, false, true>,
      attn_bwd_tc_kernel<true, true, false>,   attn_bwd_tc_kernel<true, true, true>};
  static bool attr_done = false;  // benign race: the attribute is idempotent (and per function, not per device)
  if (!attr_done) {
    for (Kernel k : table) HS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  const Kernel kern = table[(a.drop_thresh ? 4 : 0) + (a.cos ? 2 : 0) + (bias ? 1 : 0)];
  kern<<<grid, kThreads, smem, stream>>>(map_qkv_k, map_qkv_mn, map_do_k, map_do_mn, map_dqkv, map_out, a);
  HS_LAUNCH_CHECK();
  return HS_OK;
}

}  // namespace hs
