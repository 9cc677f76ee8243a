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
// two swizzles (the second fetch hits L2).  HBM traffic per unit: q, k, v, dO in; dq, dk, dv out = 7 x 8 KB.
//
// DRAFT v2 (never run): 16 elementwise warps.  Warp roles (640 threads): warps 0-15 = four elementwise warpgroups g;
// unit n uses TMEM stage n & 1 and is swept by the TWO warpgroups with (g & 1) == (n & 1): warpgroup half g >> 1 takes
// columns [32 half, 32 half + 32) of every row (the tcgen05.ld lane restriction is per warp-in-warpgroup, so both reach
// all 128 lanes).  Row sums are combined through S.xrs / S.xsp.  Epilogue tiles: half 0 query rows -> dQ, half 0 key
// rows -> dK, half 1 key rows -> dV.  Warp 16 = load producer, warp 17 = MMA issuer, warps 18-19 = row statistics.
// 3 shared-memory slots of 56 KB; the output tiles are staged in the slot's (dead) MN-major tiles and written back by TMA.
#include <cfloat>

#include "hs_common.h"
#include "hs_kernels.h"
#include "hs_sm100.cuh"
#include "hs_tc_common.cuh"

namespace {

using namespace hs::sm100;
using namespace hs::tc;

constexpr int kSlots = 3;
constexpr int kStageCols = 256;  // D1 (128) + D2 (128)
constexpr int kTmemCols = 512;
constexpr int kThreads = 640;
constexpr int kEpiWarps = 16;
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

struct Slot {
  uint8_t qk[2 * kTile];   // [Q;K]  K-major, SWIZZLE_128B
  uint8_t dov[2 * kTile];  // [dO;V] K-major, SWIZZLE_128B
  uint8_t q_mn[kTile];     // MN-major (SWIZZLE_128B_ATOM_32B); reused as dQ staging
  uint8_t k_mn[kTile];     //   "                                reused as dK staging
  uint8_t do_mn[kTile];    //   "                                reused as dV staging
};

struct Smem {
  Slot slot[kSlots];
  float bias[kWS * kBiasPitch];     // bias[i][j] * log2(e); query rows read it row-wise (LDS.128), key rows column-wise
  SlotMeta meta[kSlots];
  float inv[kSlots][2 * kWS];  // per slot (statistics warps, from the forward): [0,64) 1/max(|q_i|,eps), [64,128) 1/max(|k_j|,eps)
  float lse[kSlots][kWS];    // per slot (written by the statistics warp): log2-domain log-sum-exp of every query row
  float delta[kSlots][kWS];  // per slot: rowsum(P o dP) = dO_i . O_i
  float xrs[2][2][2 * kWS];  // [stage][half][TMEM lane]: partial sum_c dS_c w_c over the half's 32 columns
  float2 xsp[2][2][kWS];     // [stage][half][query row]: partial (sum_c dS_c, sum_c P_c w_c) for the mean-centring
  float dbt[2][kWS * kDbtPitch];  // per stage: dbt[i][j] = sum over its units of dS[i][j] (row i, the half's 32 columns: one owner)
  uint64_t full[kSlots], empty[kSlots], meta_ready[kSlots], stats_ready[kSlots];
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
constexpr int kTraceUnits = 48, kTracePoints = 8, kTraceRoles = 6;  // roles: wg0 nat, wg0 tr, wg1 nat, wg1 tr, mma, producer
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
// Elementwise stage of one unit.  Written as ROLLED loops over 8-column chunks with everything recomputed from TMEM
// (no per-row register arrays): the whole hot path is a few hundred instructions and stays resident in the
// instruction cache -- the fully unrolled first version (4000 instructions, 64 KB) spent half of its issue slots
// waiting for instruction fetches (profiles/r1d_attn_bwd_tc_*).  TMEM loads are double-buffered: the next chunk is in
// flight while the current one is used.
struct RowCtx {
  uint32_t s_src, dp_src;  // my 64-column block of D1 (S or S^T) and D2 (dP or dP^T), lane field included
  uint32_t oinv, brow;     // shared-memory addresses: normalisation of the other index (cos), my bias row / column
  uint32_t bstep;          // byte step between consecutive bias entries along my columns (4: row, 4 * pitch: column)
  uint32_t groups;         // shared-memory address of the unit's 64 group ids
  float row_scale;         // log2(e) * scale (* my 1/|row| for cos) * truncation fix
  int my_group;
  bool cos, has_bias, masked;
  // attention dropout: element (i, j) of the unit; my row is index `drop_r`, columns run over the other index
  uint32_t drop_key, drop_thresh;  // thresh 0 = off
  float drop_scale;
  int drop_r;
  bool drop_row_is_query;  // query-row thread: (i, j) = (drop_r, c); key-row thread: (i, j) = (c, drop_r); also set without dropout
  float my_lse, my_delta;  // query-row threads: the statistics of my own row (key-row threads read vectors instead)
};

struct RowSums {
  float rs, sds, pw;  // partial sums over this thread's 32 columns
};

constexpr int kCW = 8;  // columns per chunk of the elementwise loops (16 was measured slower: 1.44 vs 1.27 ms)

__device__ __forceinline__ void tmem_ld_chunk(uint32_t taddr, uint32_t (&r)[kCW]) {
  if constexpr (kCW == 8) tmem_ld8(taddr, r); else tmem_ld16(taddr, r);
}
__device__ __forceinline__ void tmem_st_chunk(uint32_t taddr, const uint32_t (&r)[kCW]) {
  if constexpr (kCW == 8) tmem_st8(taddr, r); else tmem_st16(taddr, r);
}

// log2-domain logits of columns [c0, c0 + kCW) from the raw tensor-core products; ov = normalisation of the other index
__device__ __forceinline__ void logits_chunk(const RowCtx& R, const uint32_t (&raw)[kCW], int c0, float (&x)[kCW],
                                             float (&ov)[kCW]) {
  float bv[kCW];
#pragma unroll
  for (int q = 0; q < kCW / 4; ++q) {
    float4 o4 = make_float4(1.f, 1.f, 1.f, 1.f);
    if (R.cos) o4 = lds_f4(R.oinv + 4 * (c0 + 4 * q));
    ov[4 * q + 0] = o4.x; ov[4 * q + 1] = o4.y; ov[4 * q + 2] = o4.z; ov[4 * q + 3] = o4.w;
  }
  if (R.has_bias) {
    if (R.bstep == 4) {
#pragma unroll
      for (int q = 0; q < kCW / 4; ++q) {
        const float4 b4 = lds_f4(R.brow + 4 * (c0 + 4 * q));
        bv[4 * q + 0] = b4.x; bv[4 * q + 1] = b4.y; bv[4 * q + 2] = b4.z; bv[4 * q + 3] = b4.w;
      }
    } else {  // key rows: bias[c][r] walks down a column; consecutive lanes hit consecutive banks
#pragma unroll
      for (int e = 0; e < kCW; ++e) bv[e] = lds_f1(R.brow + R.bstep * (c0 + e));
    }
  } else {
#pragma unroll
    for (int e = 0; e < kCW; ++e) bv[e] = 0.f;
  }
#pragma unroll
  for (int e = 0; e < kCW; ++e) x[e] = fmaf(__uint_as_float(raw[e]) * R.row_scale, ov[e], bv[e]);
  if (R.masked) {
#pragma unroll
    for (int q = 0; q < kCW / 4; ++q) {
      uint32_t g4;
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(g4) : "r"(R.groups + c0 + 4 * q) : "memory");
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if ((int)((g4 >> (8 * e)) & 0xff) != R.my_group) x[4 * q + e] += kMaskFill * kLog2e;
    }
  }
}

// all threads: p = exp2(logit - lse), dS = p (dP - delta); P and dS (scaled by the other index' 1/norm for cos) go back
// to TMEM as TF32 A operands; query-row threads also accumulate dS into their warpgroup's dbias tile.
// lse_v / delta_v: shared addresses of the statistics seen along my columns, `vstep` = 1 (vectors over the query index,
// key-row threads) or 0 (my own row's 4-fold copy, query-row threads).  Returns sum_c dS_c * raw_c (for cos).
// dbt_row: shared address of my row of the dbias tile (0 = none); 16-byte chunk c4 of row r lives at chunk (c4 ^ (r & 15)).
template <bool kDrop>
__device__ __forceinline__ RowSums ds_sweep(const RowCtx& R, float fix2, uint32_t lse_v, uint32_t delta_v, uint32_t vstep,
                                            uint32_t p_dst, uint32_t ds_dst, uint32_t dbt_row, int dbt_xor, int cb) {
  // rs = sum_c dS_c * w_c with w = raw * (1/norm of the other index) (the cos logit up to my row's scale).  In exact
  // arithmetic sum_c dS_c = 0 along a query row, so any constant may be subtracted from w: the P-weighted mean of w is
  // subtracted (sds * pw) so that an error of the row's delta (it now comes from the forward output) is not amplified.
  float rs[4] = {0.f, 0.f, 0.f, 0.f}, sds[2] = {0.f, 0.f}, pw[2] = {0.f, 0.f};
  auto chunk = [&](const uint32_t (&sraw)[kCW], const uint32_t (&dpr)[kCW], int c0) {
    float x[kCW], ov[kCW], lv[kCW], dv[kCW], ds[kCW];
    logits_chunk(R, sraw, c0, x, ov);
    if (vstep) {  // key rows: the statistics of the query index run along my columns
#pragma unroll
      for (int q = 0; q < kCW / 4; ++q) {
        const float4 l4 = lds_f4(lse_v + 4 * (c0 + 4 * q));
        const float4 d4 = lds_f4(delta_v + 4 * (c0 + 4 * q));
        lv[4 * q + 0] = l4.x; lv[4 * q + 1] = l4.y; lv[4 * q + 2] = l4.z; lv[4 * q + 3] = l4.w;
        dv[4 * q + 0] = d4.x; dv[4 * q + 1] = d4.y; dv[4 * q + 2] = d4.z; dv[4 * q + 3] = d4.w;
      }
    } else {  // query rows: my own row's values
#pragma unroll
      for (int e = 0; e < kCW; ++e) {
        lv[e] = R.my_lse;
        dv[e] = R.my_delta;
      }
    }
    float4 acc[kCW / 4];
    if (dbt_row) {  // issue the tile loads early; the row is owned by this thread alone (plain read-modify-write)
#pragma unroll
      for (int q = 0; q < kCW / 4; ++q) acc[q] = lds_f4(dbt_row + 16 * (((c0 >> 2) + q) ^ dbt_xor));
    }
    uint32_t pa[kCW], ua[kCW];
#pragma unroll
    for (int e = 0; e < kCW; ++e) {
      const float pv = ex2_approx(x[e] - lv[e]);
      float dpe = __uint_as_float(dpr[e]) * fix2, pd = pv;
      if (kDrop) {  // O = dropout(P) V:  dP -> dP o m,  the P fed to dV is P o m   (m = 0 or 1 / (1 - p))
        const int c = c0 + e;
        const bool keep = R.drop_row_is_query ? hs::drop_keep(R.drop_key, R.drop_r, c, kWS, R.drop_thresh)
                                              : hs::drop_keep(R.drop_key, c, R.drop_r, kWS, R.drop_thresh);
        const float mk = keep ? R.drop_scale : 0.f;
        dpe *= mk;
        pd *= mk;
      }
      ds[e] = pv * (dpe - dv[e]);
      const float u = ds[e] * ov[e];
      rs[e & 3] = fmaf(u, __uint_as_float(sraw[e]), rs[e & 3]);
      if (R.cos) {
        sds[e & 1] += ds[e];
        pw[e & 1] = fmaf(pv * ov[e], __uint_as_float(sraw[e]), pw[e & 1]);
      }
      pa[e] = __float_as_uint(tf32_rna(pd));
      ua[e] = __float_as_uint(tf32_rna(u));
    }
    // the chunk of S / dP at these columns has been consumed: overwrite in place
    tmem_st_chunk(p_dst + c0, pa);
    tmem_st_chunk(ds_dst + c0, ua);
    if (dbt_row) {
#pragma unroll
      for (int q = 0; q < kCW / 4; ++q) {
        acc[q].x += ds[4 * q + 0]; acc[q].y += ds[4 * q + 1]; acc[q].z += ds[4 * q + 2]; acc[q].w += ds[4 * q + 3];
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dbt_row + 16 * (((c0 >> 2) + q) ^ dbt_xor)),
                     "f"(acc[q].x), "f"(acc[q].y), "f"(acc[q].z), "f"(acc[q].w)
                     : "memory");
      }
    }
  };
  // single-buffered TMEM loads: with four elementwise warps per scheduler the other warps hide the tcgen05.ld latency,
  // and the 16 registers of a second buffer are what the 104-register budget cannot afford
  uint32_t sa[kCW], da[kCW];
  const int ce = cb + kWS / 2;  // this warpgroup's half of the columns
#pragma unroll 1
  for (int c0 = cb; c0 < ce; c0 += kCW) {
    tmem_ld_chunk(R.s_src + c0, sa);
    tmem_ld_chunk(R.dp_src + c0, da);
    tmem_wait_ld();
    chunk(sa, da, c0);
  }
  tmem_wait_st();
  // the centring product (sum dS)(sum P w) needs the sums over all 64 columns: the caller combines the two halves
  RowSums out;
  out.rs = (rs[0] + rs[1]) + (rs[2] + rs[3]);
  out.sds = sds[0] + sds[1];
  out.pw = pw[0] + pw[1];
  return out;
}

template <bool kDrop>
__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap map_qkv_k, const __grid_constant__ CUtensorMap map_qkv_mn,
                   const __grid_constant__ CUtensorMap map_do_k, const __grid_constant__ CUtensorMap map_do_mn,
                   const __grid_constant__ CUtensorMap map_dqkv, const BwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y;
  const bool has_bias = a.bias != nullptr;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kSlots; ++i) {
      mbar_init(&S.full[i], 2);
      mbar_init(&S.empty[i], 256);
      mbar_init(&S.meta_ready[i], 1);
      mbar_init(&S.stats_ready[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&S.s_ready[i], 1);
      mbar_init(&S.dsn_ready[i], 128);  // dS of the 64 query rows (both column halves) is in TMEM   -> dQ
      mbar_init(&S.dst_ready[i], 128);  // P^T, dS^T of the 64 key rows (both halves)                -> dV, dK
      mbar_init(&S.o_ready[i], 1);
      mbar_init(&S.stage_free[i], 256);
    }
    mbar_fence_init();
  }
  if (warp == kEpiWarps + 1) tmem_alloc(&S.tmem_base, kTmemCols);
  if (warp == kEpiWarps && lane == 0) {
    tma_prefetch_desc(&map_qkv_k);
    tma_prefetch_desc(&map_qkv_mn);
    tma_prefetch_desc(&map_do_k);
    tma_prefetch_desc(&map_do_mn);
    tma_prefetch_desc(&map_dqkv);
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

  if (warp >= kEpiWarps) {
    // 640 threads: the kernel starts with 96 registers per thread; this warpgroup keeps them (the statistics warps need
    // them), the 16 elementwise warps take the 4096 spare ones (setmaxnreg.inc 104 below)
    if (warp == kEpiWarps) {
      // ================================================================= load producer
      int n = 0;
      for (int unit = blockIdx.x; unit < a.total; unit += gridDim.x, ++n) {
        const int slot = n % kSlots;
        const uint32_t use = (uint32_t)(n / kSlots);
        mbar_wait(&S.empty[slot], (use & 1) ^ 1);
        if (lane == 0) HS_TRACE(5, n, 0);
        SlotMeta& M = S.meta[slot];
        Slot& T = S.slot[slot];
        const int b = unit / a.nW, w = unit - b * a.nW;
        const long long s0 = (long long)w * kWS;
        int r0, r1, g0 = 0, g1 = 0;
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
          mbar_arrive_expect_tx(&S.full[slot], contig ? 7u * kTile : 0u);
          if (contig) {
            const int row = goff + rbase;
            tma_load_2d(T.qk, &map_qkv_k, &S.full[slot], h * kD, row);
            tma_load_2d(T.qk + kTile, &map_qkv_k, &S.full[slot], a.C + h * kD, row);
            tma_load_2d(T.dov, &map_do_k, &S.full[slot], h * kD, row);
            tma_load_2d(T.dov + kTile, &map_qkv_k, &S.full[slot], 2 * a.C + h * kD, row);
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
            cp_async16(T.qk + ok, g);
            cp_async16(T.qk + kTile + ok, g + a.C);
            cp_async16(T.dov + ok, gd);
            cp_async16(T.dov + kTile + ok, g + 2 * a.C);
            cp_async16(T.q_mn + om, g);
            cp_async16(T.k_mn + om, g + a.C);
            cp_async16(T.do_mn + om, gd);
          }
          cp_async_wait_all();
          fence_proxy_async_smem();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.full[slot]);  // arrival 2 of 2 (publishes the metadata too)
        if (lane == 0) HS_TRACE(5, n, 1);
      }
    } else if (warp == kEpiWarps + 1 && elect_one()) {
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
          const int slot = ns % kSlots, t = ns & 1;
          if (mbar_test_wait(&S.full[slot], (uint32_t)(ns / kSlots) & 1) &&
              mbar_test_wait(&S.stage_free[t], ((uint32_t)(ns >> 1) & 1) ^ 1)) {
            HS_TRACE(4, ns, 0);
            tc_fence_after();
            const uint32_t D1 = tmem + (uint32_t)t * kStageCols, D2 = D1 + 128;
            const uint32_t qk = smem_u32(S.slot[slot].qk), dov = smem_u32(S.slot[slot].dov);
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
          const int t = no & 1, slot = no % kSlots;
          const uint32_t ph = (uint32_t)(no >> 1) & 1;
          const uint32_t D1 = tmem + (uint32_t)t * kStageCols, D2 = D1 + 128;
          const Slot& T = S.slot[slot];
          if (ophase == 0 && mbar_test_wait(&S.dsn_ready[t], ph)) {
            HS_TRACE(4, no, 2);
            tc_fence_after();
            const uint32_t kb = smem_u32(T.k_mn);
#pragma unroll
            for (int s = 0; s < 8; ++s)  // dQ = dS k            A: D1[:, 64:128)  ->  D2[:, 64:96)
              umma_tf32_ts(D2 + 64, D1 + 64 + s * 8, umma_desc_at(kDescMN, kb + s * 1024), kIdescO, s > 0);
            ophase = 1;
            progressed = true;
          }
          if (ophase == 1 && mbar_test_wait(&S.dst_ready[t], ph)) {
            tc_fence_after();
            const uint32_t db = smem_u32(T.do_mn), qb = smem_u32(T.q_mn);
#pragma unroll
            for (int s = 0; s < 8; ++s)  // dV = P^T dO          A: D1[:, 0:64)    ->  D2[:, 96:128)
              umma_tf32_ts(D2 + 96, D1 + s * 8, umma_desc_at(kDescMN, db + s * 1024), kIdescO, s > 0);
            HS_TRACE(4, no, 3);
            // dK overwrites the dS block that dQ has read: tcgen05.mma instructions of one thread execute in issue order
#pragma unroll
            for (int s = 0; s < 8; ++s)  // dK = dS^T q          A: D2[:, 0:64)    ->  D1[:, 64:96)
              umma_tf32_ts(D1 + 64, D2 + s * 8, umma_desc_at(kDescMN, qb + s * 1024), kIdescO, s > 0);
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
    } else if (warp >= kEpiWarps + 2) {
      // ================================================================= statistics warps (18: even units, 19: odd units).
      // The softmax row statistics come from the forward pass -- lse from its saved vector, delta_i = sum_j P_ij dP_ij
      // = dO_i . O_i from its output -- so the elementwise warpgroups need no statistics sweep over S.  They start as
      // soon as the producer has published the unit's row table, i.e. in parallel with the TMA loads of the slot.
      int n = 0;
      for (int unit = blockIdx.x; unit < a.total; unit += gridDim.x, ++n) {
        if ((n & 1) != (warp & 1)) continue;
        const int slot = n % kSlots;
        mbar_wait(&S.meta_ready[slot], (uint32_t)(n / kSlots) & 1);
        const SlotMeta& M = S.meta[slot];
        const long long plane = (long long)a.H * ((long long)a.B * a.N);
#pragma unroll 1
        for (int k = 0; k < 2; ++k) {  // one 32-row half at a time: 16 float4 in flight per lane (96-register budget)
          float4 o[8], d[8];
          float qi = 1.f, ki = 1.f;
          const long long row = M.rows[lane + 32 * k];
          const float4* orow = reinterpret_cast<const float4*>(a.out + row * a.C + h * kD);
          const float4* drow = reinterpret_cast<const float4*>(a.dout + row * a.C + h * kD);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            o[c] = __ldg(orow + c);
            d[c] = __ldg(drow + c);
          }
          const float lse = __ldg(a.lse + (long long)h * ((long long)a.B * a.N) + row);
          if (a.cos) {
            qi = __ldg(a.lse + plane + (long long)h * ((long long)a.B * a.N) + row);
            ki = __ldg(a.lse + 2 * plane + (long long)h * ((long long)a.B * a.N) + row);
          }
          float dot = 0.f;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            dot += (o[c].x * d[c].x + o[c].y * d[c].y) + (o[c].z * d[c].z + o[c].w * d[c].w);
          S.lse[slot][lane + 32 * k] = lse;
          S.delta[slot][lane + 32 * k] = dot;
          S.inv[slot][lane + 32 * k] = qi;
          S.inv[slot][kWS + lane + 32 * k] = ki;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.stats_ready[slot]);
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // ================================================================= elementwise + epilogue warpgroups
    const int g4 = warp >> 2;              // warpgroup 0..3
    const int wg = g4 & 1;                 // handles units n with (n & 1) == wg, TMEM stage wg
    const int half = g4 >> 1;              // columns [32 half, 32 half + 32) of every row
    const int cb = half * (kWS / 2);
    const int L = (warp & 3) * 32 + lane;  // TMEM lane
    const bool nat = L < kWS;              // warps 0,1: query rows; warps 2,3: key rows
    const int r = L & 63;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t D1 = tmem + (uint32_t)wg * kStageCols + lane_addr, D2 = D1 + 128;
    const int half_bar = 5 + g4 * 2 + (nat ? 0 : 1);  // named barrier of the 64 threads of this (warpgroup, row kind)
    // output tile of this thread (one row of it): half 0 query rows dQ, half 0 key rows dK, half 1 key rows dV
    const int tile = nat ? (half == 0 ? 0 : -1) : (half == 0 ? 1 : 2);

    float racc = 0.f;  // running sum of dS o (logits without bias / mask): d logit_scale
    const float eff = a.cos ? __expf(fminf(__ldg(a.logit_scale + h), kLogitScaleMax)) : a.scale;
    bool store_pending = false;
    int pending_slot = -1;  // store-issuing threads (r == 0): slot whose staging tiles a TMA store may still be reading

    int n = 0;
    for (int unit = blockIdx.x; unit < a.total; unit += gridDim.x, ++n) {
      if ((n & 1) != wg) continue;
      const int slot = n % kSlots;
      const uint32_t it = (uint32_t)(n >> 1) & 1;
      if (pending_slot >= 0) {  // the staging tiles of my previous unit live in its slot: free it once they are read
        tma_store_wait_read<0>();
        mbar_arrive(&S.empty[pending_slot]);
        pending_slot = -1;
      }
      mbar_wait(&S.full[slot], (uint32_t)(n / kSlots) & 1);
      [[maybe_unused]] const int trole = wg * 2 + (nat ? 0 : 1);  // trace role (diagnostics build only; half 0 stamps)
      if (r == 0) HS_TRACE(trole, n, 0);

      const SlotMeta& M = S.meta[slot];
      Slot& T = S.slot[slot];
      const int flags = M.flags;
      const uint8_t* myrow = T.qk + L * 128;  // row L of [Q;K]: q_r for the query half, k_r for the key half

      const int my_row = M.rows[r];
      if (r == 0) HS_TRACE(trole, n, 6);
      mbar_wait(&S.stats_ready[slot], (uint32_t)(n / kSlots) & 1);  // lse / delta / norms of this unit (statistics warps)
      if (r == 0) HS_TRACE(trole, n, 7);
      const float my_inv = S.inv[slot][L];  // 1/|q_r| (query rows) or 1/|k_r| (key rows); 1 without cos attention
      const float row_scale = eff * kLog2e * my_inv * a.fix2;  // S = q k^T has two truncated operands

      mbar_wait(&S.s_ready[wg], it);
      if (r == 0) HS_TRACE(trole, n, 1);
      tc_fence_after();
      RowCtx R;
      R.s_src = D1 + (nat ? 64u : 0u);   // S (query rows) / S^T (key rows)
      R.dp_src = D2 + (nat ? 64u : 0u);  // dP / dP^T
      R.oinv = smem_u32(S.inv[slot] + (nat ? kWS : 0));
      R.brow = smem_u32(S.bias + (nat ? r * kBiasPitch : r));
      R.bstep = nat ? 4u : 4u * kBiasPitch;
      R.groups = smem_u32(M.groups);
      R.row_scale = row_scale;
      R.my_group = M.groups[r];
      R.cos = a.cos != 0;
      R.has_bias = has_bias;
      R.masked = !(flags & kFlagUniform);
      R.drop_thresh = a.drop_thresh;
      R.drop_scale = a.drop_scale;
      R.drop_key = a.drop_thresh ? hs::drop_unit_key(a.seed, unit, h, a.H) : 0u;
      R.drop_r = r;
      R.drop_row_is_query = nat;
      R.my_lse = nat ? S.lse[slot][r] : 0.f;
      R.my_delta = nat ? S.delta[slot][r] : 0.f;
      const uint32_t lse_v = smem_u32(S.lse[slot]);
      const uint32_t delta_v = smem_u32(S.delta[slot]);
      const uint32_t vstep = nat ? 0u : 1u;
      // P^T over S^T for the key rows (the query rows write P into a dead block); dS over S / dS^T over dP^T
      const RowSums part = ds_sweep<kDrop>(R, a.fix2, lse_v, delta_v, vstep, D1, nat ? D1 + 64 : D2,
                                           (nat && a.dbias) ? smem_u32(S.dbt[wg] + r * kDbtPitch) : 0u, r & 15, cb);
      // publish my partial row sums for the other half (ordered before the consumer's read by my arrive below, the MMA
      // issuer's wait and its commit on o_ready)
      S.xrs[wg][half][L] = part.rs;
      if (nat) S.xsp[wg][half][r] = make_float2(part.sds, part.pw);
      tc_fence_before();
      mbar_arrive(nat ? &S.dsn_ready[wg] : &S.dst_ready[wg]);
      if (r == 0 && half == 0) HS_TRACE(trole, n, 2);

      mbar_wait(&S.o_ready[wg], it);
      if (r == 0 && half == 0) HS_TRACE(trole, n, 3);
      tc_fence_after();
      // row sum over all 64 columns; along a query row the P-weighted mean of w is subtracted (see ds_sweep)
      float rs = S.xrs[wg][0][L] + S.xrs[wg][1][L];
      if (nat) {
        const float2 p0 = S.xsp[wg][0][r], p1 = S.xsp[wg][1][r];
        rs -= (p0.x + p1.x) * (p0.y + p1.y);
      }
      rs *= row_scale * (1.0f / kLog2e);  // sum_c dS[r][c] * (eff * cos(r, c))   (meaningful for cos attention only)
      if (nat && half == 0) racc += rs;
      // dQ / dK through the scaling / F.normalize:  d row = g * acc - row * corr;   dV = acc * fix1
      const float g = tile == 2 ? a.fix1 : eff * my_inv * a.fix1;  // dS or P^T (rounded) x k / q / dO (truncated)
      const bool clamped = my_inv >= 1.0f / kNormEps;
      const float corr = (a.cos && !clamped && tile != 2) ? my_inv * my_inv * rs : 0.f;
      const int row0 = M.rows[0];
      const bool contig = !kDirectStore && (flags & kFlagContig) != 0;
      uint8_t* st0 = (tile == 0 ? T.q_mn : tile == 1 ? T.k_mn : T.do_mn) + r * 128;
      float* g0 = a.dqkv + (long long)my_row * 3 * a.C + (tile > 0 ? tile : 0) * a.C + h * kD;
      const uint32_t osrc = tile == 0 ? D2 + 64 : tile == 1 ? D1 + 64 : D2 + 96;  // dQ, dK, dV
      if (tile >= 0) {  // warp-uniform
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {  // 16 columns at a time (register budget)
          uint32_t acc0[16];
          tmem_ld16(osrc + 16 * hh, acc0);
          tmem_wait_ld();
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const int c = 4 * hh + cc;
            float4 v4;
            v4.x = __uint_as_float(acc0[4 * cc + 0]) * g;
            v4.y = __uint_as_float(acc0[4 * cc + 1]) * g;
            v4.z = __uint_as_float(acc0[4 * cc + 2]) * g;
            v4.w = __uint_as_float(acc0[4 * cc + 3]) * g;
            if (a.cos && tile != 2) {
              const float4 x4 = *reinterpret_cast<const float4*>(myrow + ((c ^ (r & 7)) << 4));
              v4.x = fmaf(-x4.x, corr, v4.x);
              v4.y = fmaf(-x4.y, corr, v4.y);
              v4.z = fmaf(-x4.z, corr, v4.z);
              v4.w = fmaf(-x4.w, corr, v4.w);
            }
            if (contig)
              *reinterpret_cast<float4*>(st0 + ((c ^ (r & 7)) << 4)) = v4;
            else
              reinterpret_cast<float4*>(g0)[c] = v4;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&S.stage_free[wg]);
      if (r == 0 && half == 0) HS_TRACE(trole, n, 4);
      if (contig && tile >= 0) {
        fence_proxy_async_smem();
        named_bar_sync(half_bar, 64);
        if (r == 0) {
          tma_store_2d(&map_dqkv, tile == 0 ? T.q_mn : tile == 1 ? T.k_mn : T.do_mn, tile * a.C + h * kD, row0);
          tma_store_commit();
          store_pending = true;
          pending_slot = slot;  // released at the top of this thread's next unit, once the TMA engine has read it
        }
      }
      if (pending_slot != slot) mbar_arrive(&S.empty[slot]);
      if (r == 0) HS_TRACE(trole, n, 5);
    }
    if (pending_slot >= 0) {
      tma_store_wait_read<0>();
      mbar_arrive(&S.empty[pending_slot]);
    }
    if (store_pending) tma_store_wait<0>();

    if (a.dbias) {
      // one tile per TMEM stage: wait for all 16 elementwise warps, then one atomic per entry
      named_bar_sync(14, 512);
      float* gb = a.dbias + (long long)h * kWS * kWS;
      for (int idx = threadIdx.x; idx < kWS * kWS; idx += 512) {
        const int i = idx >> 6, j = idx & 63;
        const int pos = i * kDbtPitch + 4 * ((j >> 2) ^ (i & 15)) + (j & 3);
        atomicAdd(gb + idx, S.dbt[0][pos] + S.dbt[1][pos]);
      }
    }
    if (nat && half == 0 && a.cos && a.dlogit) {
      // d logit_scale = sum dS o logits_cos, zero where the clamp is active (torch.clamp backward)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) racc += __shfl_xor_sync(0xffffffffu, racc, o);
      if (lane == 0 && __ldg(a.logit_scale + h) <= kLogitScaleMax) atomicAdd(a.dlogit + h, racc);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps + 1) tmem_dealloc(tmem, kTmemCols);
}

}  // namespace

namespace hs {

int window_attn_bwd_tc_v2(const float* qkv, const float* out, const float* lse, const float* dout, const int32_t* src,
                       const uint8_t* groups, const float* bias, const float* logit_scale, float scale, DropCfg drop,
                       float* dqkv, float* dbias, float* dlogit, int B, int64_t N, int C, int H, uint32_t flags,
                       cudaStream_t stream) {
  HS_REQUIRE(qkv && dout && dqkv && out && lse, "hs_window_attn_bwd: null qkv/out/lse/dout/dqkv");
  HS_REQUIRE(!(flags & HS_ATTN_COS) || logit_scale, "hs_window_attn_bwd: cos attention needs logit_scale");
  CUtensorMap map_qkv_k, map_qkv_mn, map_do_k, map_do_mn, map_dqkv;
  const long long rows = (long long)B * N;
  int rc;
  if ((rc = make_map(&map_qkv_k, qkv, rows, 3 * C, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_map(&map_qkv_mn, qkv, rows, 3 * C, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
  if ((rc = make_map(&map_do_k, dout, rows, C, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_map(&map_do_mn, dout, rows, C, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
  if ((rc = make_map(&map_dqkv, dqkv, rows, 3 * C, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
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
  static bool attr_done = false;  // benign race: the attribute is idempotent
  if (!attr_done) {
    HS_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HS_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  int gx = sm_count() / H;
  if (gx < 1) gx = 1;
  if (gx > a.total) gx = a.total;
  dim3 grid(gx, H);
  if (a.drop_thresh)
    attn_bwd_tc_kernel<true><<<grid, kThreads, smem, stream>>>(map_qkv_k, map_qkv_mn, map_do_k, map_do_mn, map_dqkv, a);
  else
    attn_bwd_tc_kernel<false><<<grid, kThreads, smem, stream>>>(map_qkv_k, map_qkv_mn, map_do_k, map_do_mn, map_dqkv, a);
  HS_LAUNCH_CHECK();
  return HS_OK;
}

}  // namespace hs
