// Backward of the windowed attention core on the sm_100a tensor cores (tcgen05 + TMEM + TMA); window = 64
// tokens, head_dim = 32, fp32 in / fp32 out, TF32 operands with fp32 accumulation.  Adjoint of
// hs_attn_tc.cu (reference: autograd of swin_hp_transformer.py:136-171 + the shift / partition / reverse at
// :319-330).  P is recomputed from q, k (no forward state is needed besides qkv itself).
//
// One work unit = one (window, head).  Per unit, with S = q k^T, P = softmax(S*scale + bias + mask), dP = dO v^T,
// dS = P o (dP - rowsum(P o dP)):
//      dV = P^T dO      dQ = scale * dS k      dK = scale * dS^T q            (+ the F.normalize terms for cos attention)
// The TMEM operand of tcgen05.mma is the A matrix with M on the TMEM lanes, so dQ needs dS with the query index on
// the lanes while dV and dK need P^T / dS^T with the key index on the lanes.  Both orientations are produced by the
// tensor cores themselves, stacked into one M = 128 tile (lanes 0-63 "natural", lanes 64-127 "transposed"):
//      D1[:, 0:64)   = [Q;K] K^T   -> lanes 0-63  : S        D1[:, 64:128) = [Q;K] Q^T   -> lanes 64-127: S^T
//      D2[:, 0:64)   = [dO;V] V^T  -> lanes 0-63  : dP       D2[:, 64:128) = [dO;V] dO^T -> lanes 64-127: dP^T
// (the other half of each product is unused).  Threads 0-63 of a warpgroup own one query row each (softmax statistics,
// dS row, dbias row), threads 64-127 one key row each (P^T, dS^T from the row statistics published in shared memory).
// They write dS (over S), P^T (over S^T) and dS^T (over dP^T) back to TMEM as TF32, and three more MMAs with
// MN-major B tiles (K, dO, Q) produce dQ -> D2[:, 0:32), dV -> D2[:, 32:64), dK -> D1[:, 0:32).
//
// Every input tile is needed K-major (scores) and, except V, MN-major (outputs): they are fetched twice by TMA with the
// two swizzles (the second fetch hits L2).  HBM traffic per unit: q, k, v, dO in; dq, dk, dv out = 7 x 8 KB.
//
// Warp roles (384 threads): warps 0-3 / 4-7 = two elementwise warpgroups (unit n -> group n & 1, TMEM stage n & 1),
// warp 8 = load producer, warp 9 = MMA issuer.  3 shared-memory slots of 56 KB; the output tiles are staged in the
// slot's (dead) MN-major tiles and written back by TMA.
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
constexpr int kThreads = 384;
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
  float bias[2][kWS * kBiasPitch];  // [0]: bias[i][j], [1]: bias[j][i]; both pre-multiplied by log2(e)
  SlotMeta meta[kSlots];
  float inv[2][2 * kWS];  // per warpgroup: [0,64) 1/max(|q_i|,eps), [64,128) 1/max(|k_j|,eps)
  float lse[2][kWS];      // per warpgroup: log2-domain row log-sum-exp
  float delta[2][kWS];    // per warpgroup: rowsum(P o dP)
  uint64_t full[kSlots], empty[kSlots];
  uint64_t s_ready[2], ds_ready[2], x_ready[2], o_ready[2], stage_free[2];
  uint32_t tmem_base;
};

struct BwdArgs {
  const float* qkv;
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
  int B, nW, C, H, cos;
  long long N;
  int total;  // B * nW units per head
};

__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Elementwise stage of one unit for one thread.  kNat: thread owns query row r (lanes 0-63); else key row r.
template <bool kNat>
__device__ __forceinline__ void unit_elementwise(Smem& S, const BwdArgs& a, const int wg, const int r, const uint32_t D1,
                                                 const uint32_t D2, const SlotMeta& M, const int flags,
                                                 const float row_scale, const bool has_bias, float* dB, float& rs_out) {
  constexpr uint32_t cb = kNat ? 0u : 64u;  // my column block of D1 / D2
  const float* oinv = S.inv[wg] + (kNat ? kWS : 0);  // normalisation of the *other* index
  const float* brow = S.bias[kNat ? 0 : 1] + r * kBiasPitch;
  const int my_group = M.groups[r];

  float x[kWS];
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    uint32_t sr[32];
    tmem_ld32(D1 + cb + 32 * hf, sr);
    tmem_wait_ld();
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
      const int c = 32 * hf + 4 * c4;
      float4 o4 = make_float4(1.f, 1.f, 1.f, 1.f), b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a.cos) o4 = *reinterpret_cast<const float4*>(oinv + c);
      if (has_bias) b4 = *reinterpret_cast<const float4*>(brow + c);
      x[c + 0] = fmaf(__uint_as_float(sr[4 * c4 + 0]) * row_scale, o4.x, b4.x);
      x[c + 1] = fmaf(__uint_as_float(sr[4 * c4 + 1]) * row_scale, o4.y, b4.y);
      x[c + 2] = fmaf(__uint_as_float(sr[4 * c4 + 2]) * row_scale, o4.z, b4.z);
      x[c + 3] = fmaf(__uint_as_float(sr[4 * c4 + 3]) * row_scale, o4.w, b4.w);
    }
  }
  if (!(flags & kFlagUniform)) {
    const uint32_t* gp = reinterpret_cast<const uint32_t*>(M.groups);
#pragma unroll
    for (int c4 = 0; c4 < kWS / 4; ++c4) {
      const uint32_t g4 = gp[c4];
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if ((int)((g4 >> (8 * e)) & 0xff) != my_group) x[4 * c4 + e] += kMaskFill * kLog2e;
    }
  }

  float inv_l = 1.f, delta = 0.f;
  if (kNat) {
    // row statistics: x <- exp2(x - max), l = sum, delta = sum(p dP) / l
    float mx = x[0];
#pragma unroll
    for (int c = 1; c < kWS; ++c) mx = fmaxf(mx, x[c]);
    float l = 0.f;
#pragma unroll
    for (int c = 0; c < kWS; ++c) {
      x[c] = ex2_approx(x[c] - mx);
      l += x[c];
    }
    float dot = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t dr[16];
      tmem_ld16(D2 + cb + 16 * q, dr);
      tmem_wait_ld();
#pragma unroll
      for (int cc = 0; cc < 16; ++cc) dot = fmaf(x[16 * q + cc], __uint_as_float(dr[cc]), dot);
    }
    inv_l = 1.0f / l;
    delta = dot * inv_l * a.fix2;  // dP = dO v^T has two truncated operands
    S.lse[wg][r] = mx + lg2_approx(l);
    S.delta[wg][r] = delta;
    named_bar_arrive(3 + wg, 128);  // publish (bar.arrive orders the shared-memory writes)
  } else {
    named_bar_sync(3 + wg, 128);  // row statistics of all 64 query rows are in shared memory
  }

  float rs = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t dr[16], s2[16];
    tmem_ld16(D2 + cb + 16 * q, dr);
    if (a.cos) tmem_ld16(D1 + cb + 16 * q, s2);
    tmem_wait_ld();
    uint32_t pa[16];
#pragma unroll
    for (int cc = 0; cc < 16; ++cc) {
      const int c = 16 * q + cc;
      float p, dl;
      if (kNat) {
        p = x[c] * inv_l;
        dl = delta;
      } else {
        p = ex2_approx(x[c] - S.lse[wg][c]);
        dl = S.delta[wg][c];
      }
      const float ds = p * (__uint_as_float(dr[cc]) * a.fix2 - dl);
      if (kNat) dB[c] += ds;
      float av = ds;
      if (a.cos) {
        const float oi = oinv[c];
        rs = fmaf(ds, __uint_as_float(s2[cc]) * row_scale * oi, rs);
        av = ds * oi;
      }
      dr[cc] = __float_as_uint(tf32_rna(av));
      if (!kNat) pa[cc] = __float_as_uint(tf32_rna(p));
    }
    if (kNat) {
      tmem_st16(D1 + 16 * q, dr);  // dS (scaled by 1/|k_j| for cos) over S
    } else {
      tmem_st16(D1 + 64 + 16 * q, pa);  // P^T over S^T
      tmem_st16(D2 + 64 + 16 * q, dr);  // dS^T (scaled by 1/|q_i| for cos) over dP^T
    }
  }
  rs_out = rs;
  tmem_wait_st();
}

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
      mbar_init(&S.empty[i], 128);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&S.s_ready[i], 1);
      mbar_init(&S.ds_ready[i], 128);
      mbar_init(&S.x_ready[i], 1);
      mbar_init(&S.o_ready[i], 1);
      mbar_init(&S.stage_free[i], 128);
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
  }
  if (has_bias) {
    const float* bp = a.bias + (long long)h * kWS * kWS;
    for (int idx = threadIdx.x; idx < kWS * kWS; idx += kThreads) {
      const int i = idx >> 6, j = idx & 63;
      const float v = __ldg(bp + idx) * kLog2e;
      S.bias[0][i * kBiasPitch + j] = v;
      S.bias[1][j * kBiasPitch + i] = v;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;

  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    if (warp == 8) {
      // ================================================================= load producer
      int n = 0;
      for (int unit = blockIdx.x; unit < a.total; unit += gridDim.x, ++n) {
        const int slot = n % kSlots;
        const uint32_t use = (uint32_t)(n / kSlots);
        mbar_wait(&S.empty[slot], (use & 1) ^ 1);
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
        if (lane == 0) {
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
      }
    } else if (warp == 9 && lane == 0) {
      // ================================================================= MMA issuer (one thread)
      constexpr uint64_t kDescK = umma_smem_desc(16, 1024, kLayoutSw128);       // K-major, 8-row groups 1024 B apart
      constexpr uint64_t kDescMN = umma_smem_desc(1024, 512, kLayoutSw128B32);  // MN-major, 4-row k-atoms 512 B apart
      constexpr uint32_t kIdescS = umma_idesc_tf32(128, 64, 0, 0);
      constexpr uint32_t kIdescO = umma_idesc_tf32(128, 32, 0, 1);
      auto issue_outputs = [&](int m) {
        const int t = m & 1, slot = m % kSlots;
        const uint32_t ph = (uint32_t)(m >> 1) & 1;
        mbar_wait(&S.ds_ready[t], ph);
        tc_fence_after();
        const uint32_t D1 = tmem + (uint32_t)t * kStageCols, D2 = D1 + 128;
        const Slot& T = S.slot[slot];
        const uint32_t kb = smem_u32(T.k_mn), db = smem_u32(T.do_mn), qb = smem_u32(T.q_mn);
#pragma unroll
        for (int s = 0; s < 8; ++s)  // dQ = dS k
          umma_tf32_ts(D2, D1 + s * 8, umma_desc_at(kDescMN, kb + s * 1024), kIdescO, s > 0);
#pragma unroll
        for (int s = 0; s < 8; ++s)  // dV = P^T dO
          umma_tf32_ts(D2 + 32, D1 + 64 + s * 8, umma_desc_at(kDescMN, db + s * 1024), kIdescO, s > 0);
        umma_commit(&S.x_ready[t]);
        mbar_wait(&S.x_ready[t], ph);  // dQ has consumed dS before dK overwrites D1[:, 0:32)
        tc_fence_after();
#pragma unroll
        for (int s = 0; s < 8; ++s)  // dK = dS^T q
          umma_tf32_ts(D1, D2 + 64 + s * 8, umma_desc_at(kDescMN, qb + s * 1024), kIdescO, s > 0);
        umma_commit(&S.o_ready[t]);
      };
      int n = 0;
      for (int unit = blockIdx.x; unit < a.total; unit += gridDim.x, ++n) {
        const int slot = n % kSlots, t = n & 1;
        mbar_wait(&S.full[slot], (uint32_t)(n / kSlots) & 1);
        mbar_wait(&S.stage_free[t], ((uint32_t)(n >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t D1 = tmem + (uint32_t)t * kStageCols, D2 = D1 + 128;
        const uint32_t qk = smem_u32(S.slot[slot].qk), dov = smem_u32(S.slot[slot].dov);
#pragma unroll
        for (int s = 0; s < 4; ++s)  // [Q;K] K^T
          umma_tf32_ss(D1, umma_desc_at(kDescK, qk + s * 32), umma_desc_at(kDescK, qk + kTile + s * 32), kIdescS, s > 0);
#pragma unroll
        for (int s = 0; s < 4; ++s)  // [Q;K] Q^T
          umma_tf32_ss(D1 + 64, umma_desc_at(kDescK, qk + s * 32), umma_desc_at(kDescK, qk + s * 32), kIdescS, s > 0);
#pragma unroll
        for (int s = 0; s < 4; ++s)  // [dO;V] V^T
          umma_tf32_ss(D2, umma_desc_at(kDescK, dov + s * 32), umma_desc_at(kDescK, dov + kTile + s * 32), kIdescS, s > 0);
#pragma unroll
        for (int s = 0; s < 4; ++s)  // [dO;V] dO^T
          umma_tf32_ss(D2 + 64, umma_desc_at(kDescK, dov + s * 32), umma_desc_at(kDescK, dov + s * 32), kIdescS, s > 0);
        umma_commit(&S.s_ready[t]);
        if (n > 0) issue_outputs(n - 1);
      }
      if (n > 0) issue_outputs(n - 1);
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    // ================================================================= elementwise + epilogue warpgroups
    const int wg = warp >> 2;              // handles units n with (n & 1) == wg, TMEM stage wg
    const int L = (warp & 3) * 32 + lane;  // TMEM lane
    const bool nat = L < kWS;              // warps 0,1: query rows; warps 2,3: key rows
    const int r = L & 63;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t D1 = tmem + (uint32_t)wg * kStageCols + lane_addr, D2 = D1 + 128;
    const int half_bar = 5 + wg * 2 + (nat ? 0 : 1);  // named barrier of the 64 threads of this half

    float dB[kWS];  // running sum of dS[r][:] over this CTA's units (query-row threads only)
#pragma unroll
    for (int c = 0; c < kWS; ++c) dB[c] = 0.f;
    float racc = 0.f;  // running sum of dS o (logits without bias / mask): d logit_scale
    const float eff = a.cos ? __expf(fminf(__ldg(a.logit_scale + h), kLogitScaleMax)) : a.scale;
    bool store_pending = false;

    int n = 0;
    for (int unit = blockIdx.x; unit < a.total; unit += gridDim.x, ++n) {
      if ((n & 1) != wg) continue;
      const int slot = n % kSlots;
      const uint32_t it = (uint32_t)(n >> 1) & 1;
      mbar_wait(&S.full[slot], (uint32_t)(n / kSlots) & 1);
      const SlotMeta& M = S.meta[slot];
      Slot& T = S.slot[slot];
      const int flags = M.flags;
      const uint8_t* myrow = T.qk + L * 128;  // row L of [Q;K]: q_r for the query half, k_r for the key half

      float my_inv = 1.0f;
      if (a.cos) {
        float ss = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 v = *reinterpret_cast<const float4*>(myrow + (((c + lane) & 7) << 4));
          ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
        my_inv = 1.0f / fmaxf(sqrtf(ss), kNormEps);
        S.inv[wg][L] = my_inv;
        named_bar_sync(1 + wg, 128);
      }
      const float row_scale = eff * kLog2e * my_inv * a.fix2;  // S = q k^T has two truncated operands

      mbar_wait(&S.s_ready[wg], it);
      tc_fence_after();
      float rs = 0.f;
      if (nat)
        unit_elementwise<true>(S, a, wg, r, D1, D2, M, flags, row_scale, has_bias, dB, rs);
      else
        unit_elementwise<false>(S, a, wg, r, D1, D2, M, flags, row_scale, has_bias, dB, rs);
      tc_fence_before();
      mbar_arrive(&S.ds_ready[wg]);
      rs *= (1.0f / kLog2e);  // sum_c dS[r][c] * (eff * cos(r, c))
      if (nat) racc += rs;

      mbar_wait(&S.o_ready[wg], it);
      tc_fence_after();
      uint32_t acc0[kD], acc1[kD];
      if (nat) {
        tmem_ld32(D2, acc0);  // dQ
      } else {
        tmem_ld32(D1, acc0);       // dK
        tmem_ld32(D2 + 32, acc1);  // dV
      }
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(&S.stage_free[wg]);

      // through the scaling / F.normalize:  d row = g * acc - row * corr
      const float g = eff * my_inv * a.fix1;  // dS (rounded) x k / q (truncated)
      const bool clamped = my_inv >= 1.0f / kNormEps;
      const float corr = (a.cos && !clamped) ? my_inv * my_inv * rs : 0.f;
      const int my_row = M.rows[r];
      const int row0 = M.rows[0];
      const bool contig = (flags & kFlagContig) != 0;
      uint8_t* st0 = (nat ? T.q_mn : T.k_mn) + r * 128;
      uint8_t* st1 = T.do_mn + r * 128;
      float* g0 = a.dqkv + (long long)my_row * 3 * a.C + (nat ? 0 : a.C) + h * kD;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float4 v4;
        v4.x = __uint_as_float(acc0[4 * c + 0]) * g;
        v4.y = __uint_as_float(acc0[4 * c + 1]) * g;
        v4.z = __uint_as_float(acc0[4 * c + 2]) * g;
        v4.w = __uint_as_float(acc0[4 * c + 3]) * g;
        if (a.cos) {
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
      if (!nat) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float4 v4;
          v4.x = __uint_as_float(acc1[4 * c + 0]) * a.fix1;  // P^T (rounded) x dO (truncated)
          v4.y = __uint_as_float(acc1[4 * c + 1]) * a.fix1;
          v4.z = __uint_as_float(acc1[4 * c + 2]) * a.fix1;
          v4.w = __uint_as_float(acc1[4 * c + 3]) * a.fix1;
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
          tma_store_wait_read<0>();  // staging lives in the slot: release it only after the TMA engine has read it
          store_pending = true;
        }
      }
      mbar_arrive(&S.empty[slot]);
    }
    if (store_pending) tma_store_wait<0>();

    if (nat) {
      if (a.dbias) {
        float* gb = a.dbias + ((long long)h * kWS + r) * kWS;
#pragma unroll
        for (int c = 0; c < kWS; ++c) atomicAdd(gb + c, dB[c]);
      }
      if (a.cos && a.dlogit) {
        // d logit_scale = sum dS o logits_cos, zero where the clamp is active (torch.clamp backward)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) racc += __shfl_xor_sync(0xffffffffu, racc, o);
        if (lane == 0 && __ldg(a.logit_scale + h) <= kLogitScaleMax) atomicAdd(a.dlogit + h, racc);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, kTmemCols);
}

}  // namespace

namespace hs {

int window_attn_bwd_tc(const float* qkv, const float* dout, const int32_t* src, const uint8_t* groups,
                       const float* bias, const float* logit_scale, float scale, float* dqkv, float* dbias,
                       float* dlogit, int B, int64_t N, int C, int H, uint32_t flags, cudaStream_t stream) {
  HS_REQUIRE(qkv && dout && dqkv, "hs_window_attn_bwd: null qkv/dout/dqkv");
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
  a.qkv = qkv; a.dout = dout; a.dqkv = dqkv; a.src = src; a.groups = groups; a.bias = bias;
  a.logit_scale = logit_scale; a.dbias = dbias; a.dlogit = dlogit; a.scale = scale;
  a.B = B; a.nW = (int)(N / kWS); a.C = C; a.H = H; a.cos = (flags & HS_ATTN_COS) ? 1 : 0; a.N = N;
  a.total = B * a.nW;
  a.fix1 = (flags & HS_ATTN_NO_TRUNC_COMP) ? 1.0f : kTruncFix1;
  a.fix2 = (flags & HS_ATTN_NO_TRUNC_COMP) ? 1.0f : kTruncFix2;
  const size_t smem = sizeof(Smem) + 1024;
  static bool attr_done = false;  // benign race: the attribute is idempotent
  if (!attr_done) {
    HS_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  int gx = sm_count() / H;
  if (gx < 1) gx = 1;
  if (gx > a.total) gx = a.total;
  dim3 grid(gx, H);
  attn_bwd_tc_kernel<<<grid, kThreads, smem, stream>>>(map_qkv_k, map_qkv_mn, map_do_k, map_do_mn, map_dqkv, a);
  HS_LAUNCH_CHECK();
  return HS_OK;
}

}  // namespace hs
