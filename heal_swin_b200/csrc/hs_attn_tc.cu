// Windowed attention core on the sm_100a tensor cores (tcgen05 + TMEM + TMA), window = 64 tokens,
// head_dim = 32, fp32 in / fp32 out, TF32 operands with fp32 accumulation.
//
// One work unit = one (window, head): S = Q K^T (64x64x32), P = softmax(S * scale + bias + mask),
// O = P V (64x32x64).  Two units (two consecutive windows of the same head) are stacked into one
// M = 128 accumulator tile, so every tcgen05.mma runs at the full 128-lane datapath width:
//   S pair   : A = [Q0;Q1] (128x32, K-major SW128), B = [K0;K1] (128x32, K-major SW128) -> 128x128 fp32 in TMEM,
//              only the two diagonal 64x64 blocks are used;
//   P        : written back over S in TMEM (tcgen05.st) and fed to the PV MMA as the A operand from TMEM;
//   O unit u : A = TMEM columns [64u, 64u+64), B = V_u (64x32, MN-major, SWIZZLE_128B_BASE32B) -> 128x32, rows of unit u valid.
// The HEALPix shift / window partition / window reverse is folded into the tile addressing: contiguous
// windows are fetched and written back by TMA (cp.async.bulk.tensor, 128B swizzle), permuted windows by
// 16-byte cp.async gathers / row stores.   Reference: swin_hp_transformer.py:136-171, 319-330.
//
// Warp roles (384 threads): warps 0-3 and 4-7 = two softmax/epilogue warpgroups (pair n -> group n&1, TMEM
// stage n&1), warp 8 = load producer, warp 9 = MMA issuer, warps 10-11 idle (they only donate registers).  Operand layouts and descriptor encodings were
// pinned on hardware with tools/probe_umma.cu (DESIGN.md "tcgen05 conventions").
#include <cfloat>

#include "hs_common.h"
#include "hs_kernels.h"
#include "hs_sm100.cuh"
#include "hs_tc_common.cuh"

namespace {

using namespace hs::sm100;
using namespace hs::tc;

constexpr int kSlots = 3;            // smem ring depth (pairs)
constexpr int kStageCols = 192;      // TMEM columns per stage: S/P 128 + O 2 x 32
constexpr int kTmemCols = 512;
constexpr int kThreads = 384;  // 3 warpgroups: 2 x softmax/epilogue, 1 x {producer, MMA, 2 idle warps}

struct SlotMeta {
  int rows[2][kWS];        // global row (b * N + token) of every slot of the two units
  uint8_t groups[2][kWS];  // mask group ids
  int flags[2];
  int pad[2];
};

struct Smem {
  uint8_t q[kSlots][2 * kTile];
  uint8_t k[kSlots][2 * kTile];
  uint8_t v[kSlots][2 * kTile];
  uint8_t o[2][2][kTile];  // [warpgroup][unit] output staging for the TMA store
  SlotMeta meta[kSlots];
  float kinv[2][2 * kWS];  // [warpgroup][pair row]: 1 / max(|k_j|, eps)
  uint64_t full[kSlots], empty[kSlots];
  uint64_t s_ready[2], p_ready[2], o_ready[2], stage_free[2];
  uint32_t tmem_base;
};

struct TcArgs {
  const float* qkv;
  float* out;
  float* lse;  // (P, H, B*N) or null: plane 0 log-sum-exp, planes 1, 2 (cos attention only) 1/|q|, 1/|k|
  const int32_t* src;
  const uint8_t* groups;
  const float* bias;         // (H, 64, 64) or null
  const float* logit_scale;  // (H) or null
  float scale;
  float fix1, fix2;  // TF32 truncation compensation (hs_tc_common.cuh), 1.0 when disabled
  uint32_t drop_thresh;  // attention-probability dropout (kDrop instantiation only), see hs_common.h
  float drop_scale;
  uint64_t seed;
  int B, nW, C, H, cos;
  long long N;
  int total;  // B * nW units per head
};


template <bool kDrop>
__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_qk, const __grid_constant__ CUtensorMap map_v,
                   const __grid_constant__ CUtensorMap map_o, const TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y;
  const int npairs = (a.total + 1) >> 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kSlots; ++i) {
      mbar_init(&S.full[i], 2);
      mbar_init(&S.empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&S.s_ready[i], 1);
      mbar_init(&S.p_ready[i], 128);
      mbar_init(&S.o_ready[i], 1);
      mbar_init(&S.stage_free[i], 128);
    }
    mbar_fence_init();
  }
  if (warp == 9) {
    tmem_alloc(&S.tmem_base, kTmemCols);
  }
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&map_qk);
    tma_prefetch_desc(&map_v);
    tma_prefetch_desc(&map_o);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;

  // register budget: the softmax warpgroups keep a bias row + a logit row per thread
  if (warp >= 8) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
  if (warp == 8) {
    // ================================================================= load producer
    int n = 0;
    for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x, ++n) {
      const int slot = n % kSlots;
      const uint32_t use = (uint32_t)(n / kSlots);
      mbar_wait(&S.empty[slot], (use & 1) ^ 1);
      SlotMeta& M = S.meta[slot];
      int contig[2] = {0, 0}, valid[2] = {0, 0};
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int wb = pair * 2 + u;
        int flags = 0;
        if (wb < a.total) {
          const int b = wb / a.nW, w = wb - b * a.nW;
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
          const bool c = __all_sync(0xffffffffu, (r0 == rbase + lane) && (r1 == rbase + 32 + lane));
          const bool un = __all_sync(0xffffffffu, (g0 == gbase) && (g1 == gbase));
          const int goff = (int)((long long)b * a.N);
          M.rows[u][lane] = goff + r0;
          M.rows[u][lane + 32] = goff + r1;
          M.groups[u][lane] = (uint8_t)g0;
          M.groups[u][lane + 32] = (uint8_t)g1;
          contig[u] = c ? 1 : 0;
          valid[u] = 1;
          flags = kFlagValid | (c ? kFlagContig : 0) | (un ? kFlagUniform : 0);
        }
        if (lane == 0) M.flags[u] = flags;
      }
      __syncwarp();
      // TMA part: arrival 1 of 2 carries the transaction bytes
      if (elect_one()) {
        const uint32_t tx = (uint32_t)((valid[0] & contig[0]) + (valid[1] & contig[1])) * 3u * kTile;
        mbar_arrive_expect_tx(&S.full[slot], tx);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (valid[u] && contig[u]) {
            const int row = M.rows[u][0];
            tma_load_2d(S.q[slot] + u * kTile, &map_qk, &S.full[slot], h * kD, row);
            tma_load_2d(S.k[slot] + u * kTile, &map_qk, &S.full[slot], a.C + h * kD, row);
            tma_load_2d(S.v[slot] + u * kTile, &map_v, &S.full[slot], 2 * a.C + h * kD, row);
          }
        }
      }
      // gathered part (shifted windows whose rows are not consecutive): 16 B cp.async, same swizzles
      bool gathered = false;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (valid[u] && !contig[u]) {
          gathered = true;
          const int c16 = lane & 7;
#pragma unroll 4
          for (int it = 0; it < 16; ++it) {
            const int r = it * 4 + (lane >> 3);
            const float* g = a.qkv + (long long)M.rows[u][r] * 3 * a.C + h * kD + c16 * 4;
            cp_async16(S.q[slot] + u * kTile + sw128_off(r, c16), g);
            cp_async16(S.k[slot] + u * kTile + sw128_off(r, c16), g + a.C);
            cp_async16(S.v[slot] + u * kTile + sw128b32_off(r, c16), g + 2 * a.C);
          }
        }
      }
      if (gathered) {
        cp_async_wait_all();
        fence_proxy_async_smem();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.full[slot]);  // arrival 2 of 2 (publishes the metadata too)
    }
  } else if (warp == 9) {
    // ================================================================= MMA issuer (one elected thread: with elect.sync
    // the compiler keeps the descriptors in uniform registers and emits back-to-back UTCHMMA instead of a per-MMA
    // election loop)
    if (elect_one()) {
      constexpr uint64_t kDescK = umma_smem_desc(16, 1024, kLayoutSw128);        // K-major, 8-row groups 1024 B apart
      constexpr uint64_t kDescV = umma_smem_desc(1024, 512, kLayoutSw128B32);    // MN-major, 4-row k-atoms 512 B apart
      constexpr uint32_t kIdescS = umma_idesc_tf32(128, 128, 0, 0);
      constexpr uint32_t kIdescO = umma_idesc_tf32(128, 32, 0, 1);
      auto issue_pv = [&](int m) {
        const int t = m & 1, slot = m % kSlots;
        mbar_wait(&S.p_ready[t], (uint32_t)(m >> 1) & 1);
        tc_fence_after();
        const uint32_t stage = tmem + (uint32_t)t * kStageCols;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const uint32_t vb = smem_u32(S.v[slot] + u * kTile);
#pragma unroll
          for (int s = 0; s < 8; ++s)
            umma_tf32_ts(stage + 128 + u * 32, stage + u * 64 + s * 8, umma_desc_at(kDescV, vb + s * 1024), kIdescO, s > 0);
        }
        umma_commit(&S.o_ready[t]);
        umma_commit(&S.empty[slot]);
      };
      int n = 0;
      for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x, ++n) {
        const int slot = n % kSlots, t = n & 1;
        mbar_wait(&S.full[slot], (uint32_t)(n / kSlots) & 1);
        mbar_wait(&S.stage_free[t], ((uint32_t)(n >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t qb = smem_u32(S.q[slot]), kb = smem_u32(S.k[slot]);
#pragma unroll
        for (int s = 0; s < 4; ++s)
          umma_tf32_ss(tmem + (uint32_t)t * kStageCols, umma_desc_at(kDescK, qb + s * 32), umma_desc_at(kDescK, kb + s * 32),
                       kIdescS, s > 0);
        umma_commit(&S.s_ready[t]);
        if (n > 0) issue_pv(n - 1);
      }
      if (n > 0) issue_pv(n - 1);
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    // ================================================================= softmax + epilogue warpgroups
    const int wg = warp >> 2;                 // 0 / 1 : handles pairs n with (n & 1) == wg, TMEM stage wg
    const int L = (warp & 3) * 32 + lane;     // TMEM lane = row of the pair tile
    const int u = L >> 6, i = L & 63;         // unit within the pair, query row within the window
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t stage = tmem + (uint32_t)wg * kStageCols;
    const int unit_bar = 1 + wg * 2 + u;      // named barrier of the 64 threads (2 warps) of this unit

    float bias_r[kWS];
    if (a.bias) {
      const float4* bp = reinterpret_cast<const float4*>(a.bias + ((long long)h * kWS + i) * kWS);
#pragma unroll
      for (int j = 0; j < kWS / 4; ++j) {
        const float4 b4 = __ldg(bp + j);
        bias_r[4 * j + 0] = b4.x * kLog2e;
        bias_r[4 * j + 1] = b4.y * kLog2e;
        bias_r[4 * j + 2] = b4.z * kLog2e;
        bias_r[4 * j + 3] = b4.w * kLog2e;
      }
    } else {
#pragma unroll
      for (int j = 0; j < kWS; ++j) bias_r[j] = 0.f;
    }
    const float base_scale = (a.cos ? __expf(fminf(__ldg(a.logit_scale + h), kLogitScaleMax)) : a.scale) * kLog2e * a.fix2;
    bool store_pending = false;

    int n = 0;
    for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x, ++n) {
      if ((n & 1) != wg) continue;
      const int slot = n % kSlots;
      const uint32_t it = (uint32_t)(n >> 1) & 1;
      mbar_wait(&S.full[slot], (uint32_t)(n / kSlots) & 1);
      const SlotMeta& M = S.meta[slot];
      const int flags = M.flags[u];
      const int my_row = M.rows[u][i];
      const int row0 = M.rows[u][0];
      const int my_group = M.groups[u][i];

      float row_scale = base_scale;
      if (a.cos) {
        // 1 / max(|q_i|, eps) and 1 / max(|k_i|, eps) from the rows of the staged tiles (chunk order is irrelevant)
        const uint8_t* qrow = S.q[slot] + L * 128;
        const uint8_t* krow = S.k[slot] + L * 128;
        float sq = 0.f, sk = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int cc = ((c + lane) & 7) << 4;  // rotate: conflict-free across the 8 lanes of a phase
          const float4 qv = *reinterpret_cast<const float4*>(qrow + cc);
          const float4 kv = *reinterpret_cast<const float4*>(krow + cc);
          sq += qv.x * qv.x + qv.y * qv.y + qv.z * qv.z + qv.w * qv.w;
          sk += kv.x * kv.x + kv.y * kv.y + kv.z * kv.z + kv.w * kv.w;
        }
        const float qinv = 1.0f / fmaxf(sqrtf(sq), kNormEps), kinv = 1.0f / fmaxf(sqrtf(sk), kNormEps);
        row_scale *= qinv;
        S.kinv[wg][L] = kinv;
        if (a.lse && (flags & kFlagValid)) {  // planes 1, 2 of the statistics buffer: the backward reuses the norms
          const long long plane = (long long)a.H * ((long long)a.B * a.N);
          a.lse[plane + (long long)h * ((long long)a.B * a.N) + my_row] = qinv;
          a.lse[2 * plane + (long long)h * ((long long)a.B * a.N) + my_row] = kinv;
        }
        named_bar_sync(unit_bar, 64);
      }

      mbar_wait(&S.s_ready[wg], it);
      tc_fence_after();
      uint32_t sr[kWS];
      tmem_ld32(stage + lane_addr + u * 64, sr);
      tmem_ld32(stage + lane_addr + u * 64 + 32, sr + 32);
      tmem_wait_ld();

      float x[kWS];
      if (a.cos) {
        const float4* kp = reinterpret_cast<const float4*>(&S.kinv[wg][u * 64]);
#pragma unroll
        for (int j = 0; j < kWS / 4; ++j) {
          const float4 k4 = kp[j];
          x[4 * j + 0] = fmaf(__uint_as_float(sr[4 * j + 0]) * row_scale, k4.x, bias_r[4 * j + 0]);
          x[4 * j + 1] = fmaf(__uint_as_float(sr[4 * j + 1]) * row_scale, k4.y, bias_r[4 * j + 1]);
          x[4 * j + 2] = fmaf(__uint_as_float(sr[4 * j + 2]) * row_scale, k4.z, bias_r[4 * j + 2]);
          x[4 * j + 3] = fmaf(__uint_as_float(sr[4 * j + 3]) * row_scale, k4.w, bias_r[4 * j + 3]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < kWS; ++j) x[j] = fmaf(__uint_as_float(sr[j]), row_scale, bias_r[j]);
      }
      if (!(flags & kFlagUniform)) {
        const uint32_t* gp = reinterpret_cast<const uint32_t*>(M.groups[u]);
#pragma unroll
        for (int j4 = 0; j4 < kWS / 4; ++j4) {
          const uint32_t g4 = gp[j4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if ((int)((g4 >> (8 * e)) & 0xff) != my_group) x[4 * j4 + e] += kMaskFill * kLog2e;
        }
      }
      float mx = x[0];
#pragma unroll
      for (int j = 1; j < kWS; ++j) mx = fmaxf(mx, x[j]);
      float sum = 0.f;
      uint32_t drop_key = 0;
      if (kDrop) drop_key = hs::drop_unit_key(a.seed, (long long)pair * 2 + u, h, a.H);
#pragma unroll
      for (int j = 0; j < kWS; ++j) {
        const float p = tf32_rna(ex2_approx(x[j] - mx));
        sum += p;  // softmax normalisation is over the undropped probabilities
        if (kDrop)
          sr[j] = hs::drop_keep(drop_key, i, j, kWS, a.drop_thresh) ? __float_as_uint(tf32_rna(p * a.drop_scale)) : 0u;
        else
          sr[j] = __float_as_uint(p);
      }
      tmem_st32(stage + lane_addr + u * 64, sr);
      tmem_st32(stage + lane_addr + u * 64 + 32, sr + 32);
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&S.p_ready[wg]);

      mbar_wait(&S.o_ready[wg], it);
      tc_fence_after();
      uint32_t orr[kD];
      tmem_ld32(stage + lane_addr + 128 + u * 32, orr);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(&S.stage_free[wg]);

      const float inv = a.fix1 / sum;
      if (a.lse && (flags & kFlagValid)) {  // log2-domain log-sum-exp of my logits row, for the backward
        float lg;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(sum));
        a.lse[(long long)h * ((long long)a.B * a.N) + my_row] = mx + lg;
      }
      if (flags & kFlagValid) {
        if (flags & kFlagContig) {
          // stage the tile (128B swizzle) and let one thread write it back with a TMA store
          if (store_pending) {
            if (i == 0) tma_store_wait_read<0>();
            named_bar_sync(unit_bar, 64);
          }
          uint8_t* orow = S.o[wg][u] + i * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float4 v4;
            v4.x = __uint_as_float(orr[4 * c + 0]) * inv;
            v4.y = __uint_as_float(orr[4 * c + 1]) * inv;
            v4.z = __uint_as_float(orr[4 * c + 2]) * inv;
            v4.w = __uint_as_float(orr[4 * c + 3]) * inv;
            *reinterpret_cast<float4*>(orow + ((c ^ (i & 7)) << 4)) = v4;
          }
          fence_proxy_async_smem();
          named_bar_sync(unit_bar, 64);
          if (i == 0) {
            tma_store_2d(&map_o, S.o[wg][u], h * kD, row0);
            tma_store_commit();
          }
          store_pending = true;
        } else {
          float4* dst = reinterpret_cast<float4*>(a.out + (long long)my_row * a.C + h * kD);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float4 v4;
            v4.x = __uint_as_float(orr[4 * c + 0]) * inv;
            v4.y = __uint_as_float(orr[4 * c + 1]) * inv;
            v4.z = __uint_as_float(orr[4 * c + 2]) * inv;
            v4.w = __uint_as_float(orr[4 * c + 3]) * inv;
            dst[c] = v4;
          }
        }
      }
    }
    if (store_pending && i == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, kTmemCols);
}

// ------------------------------------------------------------------------------------------ host side
}  // namespace

namespace hs {

bool window_attn_tc_supported(const float* qkv, const float* out, const float* mask, int B, int64_t N, int C, int H,
                              int ws) {
  if (ws != kWS || H <= 0 || C != H * kD || mask != nullptr) return false;
  if (N % kWS != 0 || (long long)B * N >= (1ll << 31)) return false;
  if ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out)) & 15) return false;
  return true;
}

int window_attn_fwd_tc(const float* qkv, const int32_t* src, const uint8_t* groups, const float* bias,
                       const float* logit_scale, float scale, DropCfg drop, float* out, float* lse, int B, int64_t N,
                       int C, int H, uint32_t flags, cudaStream_t stream) {
  HS_REQUIRE(qkv && out, "hs_window_attn_fwd: null qkv/out");
  HS_REQUIRE(!(flags & HS_ATTN_COS) || logit_scale, "hs_window_attn_fwd: cos attention needs logit_scale");
  CUtensorMap map_qk, map_v, map_o;
  int rc;
  if ((rc = make_map(&map_qk, qkv, (long long)B * N, 3 * C, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_map(&map_v, qkv, (long long)B * N, 3 * C, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
  if ((rc = make_map(&map_o, out, (long long)B * N, C, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  TcArgs a{};
  a.qkv = qkv; a.out = out; a.lse = lse; a.src = src; a.groups = groups; a.bias = bias; a.logit_scale = logit_scale;
  a.scale = scale; a.B = B; a.nW = (int)(N / kWS); a.C = C; a.H = H; a.cos = (flags & HS_ATTN_COS) ? 1 : 0;
  a.N = N; a.total = B * a.nW;
  a.fix1 = (flags & HS_ATTN_NO_TRUNC_COMP) ? 1.0f : kTruncFix1;
  a.fix2 = (flags & HS_ATTN_NO_TRUNC_COMP) ? 1.0f : kTruncFix2;
  a.drop_thresh = drop.p > 0.f ? hs::drop_thresh(drop.p) : 0u;
  a.drop_scale = 1.0f / (1.0f - drop.p);
  a.seed = drop.seed;
  const int npairs = (a.total + 1) / 2;
  const size_t smem = sizeof(Smem) + 1024;
  static bool attr_done = false;  // benign race: the attribute is idempotent
  if (!attr_done) {
    HS_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HS_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  int gx = sm_count() / H;
  if (gx < 1) gx = 1;
  if (gx > npairs) gx = npairs;
  dim3 grid(gx, H);
  if (a.drop_thresh)
    attn_fwd_tc_kernel<true><<<grid, kThreads, smem, stream>>>(map_qk, map_v, map_o, a);
  else
    attn_fwd_tc_kernel<false><<<grid, kThreads, smem, stream>>>(map_qk, map_v, map_o, a);
  HS_LAUNCH_CHECK();
  return HS_OK;
}

}  // namespace hs
