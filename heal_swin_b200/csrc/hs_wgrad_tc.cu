// Weight gradient of a Linear on the sm_100a tensor cores: dW[n][k] += sum_t dY[t][n] * X[t][k]  (t = tokens).
// This is the one dense GEMM of the path where the library falls well short of the HBM roofline (cuBLAS picks a legacy
// split-K kernel: 1.8-3x the roofline time at the BASELINE stage-0/1 shapes, scripts/wgrad_check.py), because the
// contraction runs over the 10^5..10^6 tokens and both operands are "MN-major" (tokens are rows in memory).
//
// D[p][q] = sum_t P[t][p] * Q[t][q]:  P = the operand with the larger feature count (128-row blocks of D, one per CTA
// column), Q = the smaller one (<= 256 columns = one tcgen05.mma N).  Both tiles are fetched by TMA exactly as they lie in
// memory (64 tokens x 32 features per box, SWIZZLE_128B_ATOM_32B) and consumed as MN-major TF32 operands: a 32-feature
// slab is 64 rows of 128 B; descriptor LBO = slab stride (8192 B), SBO = 512 B, a K = 8 (token) step advances 1024 B
// (operand form pinned on hardware with tools/probe_mn.cu, profiles/r1j_probe_mn_major.log).  The token range is split
// over the CTAs (persistent accumulation in TMEM over a CTA's whole token range), partial results are added atomically.
//
// CTA pairs (template PAIR; the wide shapes of stages 2-3): per 32-token stage a CTA stages 16 KB of P and up to 48 KB of Q
// and the MMAs read all of it again -- 128 KB through a 128 B/clk shared-memory port against 768 cycles of TF32 MMA time.
// In a pair (cluster of 2, tcgen05.mma.cta_group::2, M = 256) the two CTAs take two neighbouring 128-row blocks of P over
// the same token range and each stages only HALF of the Q slabs; the leader issues, the peer's extra warp forwards the
// arrival of its loads to the leader's barriers (same protocol as hs_gemm3_tc.cu).
#include <cstdlib>

#include "hs_common.h"
#include "hs_sm100.cuh"
#include "hs_tc_common.cuh"

namespace {

using namespace hs::sm100;

// tokens per pipeline stage: 64 (one slab = 8 KB), or 32 when the smaller feature dimension exceeds 256 (up to 512: two
// MMAs of N <= 256 per K step into adjacent TMEM columns; 16 + 4 slabs of 4 KB per stage)
constexpr int kThreads = 192;            // warps 0-3 epilogue, 4 producer, 5 MMA (+ warp 6 of a pair: forwarder)
constexpr int kMaxStages = 4;

struct WgArgs {
  float* out;
  float* colsum;  // optional (NP): colsum[p] += sum_t P[t][p]  -- the bias gradient when P = dY; needs NQ + 32 <= 256
  long long ldo_p, ldo_q;  // out[p * ldo_p + q * ldo_q]
  long long T;             // tokens
  int NP, NQ;              // feature counts of P (rows of D) and Q (columns of D, <= 256, multiple of 32)
  int p_blocks, splits, stages;
  int tt;                  // tokens per stage (64 or 32)
  float fix, fix1;  // truncation compensation for two / one truncated operand
};

template <bool PAIR>
__global__ void __launch_bounds__(kThreads + (PAIR ? 32 : 0), 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_p, const __grid_constant__ CUtensorMap map_q, const WgArgs a) {
  constexpr int kAll = kThreads + (PAIR ? 32 : 0);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[kMaxStages], full2[kMaxStages], empty[kMaxStages], done;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kTT = a.tt, kSlab = a.tt * 128;
  const int q_slabs = (a.NQ + 31) / 32;  // a ragged last slab (NQ % 32 != 0) is zero-filled by TMA beyond the tensor
  const int ones = a.colsum ? 1 : 0;  // one extra Q slab holding the constant column (1, 0, ..., 0)
  // D columns: one MMA of N = NQ (+32 for the ones slab) when that is <= 256, else two halves (multiples of 32)
  const int n_all = 32 * (q_slabs + ones);
  const int n1 = n_all <= 256 ? n_all : ((q_slabs + 1) / 2) * 32;
  const int n2 = n_all - n1;
  const int crank = PAIR ? (int)cluster_ctarank() : 0;
  // Q slabs staged by this CTA: all of them, or (pair) my half of each MMA's slabs, the two parts back to back
  const int s1 = n1 / 32, s2 = n2 / 32;
  const int my_q = PAIR ? (s1 + s2) / 2 : q_slabs;
  const int stage_bytes = (4 + my_q + ones) * kSlab;
  int pb, split;
  if (PAIR) {  // a cluster takes the P blocks 2 j, 2 j + 1 over one token range
    const int cl = blockIdx.x >> 1, half_blocks = a.p_blocks >> 1;
    pb = 2 * (cl % half_blocks) + crank;
    split = cl / half_blocks;
  } else {
    pb = blockIdx.x % a.p_blocks;
    split = blockIdx.x / a.p_blocks;
  }
  const int p0 = pb * 128;
  int p_slabs = (a.NP - p0 + 31) / 32;  // slabs of this block that touch the tensor (the rest stay zero)
  if (p_slabs > 4) p_slabs = 4;
  const long long tiles = (a.T + kTT - 1) / kTT;
  const long long per = (tiles + a.splits - 1) / a.splits;
  const long long t_lo = split * per, t_hi = (t_lo + per < tiles) ? t_lo + per : tiles;

  // slabs of P that lie entirely outside the tensor are never loaded: zero them once
  for (int s = 0; s < a.stages; ++s)
    for (int i = threadIdx.x; i < (4 - p_slabs) * (kSlab / 16); i += kAll)
      reinterpret_cast<float4*>(sm + s * stage_bytes + p_slabs * kSlab)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ones)
    for (int s = 0; s < a.stages; ++s) {
      uint8_t* slab = sm + s * stage_bytes + (4 + q_slabs) * kSlab;
      for (int i = threadIdx.x; i < kSlab / 16; i += kAll) reinterpret_cast<float4*>(slab)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  __syncthreads();
  if (ones && threadIdx.x < kTT)  // feature 0 of every token row = 1 (16-byte chunk 0 of row r sits at (r & 3) << 5)
    for (int s = 0; s < a.stages; ++s)
      *reinterpret_cast<float*>(sm + s * stage_bytes + (4 + q_slabs) * kSlab + threadIdx.x * 128 + ((threadIdx.x & 3) << 5)) = 1.0f;
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&full2[s], 1);  // pair, leader: the peer's loads of the stage have landed
      mbar_init(&empty[s], 1);
    }
    mbar_init(&done, 1);
    mbar_fence_init();
  }
  const uint32_t tmem_cols = (n_all > 256) ? 512u : 256u;
  if (warp == 5) {
    if (PAIR) tmem_alloc_pair(&tmem_base, tmem_cols);
    else tmem_alloc(&tmem_base, tmem_cols);
  }
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&map_p);
    tma_prefetch_desc(&map_q);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // both CTAs' barriers exist before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem = tmem_base;

  if (PAIR && warp == 6) {
    if (crank == 1 && elect_one()) {  // peer: forward the completion of my loads to the leader
      int n = 0;
      for (long long t = t_lo; t < t_hi; ++t, ++n) {
        const int s = n % a.stages;
        mbar_wait(&full[s], ((uint32_t)(n / a.stages)) & 1);
        mbar_arrive_cluster(&full2[s], 0);
      }
    }
  } else if (warp == 4) {
    if (elect_one()) {
      int n = 0;
      for (long long t = t_lo; t < t_hi; ++t, ++n) {
        const int s = n % a.stages;
        mbar_wait(&empty[s], (((uint32_t)(n / a.stages)) & 1) ^ 1);
        uint8_t* st = sm + s * stage_bytes;
        mbar_arrive_expect_tx(&full[s], (uint32_t)((p_slabs + my_q) * kSlab));
        const int row = (int)(t * kTT);
        for (int j = 0; j < p_slabs; ++j) tma_load_2d(st + j * kSlab, &map_p, &full[s], p0 + 32 * j, row);
        if (PAIR) {  // my half of the first MMA's slabs, then my half of the second MMA's
          for (int j = 0; j < s1 / 2; ++j)
            tma_load_2d(st + (4 + j) * kSlab, &map_q, &full[s], 32 * (crank * (s1 / 2) + j), row);
          for (int j = 0; j < s2 / 2; ++j)
            tma_load_2d(st + (4 + s1 / 2 + j) * kSlab, &map_q, &full[s], 32 * (s1 + crank * (s2 / 2) + j), row);
        } else {
          for (int j = 0; j < q_slabs; ++j) tma_load_2d(st + (4 + j) * kSlab, &map_q, &full[s], 32 * j, row);
        }
      }
    }
  } else if (warp == 5) {
    if (elect_one() && (!PAIR || crank == 0)) {
      const uint64_t kDesc = umma_smem_desc((uint32_t)kSlab, 512, kLayoutSw128B32);
      constexpr int kM = PAIR ? 256 : 128;
      const uint32_t idesc1 = umma_idesc_tf32(kM, n1, 1, 1), idesc2 = umma_idesc_tf32(kM, n2 > 0 ? n2 : 32, 1, 1);
      const int q2_off = (PAIR ? s1 / 2 : n1 / 32) * kSlab;  // where the second MMA's B slabs start in a stage
      int n = 0;
      for (long long t = t_lo; t < t_hi; ++t, ++n) {
        const int s = n % a.stages;
        mbar_wait(&full[s], ((uint32_t)(n / a.stages)) & 1);
        if (PAIR) mbar_wait(&full2[s], ((uint32_t)(n / a.stages)) & 1);
        tc_fence_after();
        const uint32_t pa = smem_u32(sm + s * stage_bytes), qa = pa + 4 * kSlab;
        for (int ks = 0; ks < kTT / 8; ++ks) {
          const uint32_t acc = (n > 0 || ks > 0) ? 1u : 0u;
          const uint64_t dp = umma_desc_at(kDesc, pa + ks * 1024);
          if (PAIR) {
            umma_tf32_ss_pair(tmem, dp, umma_desc_at(kDesc, qa + ks * 1024), idesc1, acc);
            if (n2 > 0) umma_tf32_ss_pair(tmem + n1, dp, umma_desc_at(kDesc, qa + q2_off + ks * 1024), idesc2, acc);
          } else {
            umma_tf32_ss(tmem, dp, umma_desc_at(kDesc, qa + ks * 1024), idesc1, acc);
            if (n2 > 0) umma_tf32_ss(tmem + n1, dp, umma_desc_at(kDesc, qa + q2_off + ks * 1024), idesc2, acc);
          }
        }
        if (PAIR) umma_commit_pair(&empty[s], 3);
        else umma_commit(&empty[s]);
      }
      if (PAIR) umma_commit_pair(&done, 3);
      else umma_commit(&done);
    }
  } else if (warp < 4) {
    // ================================================================= epilogue: D tile -> global (atomic add)
    if (t_hi > t_lo) {
      mbar_wait(&done, 0);
      tc_fence_after();
      const int p = p0 + warp * 32 + lane;
      const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
      for (int c0 = 0; c0 < a.NQ; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem + lane_addr + c0, r);
        tmem_wait_ld();
        if (p < a.NP) {
          float* dst = a.out + (long long)p * a.ldo_p + (long long)c0 * a.ldo_q;
          const int nc = a.NQ - c0 < 32 ? a.NQ - c0 : 32;  // (NQ is a multiple of 4)
          if (a.ldo_q == 1) {  // my 32 columns are contiguous in memory: 128-bit vector reductions
#pragma unroll
            for (int c = 0; c < 32; c += 4)
              if (c < nc)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c),
                             "f"(__uint_as_float(r[c]) * a.fix), "f"(__uint_as_float(r[c + 1]) * a.fix),
                             "f"(__uint_as_float(r[c + 2]) * a.fix), "f"(__uint_as_float(r[c + 3]) * a.fix)
                             : "memory");
          } else {  // transposed output: consecutive lanes (rows p) are contiguous, one coalesced reduction per column
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (c < nc) atomicAdd(dst + (long long)c * a.ldo_q, __uint_as_float(r[c]) * a.fix);
          }
        }
      }
      if (ones) {  // column NQ of D = sum_t P[t][p] * 1
        uint32_t r[32];
        tmem_ld32(tmem + lane_addr + 32 * q_slabs, r);
        tmem_wait_ld();
        if (p < a.NP) atomicAdd(a.colsum + p, __uint_as_float(r[0]) * a.fix1);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // no CTA leaves (or frees tensor memory) while its peer may still use it
  if (warp == 5) {
    if (PAIR) tmem_dealloc_pair(tmem, tmem_cols);
    else tmem_dealloc(tmem, tmem_cols);
  }
}

}  // namespace

extern "C" {

// 0: not covered (the caller keeps the library GEMM); 1: dW covered; 2: dW and the fused bias gradient covered
int hs_linear_wgrad_supported(int64_t T, int N, int K) {
  if (T < 4096 || N < 4 || K < 4 || (N % 4) || (K % 4)) return 0;
  const int q = N < K ? N : K;
  if ((N < K ? K : N) < 32) return 0;
  if (q > 512) return (q <= 1024 && q % 64 == 0) ? 1 : 0;  // two launches over the column halves of the smaller operand
  // (a smaller operand that is not a multiple of 32 -- the patch embedding's 48 input features -- takes a zero-filled
  // last slab; above 256 columns the two-MMA split wants whole slabs)
  if (q % 32 != 0 && q > 224) return 0;
  return (N >= K && (K + 31) / 32 * 32 + 32 <= 256) ? 2 : 1;
}

int hs_linear_wgrad(const float* dy, const float* x, float* dw, float* dbias, int64_t T, int N, int K, uint32_t flags,
                    void* stream) {
  HS_REQUIRE(dy && x && dw && T > 0 && N > 0 && K > 0, "hs_linear_wgrad: bad arguments");
  if (!hs_linear_wgrad_supported(T, N, K))
    return hs::fail(HS_ERR_UNSUPPORTED, "hs_linear_wgrad: shape T=%lld N=%d K=%d is not covered (see "
                    "hs_linear_wgrad_supported in include/healswin_b200.h)", (long long)T, N, K);
  HS_REQUIRE(!((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x)) & 15), "hs_linear_wgrad: unaligned input");
  // P = operand with more features (rows of D), Q = the other (<= 512 columns per launch)
  const bool p_is_dy = N >= K;
  const float* P = p_is_dy ? dy : x;
  const float* Q = p_is_dy ? x : dy;
  const int NP = p_is_dy ? N : K, NQ_all = p_is_dy ? K : N;
  // a smaller operand wider than 512 columns (stage 3: 768) does not fit the accumulator: one launch per column half,
  // each reading its half of Q through the row pitch and writing its column block of dW
  const int parts = NQ_all > 512 ? 2 : 1;
  const int NQ = NQ_all / parts;
  if (dbias && !(p_is_dy && parts == 1 && (NQ + 31) / 32 * 32 + 32 <= 256))
    return hs::fail(HS_ERR_UNSUPPORTED, "hs_linear_wgrad: the fused bias gradient needs N >= K and K <= 224 (N=%d K=%d)", N, K);
  for (int part = 0; part < parts; ++part) {
    WgArgs a{};
    a.T = T;
    a.NP = NP; a.NQ = NQ;
    a.colsum = dbias;
    a.ldo_p = p_is_dy ? K : 1; a.ldo_q = p_is_dy ? 1 : K;
    a.out = dw + (long long)part * NQ * a.ldo_q;
    a.p_blocks = (a.NP + 127) / 128;
    a.tt = (a.NQ + 31) / 32 * 32 + (dbias ? 32 : 0) <= 256 ? 64 : 32;
    const int kTT = a.tt, kSlab = a.tt * 128;
    const long long tiles = (T + kTT - 1) / kTT;
    int splits = hs::tc::sm_count() / a.p_blocks;
    if (splits < 1) splits = 1;
    if (splits > tiles) splits = (int)tiles;
    a.splits = splits;
    // CTA pairs where the stage is dominated by Q (>= 6 slabs), the P blocks come in twos and every MMA's slabs halve
    const int q_slabs = (a.NQ + 31) / 32, s1 = 32 * q_slabs <= 256 ? q_slabs : (q_slabs + 1) / 2, s2 = q_slabs - s1;
    bool pair = !dbias && q_slabs >= 6 && a.p_blocks % 2 == 0 && s1 % 2 == 0 && s2 % 2 == 0 && splits >= 1;
    if (const char* e = getenv("HEALSWIN_WGRAD_PAIR")) pair = pair && atoi(e) != 0;
    const int stage_bytes = (4 + (pair ? q_slabs / 2 : q_slabs) + (dbias ? 1 : 0)) * kSlab;
    int stages = (200 * 1024) / stage_bytes;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) return hs::fail(HS_ERR_UNSUPPORTED, "hs_linear_wgrad: stage does not fit");
    a.stages = stages;
    a.fix = (flags & HS_ATTN_NO_TRUNC_COMP) ? 1.0f : hs::tc::kTruncFix2;
    a.fix1 = (flags & HS_ATTN_NO_TRUNC_COMP) ? 1.0f : hs::tc::kTruncFix1;
    CUtensorMap map_p, map_q;
    int rc;
    if ((rc = hs::tc::make_map(&map_p, P, T, a.NP, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, 32, kTT))) return rc;
    if ((rc = hs::tc::make_map(&map_q, Q + (long long)part * NQ, T, a.NQ, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, 32, kTT,
                               NQ_all)))
      return rc;
    const size_t smem = (size_t)stages * stage_bytes + 1024;
    if (pair) {
      HS_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((unsigned)(a.p_blocks * a.splits));
      cfg.blockDim = dim3(kThreads + 32);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = (cudaStream_t)stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      HS_CUDA(cudaLaunchKernelEx(&cfg, wgrad_tc_kernel<true>, map_p, map_q, a));
    } else {
      HS_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      wgrad_tc_kernel<false><<<a.p_blocks * a.splits, kThreads, smem, (cudaStream_t)stream>>>(map_p, map_q, a);
    }
    HS_LAUNCH_CHECK();
  }
  return HS_OK;
}

}  // extern "C"
