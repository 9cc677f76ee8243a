// Dense linear layers of the hot path on the sm_100a tensor cores at fp32-class accuracy ("bf16x3"):
//     D[t][n] = sum_k A[t][k] * W[n][k]   (+ epilogue)        A: (T, K) fp32, W: (N, K), D: (T, N) fp32
// Serves the forward of every nn.Linear of the block (qkv / proj: swin_hp_transformer.py:131-135, 172; Mlp fc1 / fc2:
// :21-44; PatchMerging.reduction :394; PatchExpand.expand :421; FinalPatchExpand_X4.expand :444; concat_back_dim :774)
// and, with the transposed split weight, their input gradients.
//
// Why not kind::tf32: tcgen05 TF32 truncates both operands to 10 mantissa bits; through the 22 blocks of the network
// that puts the output 1.1-1.4e-3 from the fp32 reference (tolerance 1e-3).  Here every fp32 operand is split into two
// bf16 terms, x = hi + lo (+ O(2^-17 x)), and the product is accumulated in fp32 from three kind::f16 MMAs
//     A_hi W_hi + A_lo W_hi + A_hi W_lo                    (the dropped lo*lo term is O(2^-16))
// i.e. ~2^-16 relative per product instead of 2^-11, at 1.5x the tensor-pipe time of one TF32 MMA (bf16 runs at twice
// the TF32 rate) -- free where the GEMM is HBM-bound (stages 0-1).
//
// Operand staging.  W is split once per optimizer step by hs_weight_split into rows of [hi(32) | lo(32)] bf16 per
// 32-wide K chunk (128 B: one SWIZZLE_128B row), so it comes in by TMA ready to use as the shared-memory B operand (hi at
// +0/+32 B, lo at +64/+96 B for the two K=16 steps).  A arrives as raw fp32: a 128-row x 32-column chunk (16 KB, TMA
// SWIZZLE_128B) is read by 8 converter warps (thread = row, two warps per TMEM lane quadrant, 16 columns each), split, and
// written with tcgen05.st into a ring of 32-column TENSOR-MEMORY slots (columns 0-15 hi, 16-31 lo, two bf16 per column):
// the MMAs take A from tensor memory.  Shared memory bandwidth (128 B/clk) is the co-limiter of this kernel -- with A in
// shared memory the operand reads of the three MMAs were 35 % of all shared-memory traffic.
//
// A CTA owns one column chunk (<= 192 columns) and walks over 128-token tiles; its W chunk stays resident in shared
// memory when it fits, otherwise its K slices stream through a short ring of their own (L2 hits).  Two 192-column
// TMEM stages overlap the epilogue of one tile with the MMAs of the next.  The epilogue (groups of 4 warps, one 32-column
// slab at a time; every warp has private 4 KB staging regions and issues its own TMA stores / aux loads, so the warps never
// synchronise with each other) goes TMEM -> registers -> swizzled staging region -> TMA store, and can
//     MODE_PLAIN      add the bias
//     MODE_ADD        add the bias and an fp32 tensor of D's shape (aux) -- the residual-shortcut gradient in a dgrad
//     MODE_GELU       write z = acc (D) and h = dropout(GELU(acc + bias)) (D2): Mlp fc1 + act + drop in one pass
//     MODE_GELU_GRAD  D = acc * GELU'(aux + bias) * dropmask: the fc2 input gradient through the activation
//     MODE_GELU_C     as MODE_GELU, but instead of z the epilogue writes g' = GELU'(acc + bias) * dropmask as FP16 (a quarter
//                     of the bytes of z and h together; g' lies in [-0.13, 1.13] / (1 - p)): all the backward needs
//     MODE_GELU_GRAD_C  D = acc * g' with that FP16 tensor read straight from global memory (no GELU arithmetic, no mask)
//     MODE_LN         D2 = [aux +] LayerNorm_G(acc + bias) * gamma + beta over groups of G consecutive output columns, with
//                     D = acc + bias (the pre-norm tensor the backward needs) and the row statistics as optional outputs:
//                     PatchExpand's Linear -> view (B, 4N, C/2) -> LayerNorm (swin_hp_transformer.py:420-430) and the
//                     `x + norm(branch(x))` tail of a v2 block (:333-338) without a pass of their own over (T, N)
//     MODE_LN_IN      D = LayerNorm_K(a) W^T for a LayerNorm over the whole input row: PatchMerging's view -> LayerNorm(4C) ->
//                     Linear(4C -> 2C) (swin_hp_transformer.py:378-395) without materialising the normalised tensor.  With
//                     W' = W diag(gamma):  D = rstd (a W'^T - mean s) + b0,  s = W' 1,  b0 = W beta  -- the GEMM runs on
//                     the RAW rows, the converters (which see every element anyway) take the row statistics, and the
//                     epilogue applies them.
// aux sub-slabs are brought in by TMA into the staging regions ahead of time and combined in place.
//
// Warps: [0, E) epilogue (TMEM lane quadrant = warp % 4), [E, E+8) converters, E+8 A producer, E+9 MMA issuer,
// E+10 W producer (+ E+11, pair kernels: forwards the arrival of the peer's W half-slices to the leader).
//
// CTA pairs (template PAIR; the tensor-bound "ss" launches): the two CTAs of a cluster take neighbouring token tiles of the
// same column chunk and run tcgen05.mma.cta_group::2 (M = 256) issued by the leader; each stages only half of every W
// slice, which halves the B-operand traffic on its shared-memory port.
#include <cstdlib>

#include <cuda_fp16.h>

#include "hs_common.h"
#include "hs_gelu.cuh"
#include "hs_sm100.cuh"
#include "hs_tc_common.cuh"

namespace {

using namespace hs::sm100;

constexpr int kBM = 128;               // tokens per tile
constexpr int kChunk = kBM * 128;      // 16 KB: one A chunk, one staging slab
constexpr int kMaxRing = 10;
constexpr int kMaxWRing = 6;            // streamed W slices (L2 hits: a short ring is enough)
constexpr int kMaxRw = 3;              // staging regions per epilogue warp
constexpr int kRegion = 32 * 128;      // 4 KB: 32 rows x 32 columns, one warp's part of a slab
constexpr int kCvtWarps = 8;
constexpr int kTeam = 4;               // converter warps per chunk (one per TMEM lane quadrant); two teams alternate
constexpr int kASlots = 4;             // A-operand slots of 32 TMEM columns behind two 192-column accumulator stages
constexpr int kACol0 = 2 * 192;

enum : int { MODE_PLAIN = 0, MODE_ADD = 1, MODE_GELU = 2, MODE_GELU_GRAD = 3, MODE_LN = 4, MODE_LN_IN = 5, MODE_GELU_C = 6,
              MODE_GELU_GRAD_C = 7 };
constexpr int kLnExch = 2 * 2 * 128 * 2 * 8;  // MODE_LN: partial row statistics exchanged between the two epilogue groups
// operand precision: three bf16 MMAs per product (fp32-class), one TF32 MMA (A straight from the fp32 tile, no conversion;
// input gradients of the tensor-bound stages), one bf16 MMA ("bf16 operands, fp32 accumulate": BASELINE configs[3])
enum : int { PREC_BF16X3 = 0, PREC_TF32 = 1, PREC_BF16 = 2 };

struct G3Args {
  const float* bias;  // (N) or null
  float* colsum;      // (K) or null: += column sums of A (the bias gradient when A is the output gradient of a linear)
  long long T;
  int N, K;
  int n_stride, n_box, n_chunks;  // column chunks start every n_stride columns and compute n_box (multiple of 32)
  long long tiles;
  int prec;     // PREC_*
  int ss;       // 1: the A operand stays in shared memory (TF32: as delivered; bf16: split in place), 256-column stages
  int stage_cols;  // TMEM columns per accumulator stage: 192 (+ 4 A slots of 32 columns) or 256 (ss)
  int cluster;  // CTAs per cluster (1, 2 or 4): streamed W slices are loaded once per cluster and multicast
  int pair;     // 1: the two CTAs of a cluster run cta_group::2 MMAs (M = 256; each holds half of the W slice); ss modes only
  int ring, rw, resident, wring;  // A ring depth, staging regions per epilogue warp, W resident?, W ring depth
  uint32_t drop_thresh;
  float drop_scale;
  uint64_t seed;
  // MODE_LN: LayerNorm over groups of G output columns (G a multiple of 32, <= 192, dividing N and the column chunk)
  const float* gamma;  // (G)
  const float* beta;   // (G)
  float* mean_out;     // (T * N / G) or null: row r of the (T N / G, G) view = token r / (N / G), group r % (N / G)
  float* rstd_out;
  int G;
  float eps;
  int save_pre;  // D (= acc + bias) is written as well
  int has_aux;   // D2 += aux
  // MODE_LN_IN: s[n] = sum_k gamma[k] W[n][k] (bias holds b0 = W beta; mean_out / rstd_out (T) optional; eps as above)
  const float* wsum;
  // MODE_GELU_C (written) / MODE_GELU_GRAD_C (read): g' (T, N) as FP16, row-major; N a multiple of 32
  uint16_t* gp16;
};

// instruction descriptor: D = f32, A = B = bf16, both K-major
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem], TF32 operands (fp32 words, the tensor core reads their top 19 bits)
__device__ __forceinline__ void umma_tf32_ss_(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 operands
__device__ __forceinline__ void umma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// {bf16(hi_of_pair), bf16(lo_of_pair)}: first argument lands in the upper 16 bits
__device__ __forceinline__ uint32_t cvt_bf16x2(float upper, float lower) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(upper), "f"(lower));
  return d;
}
// split two consecutive fp32 values (e0 at the lower address) into a packed hi word and a packed lo word
__device__ __forceinline__ void split2(float e0, float e1, uint32_t& hi, uint32_t& lo) {
  hi = cvt_bf16x2(e1, e0);
  const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
  lo = cvt_bf16x2(e1 - h1, e0 - h0);
}

// threads per CTA: E epilogue warps, 8 converters, A producer, MMA issuer, W producer (+ the W forwarder of a pair)
template <int E, bool PAIR>
constexpr int block_threads() { return (E + kCvtWarps + 3 + (PAIR ? 1 : 0)) * 32; }

template <int E, int MODE, bool PAIR>
__global__ void __launch_bounds__(block_threads<E, PAIR>(), 1)
gemm3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
             const __grid_constant__ CUtensorMap map_aux, const __grid_constant__ CUtensorMap map_d,
             const __grid_constant__ CUtensorMap map_d2, const G3Args a) {
  constexpr bool kAux = (MODE == MODE_ADD || MODE == MODE_GELU_GRAD);
  constexpr int NG = E / 4;                    // epilogue groups (4 warps = the 4 TMEM lane quadrants)
  constexpr int kStores = (MODE == MODE_GELU || MODE == MODE_GELU_C) ? 2 : 1;
  constexpr bool kColsumOk = (MODE == MODE_PLAIN || MODE == MODE_ADD);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int nk = (a.K + 31) / 32;              // K chunks (TMA zero-fills the tail of a ragged last chunk)
  constexpr bool pair = PAIR;  // (a kernel that contains cta_group::2 instructions can only be launched as a cluster)
  const int w_slice = (pair ? a.n_box / 2 : a.n_box) * 128;  // bytes of one K slice of the W chunk (pair: this CTA's half)
  uint8_t* s_w = sm;                           // resident: nk slices; streamed: a ring of wring slices
  uint8_t* s_ring = s_w + (a.resident ? nk : a.wring) * w_slice;
  uint8_t* s_buf = s_ring + a.ring * kChunk;   // E warps x rw regions of 4 KB
  float* s_colsum = reinterpret_cast<float*>(s_buf + E * a.rw * kRegion);  // nk * 32 floats (chunk-0 CTAs with a.colsum)
  __shared__ uint64_t w_full, wr_full[kMaxWRing], wr_full2[kMaxWRing], wr_empty[kMaxWRing], raw_full[kMaxRing], slot_empty[kMaxRing], cvt_full[kMaxRing],
      a_full[kASlots], a_empty[kASlots], acc_full[2], acc_empty[2], aux_full[E * kMaxRw];
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // CTAs of a cluster share the column chunk and take neighbouring token tiles; every CTA of a cluster runs the same number
  // of tile iterations (a tile index beyond the end is harmless: TMA zero-fills its loads and clips its stores)
  const int cs = a.cluster, crank = cs > 1 ? (int)cluster_ctarank() : 0;
  const int chunk = (blockIdx.x / cs) % a.n_chunks;
  const long long t0 = (long long)(blockIdx.x / (cs * a.n_chunks)) * cs + crank, tstep = gridDim.x / a.n_chunks;
  const long long t_end = cs > 1 ? t0 + ((a.tiles + tstep - 1) / tstep) * tstep : a.tiles;
  const uint16_t cmask = (uint16_t)((1u << cs) - 1);
  // column chunk of this CTA: starts every n_stride columns and computes n_box >= n_stride of them (a multiple of 32);
  // where n_box > n_stride the first columns of the next chunk are computed -- and stored -- twice, with identical values
  const int n0 = chunk * a.n_stride;
  const int n_this = min(a.n_box, a.N - n0);
  const int S = (n_this + 31) / 32;            // 32-column slabs (stores clip a ragged last slab at N)
  const int ring = a.ring;

  if (threadIdx.x == 0) {
    mbar_init(&w_full, 1);
    for (int i = 0; i < kMaxWRing; ++i) {
      mbar_init(&wr_full[i], 1);
      mbar_init(&wr_full2[i], 1);   // pair, leader: the peer's half of the slice has landed (remote arrive)
      // released by the MMA issuers of all CTAs of the cluster (multicast commit); pair: by the leader's commit alone
      mbar_init(&wr_empty[i], pair ? 1 : cs);
    }
    for (int i = 0; i < kMaxRing; ++i) {
      mbar_init(&raw_full[i], 1);
      // released by the converters once the chunk is in their registers; in TF32 mode by the MMAs that read it (and by the
      // converters as well when they take the column sums)
      mbar_init(&slot_empty[i], a.prec == PREC_TF32 ? 1 + ((kColsumOk && a.colsum && chunk == 0) ? kTeam : 0)
                                                    : (a.ss ? 1 : kTeam));
    }
    for (int i = 0; i < kASlots; ++i) {
      mbar_init(&a_full[i], kTeam);
      mbar_init(&a_empty[i], 1);
    }
    // ss bf16: the chunk is split in place.  pair: the LEADER's barrier collects both CTAs' chunks (TF32: one forwarder each)
    for (int i = 0; i < kMaxRing; ++i) mbar_init(&cvt_full[i], pair ? 2 * (a.prec == PREC_TF32 ? 1 : kTeam) : kTeam);
    {
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], pair ? 2 * E : E * 32);  // pair: one arrive per epilogue warp of both CTAs, on the leader
    }
    for (int i = 0; i < E * kMaxRw; ++i) mbar_init(&aux_full[i], 1);
    mbar_fence_init();
  }
  if (warp == E + kCvtWarps + 1) {
    if (pair) tmem_alloc_pair(&tmem_base, 512);
    else tmem_alloc(&tmem_base, 512);
  }
  if (warp == E + kCvtWarps && lane == 0) tma_prefetch_desc(&map_a);
  if (warp == E + kCvtWarps + 2 && lane == 0) tma_prefetch_desc(&map_w);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_d);
    if (kAux || MODE == MODE_LN) tma_prefetch_desc(&map_aux);
    if (MODE == MODE_GELU || MODE == MODE_GELU_C || MODE == MODE_LN) tma_prefetch_desc(&map_d2);
  }
  tc_fence_before();
  __syncthreads();
  if (cs > 1) cluster_sync_all();  // every CTA's barriers are initialised before any multicast touches them
  tc_fence_after();
  const uint32_t tmem = tmem_base;

  if (warp == E + kCvtWarps) {
    // ================================================================ A producer
    if (elect_one()) {
      int slot = 0;
      uint32_t ph = 0;
      for (long long tile = t0; tile < t_end; tile += tstep)
        for (int kc = 0; kc < nk; ++kc) {
          mbar_wait(&slot_empty[slot], ph ^ 1);
          mbar_arrive_expect_tx(&raw_full[slot], (uint32_t)kChunk);
          tma_load_2d(s_ring + slot * kChunk, &map_a, &raw_full[slot], 32 * kc, (int)(tile * kBM));
          if (++slot == ring) { slot = 0; ph ^= 1; }
        }
    }
  } else if (warp == E + kCvtWarps + 2) {
    // ================================================================ W producer: resident chunk once, or a ring of K slices
    if (elect_one()) {
      if (a.resident) {
        mbar_arrive_expect_tx(&w_full, (uint32_t)(nk * w_slice));
        for (int kc = 0; kc < nk; ++kc) tma_load_2d(s_w + kc * w_slice, &map_w, &w_full, 64 * kc, n0);
      } else {
        int ws = 0;
        uint32_t wph = 0;
        // with a cluster this CTA loads rows [crank, crank + 1) * n_box / cs of every slice and multicasts them to all
        // CTAs of the cluster; its own barrier expects the whole slice (the other parts arrive from the peers)
        const int part_rows = a.n_box / cs;
        for (long long tile = t0; tile < t_end; tile += tstep)
          for (int kc = 0; kc < nk; ++kc) {
            mbar_wait(&wr_empty[ws], wph ^ 1);
            mbar_arrive_expect_tx(&wr_full[ws], (uint32_t)w_slice);
            if (pair)  // my half of the rows the MMA uses: [crank, crank + 1) * 16 S
              tma_load_2d(s_w + ws * w_slice, &map_w, &wr_full[ws], 64 * kc, n0 + crank * 16 * S);
            else if (cs > 1)
              tma_load_2d_multicast(s_w + ws * w_slice + crank * part_rows * 128, &map_w, &wr_full[ws], 64 * kc,
                                    n0 + crank * part_rows, cmask);
            else
              tma_load_2d(s_w + ws * w_slice, &map_w, &wr_full[ws], 64 * kc, n0);
            if (++ws == a.wring) { ws = 0; wph ^= 1; }
          }
      }
    }
  } else if (pair && warp == E + kCvtWarps + 3) {
    // ================================================================ pair, peer CTA: tell the leader that this CTA's half of
    // each W slice has landed (the leader's MMA issuer can only wait on its own CTA's barriers)
    if (crank == 1 && elect_one()) {
      int ws = 0;
      uint32_t wph = 0;
      for (long long tile = t0; tile < t_end; tile += tstep)
        for (int kc = 0; kc < nk; ++kc) {
          mbar_wait(&wr_full[ws], wph);
          mbar_arrive_cluster(&wr_full2[ws], 0);
          if (++ws == a.wring) { ws = 0; wph ^= 1; }
        }
    }
  } else if (warp == E + kCvtWarps + 1) {
    // ================================================================ MMA issuer (pair: the leader CTA only)
    if (elect_one() && (!pair || crank == 0)) {
      constexpr uint64_t kDesc = umma_smem_desc(16, 1024, kLayoutSw128);
      const uint32_t idesc = idesc_bf16(pair ? 256 : 128, 32 * S);
      if (a.resident) mbar_wait(&w_full, 0);
      const uint32_t wb = smem_u32(s_w);
      const uint32_t idesc32 = umma_idesc_tf32(pair ? 256 : 128, 32 * S, 0, 0);
      const uint32_t rb = smem_u32(s_ring);
      int as_ = 0, ws = 0, slot = 0;
      uint32_t aph = 0, wph = 0, ph = 0;
      long long it = 0;
      for (long long tile = t0; tile < t_end; tile += tstep, ++it) {
        const int st = (int)(it & 1);
        if (pair) mbar_wait_cluster(&acc_empty[st], (((uint32_t)(it >> 1)) & 1) ^ 1);
        else mbar_wait(&acc_empty[st], (((uint32_t)(it >> 1)) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d = tmem + (uint32_t)st * a.stage_cols;
        for (int kc = 0; kc < nk; ++kc) {
          if (!a.resident) mbar_wait(&wr_full[ws], wph);
          if (pair) mbar_wait_cluster(&wr_full2[ws], wph);
          const uint32_t wk = wb + (a.resident ? kc : ws) * w_slice;
          if (a.prec == PREC_TF32) {
            // A: the fp32 chunk as TMA delivered it (K-major, SWIZZLE_128B); W slice: 32 tf32-rounded fp32 per row
            if (pair) mbar_wait_cluster(&cvt_full[slot], ph);  // both CTAs' chunks (forwarded by their A-producer warps)
            else mbar_wait(&raw_full[slot], ph);
            tc_fence_after();
            const uint32_t ab = rb + slot * kChunk;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t da = umma_desc_at(kDesc, ab + 32 * ks), db = umma_desc_at(kDesc, wk + 32 * ks);
              if (pair) umma_tf32_ss_pair(d, da, db, idesc32, (kc > 0 || ks > 0) ? 1u : 0u);
              else umma_tf32_ss_(d, da, db, idesc32, (kc > 0 || ks > 0) ? 1u : 0u);
            }
            if (pair) umma_commit_pair(&slot_empty[slot], 3);
            else umma_commit(&slot_empty[slot]);
            if (++slot == ring) { slot = 0; ph ^= 1; }
          } else if (a.ss) {
            // bf16x3 with the chunk split IN PLACE in shared memory: [hi(32) | lo(32)] bf16 per 128-byte row
            if (pair) mbar_wait_cluster(&cvt_full[slot], ph);
            else mbar_wait(&cvt_full[slot], ph);
            tc_fence_after();
            const uint32_t ab = rb + slot * kChunk;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint64_t a_hi = umma_desc_at(kDesc, ab + 32 * ks), a_lo = umma_desc_at(kDesc, ab + 64 + 32 * ks);
              const uint64_t w_hi = umma_desc_at(kDesc, wk + 32 * ks), w_lo = umma_desc_at(kDesc, wk + 64 + 32 * ks);
              const uint32_t acc0 = (kc > 0 || ks > 0) ? 1u : 0u;
              if (a.prec == PREC_BF16) {  // "bf16 operands, fp32 accumulate": the hi terms only
                if (pair) umma_bf16_ss_pair(d, a_hi, w_hi, idesc, acc0);
                else umma_bf16_ss(d, a_hi, w_hi, idesc, acc0);
              } else if (pair) {
                umma_bf16_ss_pair(d, a_lo, w_hi, idesc, acc0);
                umma_bf16_ss_pair(d, a_hi, w_lo, idesc, 1u);
                umma_bf16_ss_pair(d, a_hi, w_hi, idesc, 1u);
              } else {
                umma_bf16_ss(d, a_lo, w_hi, idesc, acc0);
                umma_bf16_ss(d, a_hi, w_lo, idesc, 1u);
                umma_bf16_ss(d, a_hi, w_hi, idesc, 1u);
              }
            }
            if (pair) umma_commit_pair(&slot_empty[slot], 3);
            else umma_commit(&slot_empty[slot]);
            if (++slot == ring) { slot = 0; ph ^= 1; }
          } else {
            mbar_wait(&a_full[as_], aph);
            tc_fence_after();
            const uint32_t at = tmem + kACol0 + 32 * as_;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint32_t a_hi = at + 8 * ks, a_lo = at + 16 + 8 * ks;
              const uint64_t w_hi = umma_desc_at(kDesc, wk + 32 * ks), w_lo = umma_desc_at(kDesc, wk + 64 + 32 * ks);
              if (a.prec == PREC_BF16X3) {
                umma_bf16_ts(d, a_lo, w_hi, idesc, (kc > 0 || ks > 0) ? 1u : 0u);
                umma_bf16_ts(d, a_hi, w_lo, idesc, 1u);
                umma_bf16_ts(d, a_hi, w_hi, idesc, 1u);
              } else {
                umma_bf16_ts(d, a_hi, w_hi, idesc, (kc > 0 || ks > 0) ? 1u : 0u);
              }
            }
            umma_commit(&a_empty[as_]);
            if (++as_ == kASlots) { as_ = 0; aph ^= 1; }
          }
          if (!a.resident) {
            if (pair) umma_commit_pair(&wr_empty[ws], 3);
            else if (cs > 1) umma_commit_multicast(&wr_empty[ws], cmask);
            else umma_commit(&wr_empty[ws]);
            if (++ws == a.wring) { ws = 0; wph ^= 1; }
          }
        }
        if (pair) umma_commit_pair(&acc_full[st], 3);
        else umma_commit(&acc_full[st]);
      }
    }
  } else if (warp >= E) {
    // ================================================================ converters: fp32 chunk (smem) -> hi / lo bf16
    // Two TEAMS of four warps (one per TMEM lane quadrant = warp % 4, the tcgen05.st restriction) take alternate chunks,
    // thread = one full 32-column row: two chunks are in flight, and a thread's row is its own (the in-place variant
    // needs no cross-warp synchronisation).  hi / lo go to a tensor-memory slot (A from TMEM) or back into the chunk
    // (ss: A stays in shared memory); in TF32 mode the converters only take the column sums, if those are wanted.
    const int cw = warp - E;
    const int q = cw & 3, team = cw >> 2;
    const int row = q * 32 + lane;
    const int sw = row & 7;
    const uint32_t ring_u32 = smem_u32(s_ring);
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    // (column sums ride with the plain / add modes only: the input-gradient GEMMs; compiled out of the GELU variants, whose
    // 16 epilogue warps leave 72 registers per thread)
    constexpr bool kColsumMode = (MODE == MODE_PLAIN || MODE == MODE_ADD);
    const bool do_colsum = kColsumMode && a.colsum != nullptr && chunk == 0;
    const bool convert = a.prec != PREC_TF32;  // TF32: the MMAs read the fp32 chunk directly
    const uint32_t cs_u32 = smem_u32(s_colsum);
    if (do_colsum) {
      for (int i = threadIdx.x - E * 32; i < nk * 32; i += kCvtWarps * 32) s_colsum[i] = 0.f;
      named_bar_sync(1, kCvtWarps * 32);
    }
    // (pair, TF32 without column sums: one warp per team only forwards the chunk's arrival to the leader; the others must
    // not wait on barriers whose progress does not depend on them -- they could fall a whole ring cycle behind)
    float ln_piv = 0.f, ln_s1 = 0.f, ln_s2 = 0.f;
    if (convert || do_colsum || (pair && q == 0)) {
      long long c = 0;  // running chunk index of this CTA
      for (long long tile = t0; tile < t_end; tile += tstep)
        for (int kc = 0; kc < nk; ++kc, ++c) {
          // The planner keeps the ring depth EVEN, so a ring slot always belongs to the same team and each of its warps
          // visits the slot's uses one after the other: a parity wait is then never more than one phase away from its
          // barrier.  (With an odd ring a slot alternates between the teams and a team can poll it one phase early, TMA
          // loads completing out of order; making every warp observe every chunk instead deadlocks, because the shared-
          // memory ring is released before the tensor-memory slot is awaited and a skipping warp can fall a whole ring
          // cycle behind.  Both were seen on the B200 at full size only.)
          if ((int)(c & 1) != team) continue;
          const int slot = (int)(c % ring);
          const uint32_t ph = (uint32_t)(c / ring) & 1;
          mbar_wait(&raw_full[slot], ph);
          if (pair && !convert) {
            if (q == 0 && lane == 0) mbar_arrive_cluster(&cvt_full[slot], 0);  // this CTA's chunk has landed
            if (!do_colsum) continue;
          }
          const uint32_t base = ring_u32 + slot * kChunk + row * 128;
          float4 v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = lds_f4(base + ((j ^ sw) << 4));
          if constexpr (MODE == MODE_LN_IN) {
            // row statistics of the raw input over this team's chunks of the tile (pivoted one-pass sums; the epilogue
            // merges the two teams' partial results): handed over through shared memory, four tiles deep -- the converters
            // never run that far ahead of the epilogue (rings of <= 10 chunks, >= 5 chunks per tile)
            const long long cb = c - kc;                          // running index of the tile's first chunk
            const int first_kc = ((int)(cb & 1) == team) ? 0 : 1;
            const int last_kc = ((int)((cb + nk - 1) & 1) == team) ? nk - 1 : nk - 2;
            if (kc == first_kc) { ln_piv = v[0].x; ln_s1 = 0.f; ln_s2 = 0.f; }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float d0 = v[j].x - ln_piv, d1 = v[j].y - ln_piv, d2 = v[j].z - ln_piv, d3 = v[j].w - ln_piv;
              ln_s1 += (d0 + d1) + (d2 + d3);
              ln_s2 = fmaf(d0, d0, ln_s2); ln_s2 = fmaf(d1, d1, ln_s2); ln_s2 = fmaf(d2, d2, ln_s2); ln_s2 = fmaf(d3, d3, ln_s2);
            }
            if (kc == last_kc) {
              const float inv_n = 1.0f / (float)(32 * ((last_kc - first_kc) / 2 + 1));
              reinterpret_cast<float2*>(s_colsum)[((int)((cb / nk) & 3) * 2 + team) * 128 + row] =
                  make_float2(ln_piv + ln_s1 * inv_n, fmaxf(ln_s2 - ln_s1 * ln_s1 * inv_n, 0.f));
            }
          }
          if (do_colsum) {
            // column sums of this warp's 32 rows x 32 columns: a reduce-scatter over the lanes (16 + 8 + 4 + 2 + 1
            // shuffles) leaves one column per lane, which adds it to the CTA's running sums
            const float* cc = reinterpret_cast<const float*>(v);
            const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2, b0 = lane & 1;
            float k16[16], k8[8], k4[4], k2[2];
#pragma unroll
            for (int i = 0; i < 16; ++i)
              k16[i] = (b4 ? cc[i + 16] : cc[i]) + __shfl_xor_sync(0xffffffffu, b4 ? cc[i] : cc[i + 16], 16);
#pragma unroll
            for (int i = 0; i < 8; ++i)
              k8[i] = (b3 ? k16[i + 8] : k16[i]) + __shfl_xor_sync(0xffffffffu, b3 ? k16[i] : k16[i + 8], 8);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              k4[i] = (b2 ? k8[i + 4] : k8[i]) + __shfl_xor_sync(0xffffffffu, b2 ? k8[i] : k8[i + 4], 4);
#pragma unroll
            for (int i = 0; i < 2; ++i)
              k2[i] = (b1 ? k4[i + 2] : k4[i]) + __shfl_xor_sync(0xffffffffu, b1 ? k4[i] : k4[i + 2], 2);
            const float k1 = (b0 ? k2[1] : k2[0]) + __shfl_xor_sync(0xffffffffu, b0 ? k2[0] : k2[1], 1);
            const int col = (b4 ? 16 : 0) + (b3 ? 8 : 0) + (b2 ? 4 : 0) + (b1 ? 2 : 0) + (b0 ? 1 : 0);
            red_add_shared_f32(cs_u32 + 4 * (32 * kc + col), k1);
          }
          if (!convert) {  // TF32 + column sums: the slot is released by the MMAs and by this team
            __syncwarp();
            if (lane == 0) mbar_arrive(&slot_empty[slot]);
            continue;
          }
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            split2(v[j].x, v[j].y, hi[2 * j], lo[2 * j]);
            split2(v[j].z, v[j].w, hi[2 * j + 1], lo[2 * j + 1]);
          }
          if (a.ss) {  // in place: hi -> 16-byte chunks 0-3 of my row, lo -> chunks 4-7 (all eight are in my registers)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              sts_u4(base + ((j ^ sw) << 4), make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]));
              sts_u4(base + (((4 + j) ^ sw) << 4), make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]));
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (pair) mbar_arrive_cluster(&cvt_full[slot], 0);  // the leader issues the MMAs of both CTAs
              else mbar_arrive(&cvt_full[slot]);
            }
            continue;
          }
          __syncwarp();  // every lane holds its row in registers: the shared-memory slot can be refilled
          if (lane == 0) mbar_arrive(&slot_empty[slot]);
          const int as_ = (int)(c % kASlots);
          mbar_wait(&a_empty[as_], (((uint32_t)(c / kASlots)) & 1) ^ 1);  // the MMAs that read this slot have completed
          tc_fence_after();
          const uint32_t at = tmem + lane_addr + kACol0 + 32 * as_;
          tmem_st16(at, hi);
          if (a.prec == PREC_BF16X3) tmem_st16(at + 16, lo);
          tmem_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&a_full[as_]);
        }
    }
    if (do_colsum) {
      named_bar_sync(1, kCvtWarps * 32);
      for (int i = threadIdx.x - E * 32; i < a.K; i += kCvtWarps * 32) atomicAdd(a.colsum + i, s_colsum[i]);
    }
  } else {
    // ================================================================ epilogue
    // Warp (g, q) owns rows [32q, 32q + 32) (its TMEM lane quadrant) of the slabs g, g + NG, ... of every tile, and rw
    // private 4 KB staging regions used round-robin: its own TMA stores (and aux loads), no cross-warp synchronisation.
    const int g = warp >> 2, q = warp & 3;
    const int sw = lane & 7;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int rw = a.rw;
    uint8_t* my = s_buf + warp * rw * kRegion;
    const uint32_t my_u32 = smem_u32(my);
    uint64_t* my_full = aux_full + warp * kMaxRw;
    const int spw = S > g ? (S - g + NG - 1) / NG : 0;            // slabs of this warp per tile
    const long long my_tiles = t_end > t0 ? (t_end - t0 + tstep - 1) / tstep : 0;
    const long long total = my_tiles * spw;                       // aux loads of this warp over the whole launch
    if constexpr (MODE == MODE_LN) {
      // ---------------------------------------------------------------- LayerNorm epilogue
      // An LN group is G / 32 consecutive slabs of the row a thread owns; its slabs are spread over the NG = 2 epilogue
      // groups like all others, so the two warps that share a lane quadrant (g = 0, 1; same rows, same lanes) exchange
      // partial statistics (mean, M2 of their slabs; merged with Chan's formula) through shared memory once per tile.
      // Pass 1 reads the accumulator for the statistics, pass 2 reads it again (TMEM reads are cheap), normalises and
      // stores; the accumulator stage is released after the second read.
      static_assert(NG == 2, "the LN epilogue pairs the two epilogue groups");
      const int G = a.G, SG = G >> 5;
      const int ngr = n_this / G;                    // LN groups of this column chunk: 1 or 2
      const int gpr = a.N / G, grp0 = n0 / G;        // groups per row of D; first group of this chunk
      const bool save_pre = a.save_pre != 0, has_aux = a.has_aux != 0;
      const int rq0 = save_pre ? 1 : 0, rq = rw - rq0;  // region 0: pre-norm stores; regions [rq0, rw): aux in / D2 out
      const int nst = save_pre ? 2 : 1;
      float2* exch = reinterpret_cast<float2*>(s_colsum);  // [tile parity][g][row of the tile][group]
      const float inv_g = 1.0f / (float)G;
      auto slabs_of = [&](int gg, int j) {           // slabs of LN group j that epilogue group gg owns
        const int lo = j * SG, hi = min((j + 1) * SG, S);
        int c = 0;
        for (int s2 = lo; s2 < hi; ++s2) c += ((s2 & 1) == gg);
        return c;
      };
      auto wait_read_n = [&](int n) {
        switch (n) {
          case 0: tma_store_wait_read<0>(); break;
          case 1: tma_store_wait_read<1>(); break;
          case 2: tma_store_wait_read<2>(); break;
          case 3: tma_store_wait_read<3>(); break;
          case 4: tma_store_wait_read<4>(); break;
          default: tma_store_wait_read<5>(); break;
        }
      };
      auto issue_aux_ln = [&](long long u) {         // slab use u -> its staging region (lane 0 only)
        const long long tile = t0 + (u / spw) * tstep;
        const int s = g + NG * (int)(u % spw);
        const int reg = rq0 + (int)(u % rq);
        mbar_arrive_expect_tx(&my_full[reg], kRegion);
        tma_load_2d(my + reg * kRegion, &map_aux, &my_full[reg], n0 + 32 * s, (int)(tile * kBM) + 32 * q);
      };
      if (has_aux && lane == 0)
        for (long long u = 0; u < rq && u < total; ++u) issue_aux_ln(u);
      auto merge = [](float& n, float& mean, float& m2, float nb, float mb, float m2b) {
        const float nt = n + nb;
        if (nt > 0.f) {
          const float delta = mb - mean, f = nb / nt;
          mean += delta * f;
          m2 += m2b + delta * delta * n * f;
        }
        n = nt;
      };
      // The refill of a shortcut region (wait until its store has read it, request the sub-slab rq uses ahead) is not done
      // right after the store -- lane 0 would sit in that wait while the warp still owes the accumulator stage its last
      // reads -- but at the next point where the warp has nothing urgent: after the next accumulator read, or at the top
      // of the next tile.
      long long refill = -1;
      auto flush_refill = [&]() {
        if (refill >= 0) {
          tma_store_wait_read<0>();
          issue_aux_ln(refill);
          refill = -1;
        }
      };
      long long u = 0, it = 0;
      for (long long tile = t0; tile < t_end; tile += tstep, ++it) {
        const int as = (int)(it & 1);
        if (lane == 0) flush_refill();
        mbar_wait(&acc_full[as], ((uint32_t)(it >> 1)) & 1);
        tc_fence_after();
        const long long grow = tile * kBM + q * 32 + lane;
        const uint32_t acc_addr = tmem + lane_addr + (uint32_t)as * a.stage_cols;
        // ---- pass 1: statistics of my slabs, per LN group
        float n_0 = 0.f, m_0 = 0.f, q_0 = 0.f, n_1 = 0.f, m_1 = 0.f, q_1 = 0.f;
        for (int s = g; s < S; s += NG) {
          uint32_t acc[32];
          tmem_ld32(acc_addr + 32 * s, acc);
          tmem_wait_ld();
          const int jc = n0 + 32 * s;
          float piv = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a.bias) bv = __ldg(reinterpret_cast<const float4*>(a.bias + jc + 4 * c));
            const float v0 = __uint_as_float(acc[4 * c + 0]) + bv.x, v1 = __uint_as_float(acc[4 * c + 1]) + bv.y;
            const float v2 = __uint_as_float(acc[4 * c + 2]) + bv.z, v3 = __uint_as_float(acc[4 * c + 3]) + bv.w;
            if (c == 0) piv = v0;   // pivot inside the data range: the one-pass variance below does not cancel
            const float d0 = v0 - piv, d1 = v1 - piv, d2 = v2 - piv, d3 = v3 - piv;
            s1 += (d0 + d1) + (d2 + d3);
            s2 = fmaf(d0, d0, s2); s2 = fmaf(d1, d1, s2); s2 = fmaf(d2, d2, s2); s2 = fmaf(d3, d3, s2);
          }
          const float mb = piv + s1 * (1.0f / 32.0f), m2b = fmaxf(s2 - s1 * s1 * (1.0f / 32.0f), 0.f);
          if (s / SG == 0) merge(n_0, m_0, q_0, 32.f, mb, m2b);
          else merge(n_1, m_1, q_1, 32.f, mb, m2b);
        }
        float2* ex = exch + (size_t)(it & 1) * (2 * 128 * 2);
        ex[(g * 128 + q * 32 + lane) * 2 + 0] = make_float2(m_0, q_0);
        ex[(g * 128 + q * 32 + lane) * 2 + 1] = make_float2(m_1, q_1);
        named_bar_sync(2 + q, 64);  // the two warps of this lane quadrant (double-buffered: one barrier per tile)
        float mean0, rstd0, mean1 = 0.f, rstd1 = 0.f;
        {
          const float2 o0 = ex[((g ^ 1) * 128 + q * 32 + lane) * 2 + 0], o1 = ex[((g ^ 1) * 128 + q * 32 + lane) * 2 + 1];
          // merge in the fixed order (group 0, group 1) so that both warps compute bit-identical statistics
          float na = 32.f * slabs_of(0, 0), ma = g == 0 ? m_0 : o0.x, qa = g == 0 ? q_0 : o0.y;
          merge(na, ma, qa, 32.f * slabs_of(1, 0), g == 0 ? o0.x : m_0, g == 0 ? o0.y : q_0);
          mean0 = ma;
          rstd0 = rsqrtf(qa * inv_g + a.eps);
          if (ngr > 1) {
            float nb = 32.f * slabs_of(0, 1), mb = g == 0 ? m_1 : o1.x, qb = g == 0 ? q_1 : o1.y;
            merge(nb, mb, qb, 32.f * slabs_of(1, 1), g == 0 ? o1.x : m_1, g == 0 ? o1.y : q_1);
            mean1 = mb;
            rstd1 = rsqrtf(qb * inv_g + a.eps);
          }
        }
        if (g == 0 && a.mean_out && grow < a.T) {
          a.mean_out[grow * gpr + grp0] = mean0;
          a.rstd_out[grow * gpr + grp0] = rstd0;
          if (ngr > 1) {
            a.mean_out[grow * gpr + grp0 + 1] = mean1;
            a.rstd_out[grow * gpr + grp0 + 1] = rstd1;
          }
        }
        // ---- pass 2: normalise and store
        if (spw == 0) {
          tc_fence_before();
          mbar_arrive(&acc_empty[as]);
        }
        for (int s = g; s < S; s += NG, ++u) {
          uint32_t acc[32];
          tmem_ld32(acc_addr + 32 * s, acc);
          tmem_wait_ld();
          if (s + NG >= S) {  // this thread's last read of the accumulator stage
            tc_fence_before();
            mbar_arrive(&acc_empty[as]);
          }
          if (lane == 0) flush_refill();
          const int jc = n0 + 32 * s;
          const bool second = (s / SG) != 0;
          const float mean = second ? mean1 : mean0, rstd = second ? rstd1 : rstd0;
          const int gc = 32 * s - (second ? G : 0);  // column inside the LN group
          if (save_pre) {
            if (lane == 0) tma_store_wait_read<1>();  // my previous pre-norm store has read region 0
            __syncwarp();
            const uint32_t brow = my_u32 + lane * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
              if (a.bias) bv = __ldg(reinterpret_cast<const float4*>(a.bias + jc + 4 * c));
              sts_f4(brow + ((c ^ sw) << 4),
                     make_float4(__uint_as_float(acc[4 * c + 0]) + bv.x, __uint_as_float(acc[4 * c + 1]) + bv.y,
                                 __uint_as_float(acc[4 * c + 2]) + bv.z, __uint_as_float(acc[4 * c + 3]) + bv.w));
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&map_d, my, jc, (int)(tile * kBM) + 32 * q);
              tma_store_commit();
            }
          }
          const int reg = rq0 + (int)(u % rq);
          if (has_aux) {
            mbar_wait(&my_full[reg], ((uint32_t)(u / rq)) & 1);  // the shortcut's sub-slab has landed
          } else {
            // the D2 store that used this region rq uses ago must have read it
            if (lane == 0) wait_read_n((rq - 1) * nst + (nst - 1));
            __syncwarp();
          }
          const uint32_t brow = my_u32 + reg * kRegion + lane * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint32_t p = brow + ((c ^ sw) << 4);
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a.bias) bv = __ldg(reinterpret_cast<const float4*>(a.bias + jc + 4 * c));
            const float4 gm = __ldg(reinterpret_cast<const float4*>(a.gamma + gc + 4 * c));
            const float4 bt = __ldg(reinterpret_cast<const float4*>(a.beta + gc + 4 * c));
            float4 o;
            o.x = fmaf((__uint_as_float(acc[4 * c + 0]) + bv.x - mean) * rstd, gm.x, bt.x);
            o.y = fmaf((__uint_as_float(acc[4 * c + 1]) + bv.y - mean) * rstd, gm.y, bt.y);
            o.z = fmaf((__uint_as_float(acc[4 * c + 2]) + bv.z - mean) * rstd, gm.z, bt.z);
            o.w = fmaf((__uint_as_float(acc[4 * c + 3]) + bv.w - mean) * rstd, gm.w, bt.w);
            if (has_aux) {
              const float4 x = lds_f4(p);
              o.x += x.x; o.y += x.y; o.z += x.z; o.w += x.w;
            }
            sts_f4(p, o);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&map_d2, my + reg * kRegion, jc, (int)(tile * kBM) + 32 * q);
            tma_store_commit();
            if (has_aux && u + rq < total) refill = u + rq;  // this region's next shortcut sub-slab (see flush_refill)
          }
        }
      }
      if (lane == 0) tma_store_wait<0>();
    } else {
    auto issue_aux = [&](long long u) {                           // slab use u -> its staging region (lane 0 only)
      const long long tile = t0 + (u / spw) * tstep;
      const int s = g + NG * (int)(u % spw);
      const int reg = (int)(u % rw);
      mbar_arrive_expect_tx(&my_full[reg], kRegion);
      tma_load_2d(my + reg * kRegion, &map_aux, &my_full[reg], n0 + 32 * s, (int)(tile * kBM) + 32 * q);
    };
    if (kAux && lane == 0)
      for (long long u = 0; u < rw && u < total; ++u) issue_aux(u);
    long long cnt = 0;  // staging-region uses so far
    long long it = 0;
    // MODE_GELU_GRAD_C: this row's 32 g' values (FP16) of the NEXT slab this warp will process are requested as soon as the
    // current ones have been consumed (same registers): their latency hides behind the store and the next accumulator wait
    uint4 gq[MODE == MODE_GELU_GRAD_C ? 4 : 1];
    auto fetch_gq = [&](long long tile, int s) {
      const long long row = tile * kBM + q * 32 + lane;
      const uint4* src = reinterpret_cast<const uint4*>(a.gp16 + row * a.N + n0 + 32 * s);
#pragma unroll
      for (int i = 0; i < (MODE == MODE_GELU_GRAD_C ? 4 : 1); ++i)
        gq[i] = (tile < t_end && s < S && row < a.T) ? __ldcs(src + i) : make_uint4(0u, 0u, 0u, 0u);
    };
    if constexpr (MODE == MODE_GELU_GRAD_C) fetch_gq(t0, g);
    for (long long tile = t0; tile < t_end; tile += tstep, ++it) {
      const int as = (int)(it & 1);
      mbar_wait(&acc_full[as], ((uint32_t)(it >> 1)) & 1);
      tc_fence_after();
      const long long grow = tile * kBM + q * 32 + lane;
      uint32_t dkey = 0;
      if (MODE == MODE_GELU || MODE == MODE_GELU_GRAD || MODE == MODE_GELU_C)
        dkey = a.drop_thresh ? hs::drop_row_key(a.seed, grow) : 0u;
      float ln_mean = 0.f, ln_rstd = 0.f;
      if constexpr (MODE == MODE_LN_IN) {  // the converters' partial row statistics of this tile -> mean, 1 / std
        const float2* st = reinterpret_cast<const float2*>(s_colsum) + (size_t)(it & 3) * 256 + q * 32 + lane;
        const float2 p0 = st[0], p1 = st[128];
        const long long cb = it * nk;
        const int n_even = (nk + 1) / 2, n_odd = nk / 2;              // chunks at even / odd offsets inside the tile
        const float na = 32.f * ((cb & 1) ? n_odd : n_even), nb = 32.f * ((cb & 1) ? n_even : n_odd);  // team 0, team 1
        const float nt = na + nb, delta = p1.x - p0.x, f = nb / nt;
        ln_mean = p0.x + delta * f;
        ln_rstd = rsqrtf((p0.y + p1.y + delta * delta * na * f) / nt + a.eps);
        if (g == 0 && chunk == 0 && a.mean_out && grow < a.T) {
          a.mean_out[grow] = ln_mean;
          a.rstd_out[grow] = ln_rstd;
        }
      }
      auto release_stage = [&]() {  // this thread's part of the accumulator stage is in registers (or unused)
        tc_fence_before();
        if (pair) {
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&acc_empty[as], 0);
        } else {
          mbar_arrive(&acc_empty[as]);
        }
      };
      if (spw == 0) release_stage();  // more groups than slabs: stay in step with the accumulator stages
      for (int s = g; s < S; s += NG) {
        const int jc = n0 + 32 * s;
        uint32_t gpk[MODE == MODE_GELU_C ? 16 : 1];  // g' of this row's 32 columns, packed FP16 pairs
        uint32_t acc[32];
        tmem_ld32(tmem + lane_addr + (uint32_t)as * a.stage_cols + 32 * s, acc);
        tmem_wait_ld();
        if (s + NG >= S) release_stage();  // this thread's last slab of the tile is in registers
#pragma unroll
        for (int st = 0; st < kStores; ++st, ++cnt) {
          const int reg = (int)(cnt % rw);
          if (kAux) {
            mbar_wait(&my_full[reg], ((uint32_t)(cnt / rw)) & 1);  // the aux sub-slab has landed
          } else {
            // the store that used this region rw uses ago must have read it
            if (lane == 0) {
              if (rw == 1) tma_store_wait_read<0>();
              else if (rw == 2) tma_store_wait_read<1>();
              else tma_store_wait_read<2>();
            }
            __syncwarp();
          }
          const uint32_t brow = my_u32 + reg * kRegion + lane * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint32_t p = brow + ((c ^ sw) << 4);
            float4 o = make_float4(__uint_as_float(acc[4 * c + 0]), __uint_as_float(acc[4 * c + 1]),
                                   __uint_as_float(acc[4 * c + 2]), __uint_as_float(acc[4 * c + 3]));
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a.bias && !(MODE == MODE_GELU && st == 0) && jc + 4 * c < a.N)
              bv = __ldg(reinterpret_cast<const float4*>(a.bias + jc + 4 * c));
            if (MODE == MODE_PLAIN) {
              o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
            } else if (MODE == MODE_LN_IN) {
              float4 sv = make_float4(0.f, 0.f, 0.f, 0.f);
              if (jc + 4 * c < a.N) sv = __ldg(reinterpret_cast<const float4*>(a.wsum + jc + 4 * c));
              o.x = fmaf(fmaf(-ln_mean, sv.x, o.x), ln_rstd, bv.x); o.y = fmaf(fmaf(-ln_mean, sv.y, o.y), ln_rstd, bv.y);
              o.z = fmaf(fmaf(-ln_mean, sv.z, o.z), ln_rstd, bv.z); o.w = fmaf(fmaf(-ln_mean, sv.w, o.w), ln_rstd, bv.w);
            } else if (MODE == MODE_ADD) {
              const float4 x = lds_f4(p);
              o.x += bv.x + x.x; o.y += bv.y + x.y; o.z += bv.z + x.z; o.w += bv.w + x.w;
            } else if (MODE == MODE_GELU) {
              if (st == 1) {
                o.x = hs::gelu_fast(o.x + bv.x); o.y = hs::gelu_fast(o.y + bv.y);
                o.z = hs::gelu_fast(o.z + bv.z); o.w = hs::gelu_fast(o.w + bv.w);
                if (a.drop_thresh) {
                  o.x = hs::drop_keep_elem(dkey, jc + 4 * c + 0, a.drop_thresh) ? o.x * a.drop_scale : 0.f;
                  o.y = hs::drop_keep_elem(dkey, jc + 4 * c + 1, a.drop_thresh) ? o.y * a.drop_scale : 0.f;
                  o.z = hs::drop_keep_elem(dkey, jc + 4 * c + 2, a.drop_thresh) ? o.z * a.drop_scale : 0.f;
                  o.w = hs::drop_keep_elem(dkey, jc + 4 * c + 3, a.drop_thresh) ? o.w * a.drop_scale : 0.f;
                }
              }
            } else if (MODE == MODE_GELU_C) {
              if (st == 0) {
                // h and g' from ONE evaluation of the erf terms: g' goes out now (packed FP16), h replaces the
                // accumulator values in registers and goes out in the second pass
                const float u4[4] = {o.x + bv.x, o.y + bv.y, o.z + bv.z, o.w + bv.w};
                float g4[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const hs::GeluTerms t = hs::gelu_terms(u4[e]);
                  float hv = u4[e] * t.cdf;
                  g4[e] = fmaf(u4[e] * 0.39894228040143267794f, t.e, t.cdf);
                  if (a.drop_thresh) {
                    const bool keep = hs::drop_keep_elem(dkey, jc + 4 * c + e, a.drop_thresh);
                    hv = keep ? hv * a.drop_scale : 0.f;
                    g4[e] = keep ? g4[e] * a.drop_scale : 0.f;
                  }
                  acc[4 * c + e] = __float_as_uint(hv);
                }
                asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(gpk[2 * c + 0]) : "f"(g4[1]), "f"(g4[0]));
                asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(gpk[2 * c + 1]) : "f"(g4[3]), "f"(g4[2]));
                continue;  // (the FP16 rows are written after the loop)
              }
            } else if (MODE == MODE_GELU_GRAD_C) {
              const uint32_t w0 = c & 1 ? gq[c >> 1].z : gq[c >> 1].x, w1 = c & 1 ? gq[c >> 1].w : gq[c >> 1].y;
              const float2 g01 = __half22float2(*reinterpret_cast<const __half2*>(&w0));
              const float2 g23 = __half22float2(*reinterpret_cast<const __half2*>(&w1));
              o.x *= g01.x; o.y *= g01.y; o.z *= g23.x; o.w *= g23.y;
            } else {  // MODE_GELU_GRAD
              const float4 z = lds_f4(p);
              o.x *= hs::gelu_grad_fast(z.x + bv.x); o.y *= hs::gelu_grad_fast(z.y + bv.y);
              o.z *= hs::gelu_grad_fast(z.z + bv.z); o.w *= hs::gelu_grad_fast(z.w + bv.w);
              if (a.drop_thresh) {
                o.x = hs::drop_keep_elem(dkey, jc + 4 * c + 0, a.drop_thresh) ? o.x * a.drop_scale : 0.f;
                o.y = hs::drop_keep_elem(dkey, jc + 4 * c + 1, a.drop_thresh) ? o.y * a.drop_scale : 0.f;
                o.z = hs::drop_keep_elem(dkey, jc + 4 * c + 2, a.drop_thresh) ? o.z * a.drop_scale : 0.f;
                o.w = hs::drop_keep_elem(dkey, jc + 4 * c + 3, a.drop_thresh) ? o.w * a.drop_scale : 0.f;
              }
            }
            sts_f4(p, o);
          }
          if constexpr (MODE == MODE_GELU_GRAD_C) {  // (gq consumed) -> the next slab of this warp: same tile, or the next tile's first
            if (s + NG < S) fetch_gq(tile, s + NG);
            else fetch_gq(tile + tstep, g);
          }
          if constexpr (MODE == MODE_GELU_C) {
            if (st == 0) {  // the g' sub-slab: 32 rows of 64 bytes, SWIZZLE_64B (16-byte chunk ^= (row >> 1) & 3)
              const uint32_t r64 = my_u32 + reg * kRegion + lane * 64;
#pragma unroll
              for (int i = 0; i < 4; ++i)
                sts_u4(r64 + ((i ^ ((lane >> 1) & 3)) << 4),
                       make_uint4(gpk[4 * i + 0], gpk[4 * i + 1], gpk[4 * i + 2], gpk[4 * i + 3]));
            }
          }
          fence_proxy_async_smem();
          __syncwarp();  // all 32 rows of the sub-slab are in the staging region
          if (lane == 0) {
            tma_store_2d(((MODE == MODE_GELU && st == 1) || (MODE == MODE_GELU_C && st == 0)) ? &map_d2 : &map_d,
                         my + reg * kRegion, jc,
                         (int)(tile * kBM) + 32 * q);
            tma_store_commit();
            if (kAux && cnt + rw < total) {  // refill this region with the aux sub-slab of the use rw ahead
              tma_store_wait_read<0>();
              issue_aux(cnt + rw);
            }
          }
        }
      }
    }
    if (lane == 0) tma_store_wait<0>();
    }  // (plain / add / GELU epilogues)
  }
  tc_fence_before();
  __syncthreads();
  if (cs > 1) cluster_sync_all();  // no CTA leaves while a peer may still multicast into it
  if (warp == E + kCvtWarps + 1) {
    if (pair) tmem_dealloc_pair(tmem, 512);
    else tmem_dealloc(tmem, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// W (rows x cols fp32, row stride ld; element (r, c) = w[r * ld + c], or w[c * ld + r] when transposed) ->
// out (rows, ceil(cols / 32), 64) bf16 = [hi(32) | lo(32)] per 32-wide chunk of the contraction axis (zero padded).
// format 1: out (rows, ceil(cols / 32), 32) fp32 rounded to the nearest TF32 (the single-MMA TF32 mode; same bytes per row).
__device__ __forceinline__ void weight_split_tile(float (*tile)[33], const float* __restrict__ w, uint16_t* __restrict__ out,
                                                  int rows, int cols, int ld, int transposed, int format, int bx, int by,
                                                  int nbx) {
  const int r0 = by * 32, c0 = bx * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    float v = 0.f;
    if (!transposed) {
      if (r0 + i < rows && c0 + tx < cols) v = w[(long long)(r0 + i) * ld + c0 + tx];
      tile[i][tx] = v;
    } else {  // coalesced along the source's fast axis (= output rows), transposed through shared memory
      if (c0 + i < cols && r0 + tx < rows) v = w[(long long)(c0 + i) * ld + r0 + tx];
      tile[tx][i] = v;
    }
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    if (r0 + i >= rows) continue;
    const float x = tile[i][tx];
    if (format == 1) {
      uint32_t t;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(x));
      reinterpret_cast<uint32_t*>(out)[((long long)(r0 + i) * nbx + bx) * 32 + tx] = t;
      continue;
    }
    const uint32_t hb = cvt_bf16x2(0.f, x) & 0xffffu;
    const float h = __uint_as_float(hb << 16);
    const uint32_t lb = cvt_bf16x2(0.f, x - h) & 0xffffu;
    uint16_t* o = out + ((long long)(r0 + i) * nbx + bx) * 64;
    o[tx] = (uint16_t)hb;
    o[32 + tx] = (uint16_t)lb;
  }
}

__global__ void weight_split_kernel(const float* __restrict__ w, uint16_t* __restrict__ out, int rows, int cols, int ld,
                                    int transposed, int format) {
  __shared__ float tile[32][33];
  weight_split_tile(tile, w, out, rows, cols, ld, transposed, format, blockIdx.x, blockIdx.y, gridDim.x);
}

// All weight operands of a model in ONE launch (hs_weight_split_batch): block b finds its matrix by binary search over the
// running tile counts.  Same per-tile code as above.
struct SplitDesc {   // mirrored by heal_swin_b200/ops.py:_SplitDesc
  const float* w;
  uint16_t* out;
  int rows, cols, ld, transposed, format, tiles_x, tile0, pad;
};
__global__ void weight_split_batch_kernel(const SplitDesc* __restrict__ descs, int n) {
  __shared__ float tile[32][33];
  const int b = blockIdx.x;
  int lo = 0, hi = n - 1;
  while (lo < hi) {  // last descriptor with tile0 <= b
    const int mid = (lo + hi + 1) >> 1;
    if (descs[mid].tile0 <= b) lo = mid;
    else hi = mid - 1;
  }
  const SplitDesc d = descs[lo];
  const int t = b - d.tile0;
  weight_split_tile(tile, d.w, d.out, d.rows, d.cols, d.ld, d.transposed, d.format, t % d.tiles_x, t / d.tiles_x, d.tiles_x);
}

int make_map_bf16(CUtensorMap* m, const uint16_t* base, long long rows, long long cols, int box_cols, int box_rows) {
  hs::tc::EncodeTiledFn enc = hs::tc::encode_fn();
  if (!enc) return hs::fail(HS_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<uint16_t*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return hs::fail(HS_ERR_CUDA, "cuTensorMapEncodeTiled (bf16) failed with CUresult %d", (int)r);
  return HS_OK;
}

int make_map_f16(CUtensorMap* m, const uint16_t* base, long long rows, long long cols, int box_cols, int box_rows) {
  hs::tc::EncodeTiledFn enc = hs::tc::encode_fn();
  if (!enc) return hs::fail(HS_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<uint16_t*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return hs::fail(HS_ERR_CUDA, "cuTensorMapEncodeTiled (f16) failed with CUresult %d", (int)r);
  return HS_OK;
}

constexpr int kSmemAvail = 232448 - 1024 - 1024;  // dynamic shared memory minus alignment slack and the static barriers

// column chunking + shared-memory plan: the widest balanced column chunk (<= 192) that leaves an A ring of >= 4 slots
int plan(G3Args& a, int mode, int E) {
  a.tiles = (a.T + kBM - 1) / kBM;
  const int nk = (a.K + 31) / 32;
  const bool aux = (mode == MODE_ADD || mode == MODE_GELU_GRAD);
  const bool ln = (mode == MODE_LN);
  // (LN: one region for the pre-norm stores + two for shortcut-in / result-out, so that a shortcut sub-slab is requested
  // two slab uses -- about one tile -- before it is needed)
  const int rw_min = ln ? (a.save_pre ? 3 : 2) : (aux ? 2 : 1), rw_max = (aux || ln) ? 3 : 2;
  // tensor-bound launches (streamed W, long contraction) keep A in shared memory and use 256-column stages: fewer, wider
  // MMAs (ncu: 69 % tensor-pipe active against 59 % with 192-column chunks); HBM-bound launches take A through TMEM
  // (PREC_BF16, one MMA per product, takes the same route: what binds it first is not the tensor pipe but the L2 ->
  // shared-memory stream of W slices, which the pair's M = 256 tiles halve per token.  BASELINE configs[3], C = 128, one
  // training step: 191.3 ms with the TMEM-A route everywhere, 183.8 / 178.6 / 178.2 ms with the threshold at 390 / 200 / 130)
  long long ridge_k = a.prec == PREC_BF16 ? 160 : 130;
  if (const char* e = getenv("HEALSWIN_GEMM3_BF16_RIDGE")) {  // experiments only
    if (a.prec == PREC_BF16 && atoll(e) > 0) ridge_k = atoll(e);
  }
  const long long ai = (long long)a.N * a.K, ridge = ridge_k * (a.N + a.K);
  a.ss = (a.prec == PREC_TF32 || (ai > ridge && a.K >= 256)) ? 1 : 0;
  if (const char* e = getenv("HEALSWIN_GEMM3_SS")) a.ss = (a.prec == PREC_TF32) ? 1 : (atoi(e) != 0);
  if (ln) a.ss = 0;  // the LN epilogue works on 192-column stages whose chunks hold whole LN groups
  a.stage_cols = a.ss ? 256 : 192;
  // CTA pairs (cta_group::2, M = 256): each CTA of a pair stages only half of every W slice, which halves the B-operand
  // traffic of its shared memory -- the co-limiter of the ss modes (per K = 16 step and CTA: 76 KB of shared-memory
  // traffic = 594 cycles against 384 cycles of MMA time alone; 56 KB = 437 cycles in a pair)
  // Measured at stages 2-3 (profiles/r2p_gemm3_pair.log): bf16x3 5-8 % faster (tensor pipe 67 -> 75 % active under ncu),
  // the single-MMA TF32 mode 5-6 % with a W ring of four half-slices (none with three: the round trip slice released ->
  // reloaded -> forwarded by the peer is longer than in one CTA).  HEALSWIN_GEMM3_PAIR=0 / 1 forces it off / on.
  a.pair = (a.ss && a.tiles >= 2) ? 1 : 0;
  if (const char* e = getenv("HEALSWIN_GEMM3_PAIR")) a.pair = (atoi(e) != 0 && a.ss && a.tiles >= 2) ? 1 : 0;
  int first = (a.N + a.stage_cols - 1) / a.stage_cols;  // number of chunks
  first = (((a.N + first - 1) / first) + 15) / 16 * 16;  // equal chunks, 16-column granularity
  if (const char* e = getenv("HEALSWIN_GEMM3_NTILE")) {  // experiments only
    const int v = atoi(e);
    if (v >= 32 && v <= a.stage_cols && v % 16 == 0 && v < first) first = v;
  }
  if (ln) {  // one or two whole LN groups per chunk
    first = (192 / a.G >= 2 ? 2 : 1) * a.G;
    if (first > a.N) first = a.N;
  }
  const int cand[7] = {first, 192, 160, 128, 96, 64, 32};
  int best_ring = 0;
  for (int ci = 0; ci < (ln ? 1 : 7); ++ci) {
    const int stride = cand[ci];
    if (stride > first || (ci > 0 && stride == first)) continue;
    const int box = (stride + 31) / 32 * 32;
    const int w_slice = (a.pair ? box / 2 : box) * 128;
    const long long staging_min =
        (long long)E * rw_min * kRegion + (a.colsum ? nk * 128 : 0) + ((ln || mode == MODE_LN_IN) ? kLnExch : 0);
    const int resident = (!a.pair && (long long)nk * w_slice + 4 * kChunk + staging_min <= kSmemAvail) ? 1 : 0;
    int wring = a.pair ? (nk < 4 ? nk : 4) : (nk < 3 ? nk : 3);  // (a pair's slices are half the size)
    if (const char* e = getenv("HEALSWIN_GEMM3_WRING")) {  // experiments only
      const int v = atoi(e);
      if (v >= 2 && v <= kMaxWRing && v <= nk) wring = v;
    }
    long long w_bytes = (long long)(resident ? nk : wring) * w_slice;
    long long left = kSmemAvail - w_bytes - staging_min;
    int ring = left > 0 ? (int)(left / kChunk) : 0;
    if (ln && !resident && ring < 4 && wring == 3) {  // the LN staging is large: two W slices in flight are enough (L2 hits)
      wring = 2;
      w_bytes = (long long)wring * w_slice;
      left = kSmemAvail - w_bytes - staging_min;
      ring = left > 0 ? (int)(left / kChunk) : 0;
    }
    if (ring > kMaxRing) ring = kMaxRing;
    ring &= ~1;  // even: chunk c and chunk c + ring are converted by the same team
    if (ring <= best_ring) continue;
    best_ring = ring;
    a.n_stride = stride;
    a.n_box = box;
    a.n_chunks = (a.N + stride - 1) / stride;
    a.resident = resident;
    a.wring = wring;
    a.ring = ring;
    left -= (long long)ring * kChunk;
    a.rw = rw_min + (int)(left / ((long long)E * kRegion));  // (staging_min already holds rw_min regions + the column sums)
    // Streamed W slices can be loaded once per cluster of 2 / 4 CTAs and multicast (HEALSWIN_GEMM3_CLUSTER): verified
    // correct on the B200 and measured to change nothing (profiles/r2q_cluster_exp.log) -- the shapes that stream W are
    // not bound by L2 -> SM traffic but by the shared-memory port (see a.pair) -- so multicast stays off.
    a.cluster = 1;
    if (a.rw > rw_max) a.rw = rw_max;
    if (ring >= 4) break;
  }
  if (best_ring < 2) return hs::fail(HS_ERR_UNSUPPORTED, "hs_gemm3: no shared-memory plan for N=%d K=%d", a.N, a.K);
  if (const char* e = getenv("HEALSWIN_GEMM3_CLUSTER")) {  // experiments only
    const int v = atoi(e);
    if ((v == 1 || v == 2 || v == 4) && !a.resident && a.n_box % (16 * v) == 0) a.cluster = v;
    if (v == 1) a.cluster = 1;
  }
  if (a.pair) a.cluster = 2;
  return HS_OK;
}

template <int E>
size_t smem_bytes(const G3Args& a) {
  const int nk = (a.K + 31) / 32, w_slice = (a.pair ? a.n_box / 2 : a.n_box) * 128;
  return (size_t)(a.resident ? nk : a.wring) * w_slice + (size_t)a.ring * kChunk + (size_t)E * a.rw * kRegion +
         (a.colsum ? (size_t)nk * 128 : 0) + ((a.G > 0 || a.wsum) ? (size_t)kLnExch : 0) + 1024;
}

struct Maps {
  CUtensorMap a, w, aux, d, d2;
};

template <int E, int MODE, bool PAIR>
int launch_kernel(const Maps& m, const G3Args& a, size_t smem, int per_chunk, cudaStream_t stream) {
  // the dynamic shared-memory limit is raised once per (instantiation, device) to the largest plan; an immutable cache
  // (the value never changes afterwards), so that no attribute call happens inside a CUDA-graph capture
  static bool raised[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !raised[dev]) {
    HS_CUDA(cudaFuncSetAttribute(gemm3_kernel<E, MODE, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemAvail + 1024));
    if (dev >= 0 && dev < 64) raised[dev] = true;
  }
  const int cs = a.cluster;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(a.n_chunks * per_chunk));
  cfg.blockDim = dim3(block_threads<E, PAIR>());
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = cs > 1 ? 1 : 0;
  HS_CUDA(cudaLaunchKernelEx(&cfg, gemm3_kernel<E, MODE, PAIR>, m.a, m.w, m.aux, m.d, m.d2, a));
  HS_LAUNCH_CHECK();
  return HS_OK;
}

template <int E, int MODE>
int launch(const float* a_dev, const uint16_t* w_dev, const float* aux_dev, float* d_dev, float* d2_dev, G3Args& a,
           cudaStream_t stream) {
  int rc;
  if ((rc = plan(a, MODE, E))) return rc;
  Maps m;
  if ((rc = hs::tc::make_map(&m.a, a_dev, a.T, a.K, CU_TENSOR_MAP_SWIZZLE_128B, 32, kBM))) return rc;
  if ((rc = hs::tc::make_map(&m.d, d_dev, a.T, a.N, CU_TENSOR_MAP_SWIZZLE_128B, 32, 32))) return rc;
  m.aux = m.d;
  m.d2 = m.d;
  if (aux_dev && (rc = hs::tc::make_map(&m.aux, aux_dev, a.T, a.N, CU_TENSOR_MAP_SWIZZLE_128B, 32, 32))) return rc;
  if (d2_dev && (rc = hs::tc::make_map(&m.d2, d2_dev, a.T, a.N, CU_TENSOR_MAP_SWIZZLE_128B, 32, 32))) return rc;
  if (MODE == MODE_GELU_C && (rc = make_map_f16(&m.d2, a.gp16, a.T, a.N, 32, 32))) return rc;  // the FP16 g' tensor
  const size_t smem = smem_bytes<E>(a);
  const int cs = a.cluster;
  int per_chunk = hs::tc::sm_count() / a.n_chunks;
  if (per_chunk > a.tiles) per_chunk = (int)a.tiles;
  per_chunk = per_chunk / cs * cs;
  if (per_chunk < cs) per_chunk = cs;
  if ((rc = make_map_bf16(&m.w, w_dev, a.N, 2ll * ((a.K + 31) / 32 * 32), 64, a.n_box / cs))) return rc;
  if constexpr (MODE == MODE_LN) return launch_kernel<E, MODE, false>(m, a, smem, per_chunk, stream);  // (never a pair)
  else
    return a.pair ? launch_kernel<E, MODE, true>(m, a, smem, per_chunk, stream)
                  : launch_kernel<E, MODE, false>(m, a, smem, per_chunk, stream);
}

}  // namespace

extern "C" {

int hs_weight_split(const float* w, int rows, int cols, int ld, int transposed, int format, uint16_t* out, void* stream) {
  HS_REQUIRE(w && out && rows > 0 && cols > 0 && ld > 0, "hs_weight_split: bad arguments");
  HS_REQUIRE(format == 0 || format == 1, "hs_weight_split: unknown format %d", format);
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  weight_split_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(w, out, rows, cols, ld, transposed, format);
  HS_LAUNCH_CHECK();
  return HS_OK;
}

int hs_weight_split_batch(const void* descs_dev, int n, int total_tiles, void* stream) {
  HS_REQUIRE(descs_dev && n > 0 && total_tiles > 0, "hs_weight_split_batch: bad arguments");
  static_assert(sizeof(SplitDesc) == 48, "descriptor layout is part of the ABI");
  weight_split_batch_kernel<<<total_tiles, dim3(32, 8), 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const SplitDesc*>(descs_dev), n);
  HS_LAUNCH_CHECK();
  return HS_OK;
}

int hs_gemm3_supported(int64_t T, int N, int K) {
  return (T >= 1 && N >= 8 && K >= 4 && N % 4 == 0 && K % 4 == 0 && K <= 8192 && N <= 65536) ? 1 : 0;
}

int hs_gemm3(const float* a_dev, const uint16_t* wsplit_dev, const float* bias_dev, const float* aux_dev, float* d_dev,
             float* d2_dev, float* colsum_dev, int64_t T, int N, int K, int mode, int precision, float drop, uint64_t seed,
             void* stream) {
  HS_REQUIRE(a_dev && wsplit_dev && d_dev && T > 0, "hs_gemm3: bad arguments");
  HS_REQUIRE(precision >= PREC_BF16X3 && precision <= PREC_BF16, "hs_gemm3: unknown precision %d", precision);
  HS_REQUIRE((mode >= MODE_PLAIN && mode <= MODE_GELU_GRAD) || mode == MODE_GELU_C || mode == MODE_GELU_GRAD_C,
             "hs_gemm3: unknown mode %d", mode);
  HS_REQUIRE(!(mode == MODE_GELU_C || mode == MODE_GELU_GRAD_C) || N % 32 == 0,
             "hs_gemm3: the compact GELU modes need N to be a multiple of 32, got %d", N);
  HS_REQUIRE(drop >= 0.f && drop < 1.f, "hs_gemm3: drop must be in [0, 1), got %f", drop);
  if (!hs_gemm3_supported(T, N, K))
    return hs::fail(HS_ERR_UNSUPPORTED, "hs_gemm3: shape T=%lld N=%d K=%d is not covered (N, K multiples of 4)",
                    (long long)T, N, K);
  const bool aux = (mode == MODE_ADD || mode == MODE_GELU_GRAD);
  HS_REQUIRE(!aux || aux_dev, "hs_gemm3: mode %d needs the aux tensor", mode);
  HS_REQUIRE((mode != MODE_GELU && mode != MODE_GELU_C) || d2_dev, "hs_gemm3: the GELU modes need the second output");
  HS_REQUIRE(mode != MODE_GELU_GRAD_C || aux_dev, "hs_gemm3: mode 7 needs the FP16 g' tensor (aux)");
  HS_REQUIRE(!colsum_dev || mode == MODE_PLAIN || mode == MODE_ADD, "hs_gemm3: column sums ride with modes 0 and 1 only");
  HS_REQUIRE(!((reinterpret_cast<uintptr_t>(a_dev) | reinterpret_cast<uintptr_t>(wsplit_dev) |
                reinterpret_cast<uintptr_t>(bias_dev) | reinterpret_cast<uintptr_t>(aux_dev) |
                reinterpret_cast<uintptr_t>(d_dev) | reinterpret_cast<uintptr_t>(d2_dev)) & 15),
             "hs_gemm3: tensors must be 16-byte aligned");
  G3Args a{};
  a.bias = bias_dev; a.colsum = colsum_dev; a.T = T; a.N = N; a.K = K; a.prec = precision;
  a.drop_thresh = drop > 0.f ? hs::drop_thresh(drop) : 0u;
  a.drop_scale = 1.0f / (1.0f - drop);
  a.seed = seed;
  cudaStream_t st = (cudaStream_t)stream;
  switch (mode) {
    case MODE_PLAIN: return launch<8, MODE_PLAIN>(a_dev, wsplit_dev, nullptr, d_dev, nullptr, a, st);
    case MODE_ADD: return launch<8, MODE_ADD>(a_dev, wsplit_dev, aux_dev, d_dev, nullptr, a, st);
    case MODE_GELU: return launch<16, MODE_GELU>(a_dev, wsplit_dev, nullptr, d_dev, d2_dev, a, st);
    case MODE_GELU_C:  // d_dev: the FP16 g' tensor (written from registers), d2_dev: h (the one TMA-stored output)
      a.gp16 = reinterpret_cast<uint16_t*>(d_dev);
      return launch<16, MODE_GELU_C>(a_dev, wsplit_dev, nullptr, d2_dev, nullptr, a, st);
    case MODE_GELU_GRAD_C:  // aux_dev: the FP16 g' tensor (read straight from global memory)
      a.gp16 = reinterpret_cast<uint16_t*>(const_cast<float*>(aux_dev));
      if (K <= 256) return launch<16, MODE_GELU_GRAD_C>(a_dev, wsplit_dev, nullptr, d_dev, nullptr, a, st);
      return launch<8, MODE_GELU_GRAD_C>(a_dev, wsplit_dev, nullptr, d_dev, nullptr, a, st);
    default:
      // the GELU' arithmetic wants 16 epilogue warps where the launch is HBM-bound (short contraction); with a long
      // contraction it is tensor-bound and the shared memory is better spent on the operand rings
      if (K <= 256) return launch<16, MODE_GELU_GRAD>(a_dev, wsplit_dev, aux_dev, d_dev, nullptr, a, st);
      return launch<8, MODE_GELU_GRAD>(a_dev, wsplit_dev, aux_dev, d_dev, nullptr, a, st);
  }
}

int hs_gemm3_ln_supported(int64_t T, int N, int K, int G) {
  return (hs_gemm3_supported(T, N, K) && G >= 32 && G <= 192 && G % 32 == 0 && N % G == 0) ? 1 : 0;
}

int hs_gemm3_ln(const float* a_dev, const uint16_t* wsplit_dev, const float* bias_dev, const float* gamma_dev,
                const float* beta_dev, const float* aux_dev, float* pre_dev, float* y_dev, float* mean_dev, float* rstd_dev,
                int64_t T, int N, int K, int G, float eps, int precision, void* stream) {
  HS_REQUIRE(a_dev && wsplit_dev && gamma_dev && beta_dev && y_dev && T > 0, "hs_gemm3_ln: bad arguments");
  HS_REQUIRE(precision == PREC_BF16X3 || precision == PREC_BF16, "hs_gemm3_ln: precision must be bf16x3 or bf16");
  HS_REQUIRE((mean_dev == nullptr) == (rstd_dev == nullptr), "hs_gemm3_ln: mean and rstd go together");
  if (!hs_gemm3_ln_supported(T, N, K, G))
    return hs::fail(HS_ERR_UNSUPPORTED,
                    "hs_gemm3_ln: shape T=%lld N=%d K=%d G=%d is not covered (G a multiple of 32, <= 192, dividing N)",
                    (long long)T, N, K, G);
  HS_REQUIRE(!((reinterpret_cast<uintptr_t>(a_dev) | reinterpret_cast<uintptr_t>(wsplit_dev) |
                reinterpret_cast<uintptr_t>(bias_dev) | reinterpret_cast<uintptr_t>(aux_dev) |
                reinterpret_cast<uintptr_t>(pre_dev) | reinterpret_cast<uintptr_t>(y_dev) |
                reinterpret_cast<uintptr_t>(gamma_dev) | reinterpret_cast<uintptr_t>(beta_dev)) & 15),
             "hs_gemm3_ln: tensors must be 16-byte aligned");
  G3Args a{};
  a.bias = bias_dev; a.T = T; a.N = N; a.K = K; a.prec = precision;
  a.drop_scale = 1.0f;
  a.gamma = gamma_dev; a.beta = beta_dev; a.mean_out = mean_dev; a.rstd_out = rstd_dev;
  a.G = G; a.eps = eps; a.save_pre = pre_dev ? 1 : 0; a.has_aux = aux_dev ? 1 : 0;
  return launch<8, MODE_LN>(a_dev, wsplit_dev, aux_dev, pre_dev ? pre_dev : y_dev, y_dev, a, (cudaStream_t)stream);
}

int hs_gemm3_lnin_supported(int64_t T, int N, int K) {
  return (hs_gemm3_supported(T, N, K) && K % 32 == 0 && K >= 160) ? 1 : 0;
}

int hs_gemm3_lnin(const float* a_dev, const uint16_t* wsplit_dev, const float* wsum_dev, const float* b0_dev, float* d_dev,
                  float* mean_dev, float* rstd_dev, int64_t T, int N, int K, float eps, int precision, void* stream) {
  HS_REQUIRE(a_dev && wsplit_dev && wsum_dev && d_dev && T > 0, "hs_gemm3_lnin: bad arguments");
  HS_REQUIRE(precision == PREC_BF16X3 || precision == PREC_BF16, "hs_gemm3_lnin: precision must be bf16x3 or bf16");
  HS_REQUIRE((mean_dev == nullptr) == (rstd_dev == nullptr), "hs_gemm3_lnin: mean and rstd go together");
  if (!hs_gemm3_lnin_supported(T, N, K))
    return hs::fail(HS_ERR_UNSUPPORTED, "hs_gemm3_lnin: shape T=%lld N=%d K=%d is not covered (K a multiple of 32, >= 160)",
                    (long long)T, N, K);
  HS_REQUIRE(!((reinterpret_cast<uintptr_t>(a_dev) | reinterpret_cast<uintptr_t>(wsplit_dev) |
                reinterpret_cast<uintptr_t>(wsum_dev) | reinterpret_cast<uintptr_t>(b0_dev) |
                reinterpret_cast<uintptr_t>(d_dev)) & 15),
             "hs_gemm3_lnin: tensors must be 16-byte aligned");
  G3Args a{};
  a.bias = b0_dev; a.wsum = wsum_dev; a.T = T; a.N = N; a.K = K; a.prec = precision;
  a.drop_scale = 1.0f;
  a.mean_out = mean_dev; a.rstd_out = rstd_dev; a.eps = eps;
  return launch<8, MODE_LN_IN>(a_dev, wsplit_dev, nullptr, d_dev, nullptr, a, (cudaStream_t)stream);
}

}  // extern "C"
