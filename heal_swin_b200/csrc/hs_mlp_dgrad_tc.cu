// Backward of the MLP's second half on the sm_100a tensor cores, fused:   dz = (dy @ W2) o GELU'(z + b1) o dropmask
// (autograd of  fc2(drop(GELU(fc1(x) + b1)))  w.r.t. the fc1 output z, swin_hp_transformer.py:21-44).  The library path
// is a dgrad GEMM that WRITES the (T, 4C) hidden gradient, followed by a GELU-backward pass that READS it, reads z and
// writes dz: this kernel never materialises the hidden gradient -- it reads dy (T, C), z (T, 4C) and writes dz (T, 4C).
//
// A CTA owns a 128-column chunk of the hidden dimension (its slice of W2 stays resident in shared memory as the MN-major
// B operand: 32-column slabs of C rows x 128 B, SWIZZLE_128B_ATOM_32B) and walks over 128-token tiles:
//   dy tile  -> ring of K-major 32-column sub-tiles (TMA, SWIZZLE_128B), 4 MMAs (K = 8) each into a 128 x 128 TMEM tile
//   epilogue -> four groups of four warps, one 32-column slab each: tcgen05.ld the accumulator, multiply IN PLACE with
//               GELU'(z + b1) on the z slab TMA put in shared memory (and the dropout mask), TMA-store the slab as dz;
//               two TMEM stages and a ring of slab buffers overlap all of it with the next tile's loads and MMAs
// Warps: 0-15 epilogue (group = warp / 4, TMEM lane quadrant = warp % 4), 16 dy producer, 17 MMA issuer, 18 z producer.
//
// Compact form (flag HS_MLP_GRAD16): the forward (hs_gemm3 mode 6) saved g' = GELU'(z + b1) o dropmask as FP16 instead of z.
// Then dz = (dy @ W2) o g': every epilogue thread reads its row's 32 g' values (64 bytes, two full sectors) straight
// from global memory before it waits for the accumulator -- no z slabs through shared memory (a third of this kernel's
// bytes), no GELU arithmetic, no mask regeneration; the slab buffers are output staging only.
#include <cuda_fp16.h>

#include "hs_common.h"
#include "hs_gelu.cuh"
#include "hs_sm100.cuh"
#include "hs_tc_common.cuh"

namespace {

using namespace hs::sm100;

constexpr int kBM = 128;                 // tokens per tile
constexpr int kNJ = 128;                 // hidden columns per CTA
constexpr int kMaxRing = 4;              // dy sub-tile ring depth (upper bound)
constexpr int kMaxSlabs = 8;             // z / dz slab buffers (upper bound)
constexpr int kSub = kBM * 128;          // bytes of one 32-column x 128-row tile (K-major sub-tile, z slab, dz slab)
constexpr int kEpiWarps = 16;
constexpr int kThreads = (kEpiWarps + 3) * 32;

struct MdArgs {
  const uint16_t* gp16;  // compact form: g' (T, J) as FP16, else null
  const float* b1;  // (J) or null
  long long T;
  int C, J;
  int n_chunks;      // J / 128
  long long tiles;   // ceil(T / 128)
  int ring, slabs;   // dy ring depth, number of slab buffers
  float fix;         // TF32 truncation compensation (two truncated operands)
  uint32_t drop_thresh;
  float drop_scale;
  uint64_t seed;
};

__device__ __forceinline__ float gelu_grad_f(float u) { return hs::gelu_grad_fast(u); }

template <bool kCompact>
__global__ void __launch_bounds__(kThreads, 1)
mlp_dgrad_gelu_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_w,
                      const __grid_constant__ CUtensorMap map_z, const __grid_constant__ CUtensorMap map_dz, const MdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int nsub = a.C / 32;                 // dy sub-tiles per token tile
  const int w_slab = a.C * 128;              // bytes of one 32-column slab of the resident W2 chunk
  const int ring = a.ring, R = a.slabs;
  uint8_t* s_w = sm;                         // 4 slabs
  uint8_t* s_dy = s_w + 4 * w_slab;          // ring sub-tiles
  uint8_t* s_z = s_dy + ring * kSub;         // R slab buffers (z in, dz out, in place)
  __shared__ uint64_t w_full, dy_full[kMaxRing], dy_empty[kMaxRing], acc_full[2], acc_empty[2], z_full[kMaxSlabs],
      z_empty[kMaxSlabs];
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.x % a.n_chunks;
  const long long t0 = blockIdx.x / a.n_chunks, tstep = gridDim.x / a.n_chunks;
  const int j0 = chunk * kNJ;

  if (threadIdx.x == 0) {
    mbar_init(&w_full, 1);
    for (int i = 0; i < kMaxRing; ++i) {
      mbar_init(&dy_full[i], 1);
      mbar_init(&dy_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], kEpiWarps * 32);
    }
    for (int i = 0; i < kMaxSlabs; ++i) {
      mbar_init(&z_full[i], 1);
      mbar_init(&z_empty[i], 1);
    }
    mbar_fence_init();
  }
  if (warp == kEpiWarps + 1) tmem_alloc(&tmem_base, 256);
  if (warp == kEpiWarps && lane == 0) {
    tma_prefetch_desc(&map_dy);
    tma_prefetch_desc(&map_w);
  }
  if (warp == kEpiWarps + 2 && lane == 0) {
    tma_prefetch_desc(&map_z);
    tma_prefetch_desc(&map_dz);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base;

  if (warp == kEpiWarps) {
    if (elect_one()) {
      // resident W2 chunk: 4 slabs of (C rows x 32 columns)
      mbar_arrive_expect_tx(&w_full, (uint32_t)(4 * w_slab));
      for (int s = 0; s < 4; ++s) tma_load_2d(s_w + s * w_slab, &map_w, &w_full, j0 + 32 * s, 0);
      int slot = 0;
      uint32_t ph = 0;
      for (long long tile = t0; tile < a.tiles; tile += tstep)
        for (int sub = 0; sub < nsub; ++sub) {
          mbar_wait(&dy_empty[slot], ph ^ 1);
          mbar_arrive_expect_tx(&dy_full[slot], kSub);
          tma_load_2d(s_dy + slot * kSub, &map_dy, &dy_full[slot], 32 * sub, (int)(tile * kBM));
          if (++slot == ring) { slot = 0; ph ^= 1; }
        }
    }
  } else if (warp == kEpiWarps + 1) {
    if (elect_one()) {
      constexpr uint64_t kDescK = umma_smem_desc(16, 1024, kLayoutSw128);
      const uint64_t kDescW = umma_smem_desc((uint32_t)w_slab, 512, kLayoutSw128B32);
      constexpr uint32_t kIdesc = umma_idesc_tf32(128, kNJ, 0, 1);
      mbar_wait(&w_full, 0);
      const uint32_t wb = smem_u32(s_w);
      int slot = 0;
      uint32_t ph = 0;
      long long it = 0;
      for (long long tile = t0; tile < a.tiles; tile += tstep, ++it) {
        const int as = (int)(it & 1);
        mbar_wait(&acc_empty[as], (((uint32_t)(it >> 1)) & 1) ^ 1);
        tc_fence_after();
        for (int sub = 0; sub < nsub; ++sub) {
          mbar_wait(&dy_full[slot], ph);
          tc_fence_after();
          const uint32_t ab = smem_u32(s_dy + slot * kSub);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_tf32_ss(tmem + (uint32_t)as * kNJ, umma_desc_at(kDescK, ab + ks * 32),
                         umma_desc_at(kDescW, wb + (sub * 4 + ks) * 1024), kIdesc, (sub > 0 || ks > 0) ? 1u : 0u);
          umma_commit(&dy_empty[slot]);
          if (++slot == ring) { slot = 0; ph ^= 1; }
        }
        umma_commit(&acc_full[as]);
      }
    }
  } else if (warp == kEpiWarps + 2) {
    if (!kCompact && elect_one()) {
      int buf = 0;
      uint32_t ph = 0;
      for (long long tile = t0; tile < a.tiles; tile += tstep)
        for (int s = 0; s < 4; ++s) {
          mbar_wait(&z_empty[buf], ph ^ 1);
          mbar_arrive_expect_tx(&z_full[buf], kSub);
          tma_load_2d(s_z + buf * kSub, &map_z, &z_full[buf], j0 + 32 * s, (int)(tile * kBM));
          if (++buf == R) { buf = 0; ph ^= 1; }
        }
    }
  } else {
    // ============================================ epilogue: group eg owns slab eg of every tile, thread = token row
    const int eg = warp >> 2, q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int jc = j0 + 32 * eg;
    float4 bv[kCompact ? 1 : 8];
    if constexpr (!kCompact) {
#pragma unroll
      for (int c = 0; c < 8; ++c)
        bv[c] = a.b1 ? __ldg(reinterpret_cast<const float4*>(a.b1 + jc + 4 * c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    long long it = 0;
    // compact form: this row's 32 g' values of the NEXT tile are requested as soon as the current ones have been consumed
    // (same registers), a whole tile period before they are needed
    uint4 gq[kCompact ? 4 : 1];
    auto fetch_gq = [&](long long tile) {
      const long long row = tile * kBM + r;
      const uint4* src = reinterpret_cast<const uint4*>(a.gp16 + row * a.J + jc);
#pragma unroll
      for (int i = 0; i < (kCompact ? 4 : 1); ++i)
        gq[i] = (tile < a.tiles && row < a.T) ? __ldcs(src + i) : make_uint4(0u, 0u, 0u, 0u);
    };
    if constexpr (kCompact) fetch_gq(t0);
    for (long long tile = t0; tile < a.tiles; tile += tstep, ++it) {
      const int as = (int)(it & 1);
      const long long g = it * 4 + eg;
      const int buf = (int)(g % R);
      const long long grow = tile * kBM + r;
      mbar_wait(&acc_full[as], ((uint32_t)(it >> 1)) & 1);
      tc_fence_after();
      uint32_t acc[32];
      tmem_ld32(tmem + lane_addr + (uint32_t)as * kNJ + 32 * eg, acc);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(&acc_empty[as]);  // this thread's part of the accumulator stage is in registers
      const uint32_t dkey = a.drop_thresh ? hs::drop_row_key(a.seed, grow) : 0u;
      uint8_t* zrow = s_z + buf * kSub + r * 128;
      if constexpr (kCompact) {
        // the buffer is free once the store that last used it has read it (z_empty; fresh barriers pass the first round)
        mbar_wait(&z_empty[buf], (((uint32_t)(g / R)) & 1) ^ 1);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t w0 = c & 1 ? gq[c >> 1].z : gq[c >> 1].x, w1 = c & 1 ? gq[c >> 1].w : gq[c >> 1].y;
          const float2 g01 = __half22float2(*reinterpret_cast<const __half2*>(&w0));
          const float2 g23 = __half22float2(*reinterpret_cast<const __half2*>(&w1));
          float4 o;
          o.x = __uint_as_float(acc[4 * c + 0]) * a.fix * g01.x;
          o.y = __uint_as_float(acc[4 * c + 1]) * a.fix * g01.y;
          o.z = __uint_as_float(acc[4 * c + 2]) * a.fix * g23.x;
          o.w = __uint_as_float(acc[4 * c + 3]) * a.fix * g23.y;
          *reinterpret_cast<float4*>(zrow + ((c ^ (r & 7)) << 4)) = o;
        }
        fetch_gq(tile + tstep);
      } else {
      mbar_wait(&z_full[buf], ((uint32_t)(g / R)) & 1);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float4* p = reinterpret_cast<float4*>(zrow + ((c ^ (r & 7)) << 4));
        const float4 zv = *p;
        float4 o;
        o.x = __uint_as_float(acc[4 * c + 0]) * a.fix * gelu_grad_f(zv.x + bv[c].x);
        o.y = __uint_as_float(acc[4 * c + 1]) * a.fix * gelu_grad_f(zv.y + bv[c].y);
        o.z = __uint_as_float(acc[4 * c + 2]) * a.fix * gelu_grad_f(zv.z + bv[c].z);
        o.w = __uint_as_float(acc[4 * c + 3]) * a.fix * gelu_grad_f(zv.w + bv[c].w);
        if (a.drop_thresh) {
          o.x = hs::drop_keep_elem(dkey, jc + 4 * c + 0, a.drop_thresh) ? o.x * a.drop_scale : 0.f;
          o.y = hs::drop_keep_elem(dkey, jc + 4 * c + 1, a.drop_thresh) ? o.y * a.drop_scale : 0.f;
          o.z = hs::drop_keep_elem(dkey, jc + 4 * c + 2, a.drop_thresh) ? o.z * a.drop_scale : 0.f;
          o.w = hs::drop_keep_elem(dkey, jc + 4 * c + 3, a.drop_thresh) ? o.w * a.drop_scale : 0.f;
        }
        *p = o;
      }
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + eg, 128);  // every row of the slab holds dz now
      if (r == 0) {
        tma_store_2d(&map_dz, s_z + buf * kSub, jc, (int)(tile * kBM));
        tma_store_commit();
        tma_store_wait_read<0>();   // the buffer can take the next z slab as soon as the store has read it
        mbar_arrive(&z_empty[buf]);
      }
    }
    if (r == 0) tma_store_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps + 1) tmem_dealloc(tmem, 256);
}

}  // namespace

extern "C" {

int hs_mlp_dgrad_gelu_supported(int64_t T, int C, int J) {
  return (T >= 1024 && C >= 32 && C <= 192 && C % 32 == 0 && J % 128 == 0 && J >= 128) ? 1 : 0;
}

int hs_mlp_dgrad_gelu(const float* dy, const float* w2, const float* z, const float* b1, float drop, uint64_t seed,
                      float* dz, int64_t T, int C, int J, uint32_t flags, void* stream) {
  HS_REQUIRE(dy && w2 && z && dz && T > 0, "hs_mlp_dgrad_gelu: bad arguments");
  HS_REQUIRE(drop >= 0.f && drop < 1.f, "hs_mlp_dgrad_gelu: drop must be in [0, 1), got %f", drop);
  if (!hs_mlp_dgrad_gelu_supported(T, C, J))
    return hs::fail(HS_ERR_UNSUPPORTED, "hs_mlp_dgrad_gelu: shape T=%lld C=%d J=%d is not covered (C <= 192, C %% 32 == 0, "
                    "J %% 128 == 0)", (long long)T, C, J);
  HS_REQUIRE(!((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(w2) | reinterpret_cast<uintptr_t>(z) |
                reinterpret_cast<uintptr_t>(dz) | reinterpret_cast<uintptr_t>(b1)) & 15), "hs_mlp_dgrad_gelu: unaligned tensor");
  MdArgs a{};
  if (flags & HS_MLP_GRAD16) {  // z_dev is the FP16 g' tensor of hs_gemm3 mode 6: no bias, no mask left to apply
    a.gp16 = reinterpret_cast<const uint16_t*>(z);
    b1 = nullptr;
    drop = 0.f;
  }
  a.b1 = b1; a.T = T; a.C = C; a.J = J; a.n_chunks = J / kNJ; a.tiles = (T + kBM - 1) / kBM;
  a.fix = (flags & HS_ATTN_NO_TRUNC_COMP) ? 1.0f : hs::tc::kTruncFix2;
  a.drop_thresh = drop > 0.f ? hs::drop_thresh(drop) : 0u;
  a.drop_scale = 1.0f / (1.0f - drop);
  a.seed = seed;
  CUtensorMap map_dy, map_w, map_z, map_dz;
  int rc;
  if ((rc = hs::tc::make_map(&map_dy, dy, T, C, CU_TENSOR_MAP_SWIZZLE_128B, 32, kBM))) return rc;
  if ((rc = hs::tc::make_map(&map_w, w2, C, J, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, 32, C))) return rc;
  if (!a.gp16 && (rc = hs::tc::make_map(&map_z, z, T, J, CU_TENSOR_MAP_SWIZZLE_128B, 32, kBM))) return rc;
  if ((rc = hs::tc::make_map(&map_dz, dz, T, J, CU_TENSOR_MAP_SWIZZLE_128B, 32, kBM))) return rc;
  // shared memory plan: resident W2 chunk, dy ring, and as many slab buffers as fit (227 KB per CTA)
  const int avail = 232448 - 1024 - 4 * C * 128;
  a.ring = C > 128 ? 2 : 3;
  a.slabs = (avail - a.ring * kSub) / kSub;
  if (a.slabs > kMaxSlabs) a.slabs = kMaxSlabs;
  if (a.slabs < 4) return hs::fail(HS_ERR_UNSUPPORTED, "hs_mlp_dgrad_gelu: C=%d leaves no room for the slab buffers", C);
  const size_t smem = (size_t)4 * C * 128 + (size_t)(a.ring + a.slabs) * kSub + 1024;
  int per_chunk = hs::tc::sm_count() / a.n_chunks;
  if (per_chunk < 1) per_chunk = 1;
  if (per_chunk > a.tiles) per_chunk = (int)a.tiles;
  if (a.gp16) {
    HS_CUDA(cudaFuncSetAttribute(mlp_dgrad_gelu_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mlp_dgrad_gelu_kernel<true><<<a.n_chunks * per_chunk, kThreads, smem, (cudaStream_t)stream>>>(map_dy, map_w, map_dz, map_dz, a);
  } else {
    HS_CUDA(cudaFuncSetAttribute(mlp_dgrad_gelu_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mlp_dgrad_gelu_kernel<false><<<a.n_chunks * per_chunk, kThreads, smem, (cudaStream_t)stream>>>(map_dy, map_w, map_z, map_dz, a);
  }
  HS_LAUNCH_CHECK();
  return HS_OK;
}

}  // extern "C"
