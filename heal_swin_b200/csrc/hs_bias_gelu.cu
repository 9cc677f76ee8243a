// h = GELU(z + bias) and its adjoint for the MLP of every block (swin_hp_transformer.py:21-44: fc1 -> nn.GELU -> fc2).
// The fc1 GEMM runs without bias; this kernel adds the bias, applies the exact (erf) GELU in one pass over the (rows,
// 4C) hidden tensor, and the backward produces dz = dh * GELU'(z + bias) together with the column sums d(bias) --
// the separate bias-gradient reduction pass over the hidden tensor disappears.  Pure HBM streaming, 128-bit accesses.
//
// Thread layout: a thread owns one float4 column group for the whole kernel (so its d(bias) partial sums stay in
// registers); a block covers blockDim / (C/4) consecutive rows per iteration.
#include "hs_common.h"
#include "hs_gelu.cuh"

namespace {


__device__ __forceinline__ float gelu_f(float u) { return hs::gelu_fast(u); }
__device__ __forceinline__ float gelu_grad_f(float u) { return hs::gelu_grad_fast(u); }

struct Drop {
  uint32_t thresh;  // 0 = off
  float scale;
  uint64_t seed;
};

__device__ __forceinline__ float4 drop_mult(const Drop& d, long long row, int col) {
  const uint32_t key = hs::drop_row_key(d.seed, row);
  float4 m;
  m.x = hs::drop_keep_elem(key, col + 0, d.thresh) ? d.scale : 0.f;
  m.y = hs::drop_keep_elem(key, col + 1, d.thresh) ? d.scale : 0.f;
  m.z = hs::drop_keep_elem(key, col + 2, d.thresh) ? d.scale : 0.f;
  m.w = hs::drop_keep_elem(key, col + 3, d.thresh) ? d.scale : 0.f;
  return m;
}

__global__ void __launch_bounds__(1024)
bias_gelu_fwd_kernel(const float4* __restrict__ z, const float4* __restrict__ bias, float4* __restrict__ h, long long rows,
                     int C4, const Drop dr) {
  const int col = threadIdx.x % C4, roff = threadIdx.x / C4, rpb = blockDim.x / C4;
  if (roff >= rpb) return;
  const float4 b = bias ? __ldg(bias + col) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long r = (long long)blockIdx.x * rpb + roff; r < rows; r += (long long)gridDim.x * rpb) {
    const float4 v = __ldcs(z + r * C4 + col);
    float4 o;
    o.x = gelu_f(v.x + b.x); o.y = gelu_f(v.y + b.y); o.z = gelu_f(v.z + b.z); o.w = gelu_f(v.w + b.w);
    if (dr.thresh) {
      const float4 m = drop_mult(dr, r, 4 * col);
      o.x *= m.x; o.y *= m.y; o.z *= m.z; o.w *= m.w;
    }
    h[r * C4 + col] = o;
  }
}

__global__ void __launch_bounds__(1024)
bias_gelu_bwd_kernel(const float4* __restrict__ dh, const float4* __restrict__ z, const float4* __restrict__ bias,
                     float4* __restrict__ dz, float* __restrict__ dbias, long long rows, int C4, const Drop dr) {
  extern __shared__ float red[];  // [4 * C4]
  const int col = threadIdx.x % C4, roff = threadIdx.x / C4, rpb = blockDim.x / C4;
  for (int i = threadIdx.x; i < 4 * C4; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  if (roff < rpb) {
    const float4 b = bias ? __ldg(bias + col) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long r = (long long)blockIdx.x * rpb + roff; r < rows; r += (long long)gridDim.x * rpb) {
      const float4 v = __ldcs(z + r * C4 + col);
      const float4 g = __ldcs(dh + r * C4 + col);
      float4 o;
      o.x = g.x * gelu_grad_f(v.x + b.x); o.y = g.y * gelu_grad_f(v.y + b.y);
      o.z = g.z * gelu_grad_f(v.z + b.z); o.w = g.w * gelu_grad_f(v.w + b.w);
      if (dr.thresh) {
        const float4 m = drop_mult(dr, r, 4 * col);
        o.x *= m.x; o.y *= m.y; o.z *= m.z; o.w *= m.w;
      }
      dz[r * C4 + col] = o;
      acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
    }
    if (dbias) {
      atomicAdd(red + 4 * col + 0, acc.x); atomicAdd(red + 4 * col + 1, acc.y);
      atomicAdd(red + 4 * col + 2, acc.z); atomicAdd(red + 4 * col + 3, acc.w);
    }
  }
  __syncthreads();
  if (dbias)
    for (int i = threadIdx.x; i < 4 * C4; i += blockDim.x) atomicAdd(dbias + i, red[i]);
}

int pick_block(int C4) {  // a multiple of C4, as close to 512 threads as possible (<= 1024)
  if (C4 <= 0 || C4 > 1024) return 0;
  int k = 512 / C4;
  if (k < 1) k = 1;
  return k * C4;
}

int num_sms() {
  // per device (a process may drive several GPUs): an immutable cache, filled on first use
  static int cached[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return 148;
  if (!cached[dev]) cudaDeviceGetAttribute(&cached[dev], cudaDevAttrMultiProcessorCount, dev);
  return cached[dev] ? cached[dev] : 148;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

extern "C" {

static Drop make_drop(float p, uint64_t seed) {
  Drop d;
  d.thresh = p > 0.f ? hs::drop_thresh(p) : 0u;
  d.scale = 1.0f / (1.0f - p);
  d.seed = seed;
  return d;
}

int hs_bias_gelu_supported(int64_t rows, int C) { return (rows > 0 && C > 0 && C % 4 == 0 && pick_block(C / 4) > 0) ? 1 : 0; }

int hs_bias_gelu_fwd(const float* z, const float* bias, float drop, uint64_t seed, float* h, int64_t rows, int C,
                     void* stream) {
  HS_REQUIRE(z && h && rows > 0 && C > 0, "hs_bias_gelu_fwd: bad arguments");
  HS_REQUIRE(drop >= 0.f && drop < 1.f, "hs_bias_gelu_fwd: drop must be in [0, 1), got %f", drop);
  const int block = (C % 4 == 0) ? pick_block(C / 4) : 0;
  if (!block || !aligned16(z) || !aligned16(h) || (bias && !aligned16(bias)))
    return hs::fail(HS_ERR_UNSUPPORTED, "hs_bias_gelu_fwd: needs C %% 4 == 0, C <= 4096 and 16-byte aligned tensors (C=%d)", C);
  const int rpb = block / (C / 4);
  long long grid = (rows + rpb - 1) / rpb;
  const long long cap = (long long)num_sms() * (2048 / block) * 2;
  if (grid > cap) grid = cap;
  bias_gelu_fwd_kernel<<<(int)grid, block, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(z), reinterpret_cast<const float4*>(bias), reinterpret_cast<float4*>(h), rows, C / 4,
      make_drop(drop, seed));
  HS_LAUNCH_CHECK();
  return HS_OK;
}

int hs_bias_gelu_bwd(const float* dh, const float* z, const float* bias, float drop, uint64_t seed, float* dz,
                     float* dbias, int64_t rows, int C, void* stream) {
  HS_REQUIRE(dh && z && dz && rows > 0 && C > 0, "hs_bias_gelu_bwd: bad arguments");
  HS_REQUIRE(drop >= 0.f && drop < 1.f, "hs_bias_gelu_bwd: drop must be in [0, 1), got %f", drop);
  const int block = (C % 4 == 0) ? pick_block(C / 4) : 0;
  if (!block || !aligned16(z) || !aligned16(dh) || !aligned16(dz) || (bias && !aligned16(bias)))
    return hs::fail(HS_ERR_UNSUPPORTED, "hs_bias_gelu_bwd: needs C %% 4 == 0, C <= 4096 and 16-byte aligned tensors (C=%d)", C);
  const int rpb = block / (C / 4);
  long long grid = (rows + rpb - 1) / rpb;
  const long long cap = (long long)num_sms() * (2048 / block) * 2;
  if (grid > cap) grid = cap;
  bias_gelu_bwd_kernel<<<(int)grid, block, (size_t)C * sizeof(float), (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(dh), reinterpret_cast<const float4*>(z), reinterpret_cast<const float4*>(bias),
      reinterpret_cast<float4*>(dz), dbias, rows, C / 4, make_drop(drop, seed));
  HS_LAUNCH_CHECK();
  return HS_OK;
}

}  // extern "C"
