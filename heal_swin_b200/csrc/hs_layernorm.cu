// Row LayerNorm (+ fused residual add) forward / backward for the HEAL-SWIN hot path: norm1 / norm2 of every block
// (swin_hp_transformer.py:316, 333-338 -- in the v2 placement  x = shortcut + LN(branch)  is one pass), the
// LayerNorm(4C) of PatchMerging (:392), the LayerNorm(C/2) of PatchExpand (:428), FinalPatchExpand_X4 (:450) and the
// encoder / decoder final norms.  Pure HBM streaming: rows of C fp32 channels, C % 4 == 0, 16 <= C <= 1536.
//
// A row is owned by T lanes of a warp (T a power of two, each lane V float4 = 4V channels, T*V = C/4), so a warp
// covers 32/T consecutive rows with every lane busy and every access a fully coalesced 128-bit load / store (C = 96:
// T = 8, V = 3, 4 rows per warp).  Statistics are two-pass in registers (mean, then centred sum of squares); the
// backward keeps its column partial sums of d(gamma) / d(beta) in registers across the rows a thread visits and
// flushes them once per CTA through shared memory + one atomicAdd per column.
#include "hs_common.h"

namespace {

constexpr int kThreads = 256;

// optional extras of the fused LayerNorm: dropout on its input (after the pre-bias) and a per-sample scale on its output
// (stochastic depth):   y = residual + row_scale[row / rows_per_scale] * (LN(dropout(x + pre_bias)) * gamma + beta)
struct LnExtra {
  const float* row_scale;  // null = 1
  long long rows_per_scale;
  uint32_t drop_thresh;    // 0 = no dropout
  float drop_scale;        // 1 / (1 - p)
  uint64_t seed;
};

__device__ __forceinline__ void drop4(float4& a, uint32_t key, int col, const LnExtra& e) {
  a.x = hs::drop_keep_elem(key, col + 0, e.drop_thresh) ? a.x * e.drop_scale : 0.f;
  a.y = hs::drop_keep_elem(key, col + 1, e.drop_thresh) ? a.y * e.drop_scale : 0.f;
  a.z = hs::drop_keep_elem(key, col + 2, e.drop_thresh) ? a.z * e.drop_scale : 0.f;
  a.w = hs::drop_keep_elem(key, col + 3, e.drop_thresh) ? a.w * e.drop_scale : 0.f;
}

__device__ __forceinline__ float group_sum(float v, int T) {
  for (int o = T >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int V, bool kDrop>
__global__ void __launch_bounds__(kThreads, (V <= 3 ? (kDrop ? 4 : 5) : (V <= 6 ? 3 : 1)))
ln_fwd_kernel(const float4* __restrict__ x, const float4* __restrict__ pre_bias, const float4* __restrict__ res,
              const float4* __restrict__ gamma, const float4* __restrict__ beta, float4* __restrict__ y,
              float* __restrict__ mean_out, float* __restrict__ rstd_out, long long rows, int T, float eps,
              const LnExtra ex) {
  const int lane = threadIdx.x & 31;
  const int t = lane & (T - 1), sub = lane / T, rpw = 32 / T;
  const int C4 = T * V;
  const float invC = 1.0f / (float)(4 * C4);
  const long long warp0 = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const long long stride = (long long)gridDim.x * (kThreads / 32) * rpw;
  for (long long base = warp0 * rpw; base < rows; base += stride) {
    const long long row = base + sub;
    const bool ok = row < rows;
    float4 a[V], r4[V];
    float s = 0.f;
    const uint32_t dkey = kDrop ? hs::drop_row_key(ex.seed, row) : 0u;
    // the residual is requested together with x: issued after the two row reductions its latency was exposed once per row
#pragma unroll
    for (int v = 0; v < V; ++v)
      r4[v] = (ok && res) ? __ldcs(res + row * C4 + t + T * v) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      a[v] = ok ? __ldcs(x + row * C4 + t + T * v) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (pre_bias) {  // the bias of the Linear that produced x (its GEMM ran without bias)
        const float4 pb = __ldg(pre_bias + t + T * v);
        a[v].x += pb.x; a[v].y += pb.y; a[v].z += pb.z; a[v].w += pb.w;
      }
      if (kDrop) drop4(a[v], dkey, 4 * (t + T * v), ex);
      s += (a[v].x + a[v].y) + (a[v].z + a[v].w);
    }
    const float mu = group_sum(s, T) * invC;
    float q = 0.f;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      a[v].x -= mu; a[v].y -= mu; a[v].z -= mu; a[v].w -= mu;
      q += (a[v].x * a[v].x + a[v].y * a[v].y) + (a[v].z * a[v].z + a[v].w * a[v].w);
    }
    const float rs = rsqrtf(group_sum(q, T) * invC + eps);
    if (ok) {
      const float osc = ex.row_scale ? __ldg(ex.row_scale + row / ex.rows_per_scale) : 1.0f;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const float4 g = __ldg(gamma + t + T * v), b = __ldg(beta + t + T * v);  // L1-resident; not kept in registers
        float4 o;
        o.x = fmaf(a[v].x * rs, g.x, b.x) * osc;
        o.y = fmaf(a[v].y * rs, g.y, b.y) * osc;
        o.z = fmaf(a[v].z * rs, g.z, b.z) * osc;
        o.w = fmaf(a[v].w * rs, g.w, b.w) * osc;
        o.x += r4[v].x; o.y += r4[v].y; o.z += r4[v].z; o.w += r4[v].w;
        y[row * C4 + t + T * v] = o;
      }
      if (t == 0 && mean_out) {
        mean_out[row] = mu;
        rstd_out[row] = rs;
      }
    }
  }
}

template <int V, bool kPreBias, bool kDrop>
__global__ void __launch_bounds__(kThreads, (V <= 3 ? ((kDrop || kPreBias) ? 2 : 3) : (V <= 6 ? 2 : 1)))
ln_bwd_kernel(const float4* __restrict__ dy, const float4* __restrict__ x, const float4* __restrict__ pre_bias,
              const float* __restrict__ mean, const float* __restrict__ rstd, const float4* __restrict__ gamma,
              float4* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
              float* __restrict__ dpre_bias, long long rows, int T, const LnExtra ex) {
  extern __shared__ float red[];  // [2 or 3][C]
  const int lane = threadIdx.x & 31;
  const int t = lane & (T - 1), sub = lane / T, rpw = 32 / T;
  const int C4 = T * V, C = 4 * C4;
  const float invC = 1.0f / (float)C;
  for (int i = threadIdx.x; i < (kPreBias ? 3 : 2) * C; i += kThreads) red[i] = 0.f;
  float4 dg[V], db[V], pb[kPreBias ? V : 1], dpb[kPreBias ? V : 1];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    dg[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kPreBias) {
      pb[v] = __ldg(pre_bias + t + T * v);
      dpb[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __syncthreads();
  const long long warp0 = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const long long stride = (long long)gridDim.x * (kThreads / 32) * rpw;
  for (long long base = warp0 * rpw; base < rows; base += stride) {
    const long long row = base + sub;
    const bool ok = row < rows;
    const float mu = ok ? __ldg(mean + row) : 0.f, rs = ok ? __ldg(rstd + row) : 0.f;
    const float osc = (ok && ex.row_scale) ? __ldg(ex.row_scale + row / ex.rows_per_scale) : 1.0f;
    const uint32_t dkey = kDrop ? hs::drop_row_key(ex.seed, row) : 0u;
    float4 xh[V], w[V];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      float4 xv = ok ? __ldcs(x + row * C4 + t + T * v) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 d = ok ? __ldcs(dy + row * C4 + t + T * v) : make_float4(0.f, 0.f, 0.f, 0.f);
      d.x *= osc; d.y *= osc; d.z *= osc; d.w *= osc;  // gradient w.r.t. the (unscaled) LayerNorm output
      if (kPreBias && ok) {
        xv.x += pb[v].x; xv.y += pb[v].y; xv.z += pb[v].z; xv.w += pb[v].w;
      }
      if (kDrop) drop4(xv, dkey, 4 * (t + T * v), ex);
      const float4 g = __ldg(gamma + t + T * v);  // L1-resident; not kept in registers
      xh[v].x = (xv.x - mu) * rs; xh[v].y = (xv.y - mu) * rs; xh[v].z = (xv.z - mu) * rs; xh[v].w = (xv.w - mu) * rs;
      w[v].x = d.x * g.x; w[v].y = d.y * g.y; w[v].z = d.z * g.z; w[v].w = d.w * g.w;
      s1 += (w[v].x + w[v].y) + (w[v].z + w[v].w);
      s2 += (w[v].x * xh[v].x + w[v].y * xh[v].y) + (w[v].z * xh[v].z + w[v].w * xh[v].w);
      dg[v].x = fmaf(d.x, xh[v].x, dg[v].x); dg[v].y = fmaf(d.y, xh[v].y, dg[v].y);
      dg[v].z = fmaf(d.z, xh[v].z, dg[v].z); dg[v].w = fmaf(d.w, xh[v].w, dg[v].w);
      db[v].x += d.x; db[v].y += d.y; db[v].z += d.z; db[v].w += d.w;
    }
    s1 = group_sum(s1, T) * invC;
    s2 = group_sum(s2, T) * invC;
    if (ok) {
#pragma unroll
      for (int v = 0; v < V; ++v) {
        float4 o;
        o.x = rs * (w[v].x - s1 - xh[v].x * s2);
        o.y = rs * (w[v].y - s1 - xh[v].y * s2);
        o.z = rs * (w[v].z - s1 - xh[v].z * s2);
        o.w = rs * (w[v].w - s1 - xh[v].w * s2);
        if (kDrop) drop4(o, dkey, 4 * (t + T * v), ex);  // back through the input dropout
        dx[row * C4 + t + T * v] = o;
        if (kPreBias) {
          dpb[v].x += o.x; dpb[v].y += o.y; dpb[v].z += o.z; dpb[v].w += o.w;
        }
      }
    }
  }
  // column sums: lanes with equal t (different sub) hold the same columns -> xor-reduce over the row index bits
#pragma unroll
  for (int v = 0; v < V; ++v) {
    for (int o = T; o < 32; o <<= 1) {
      dg[v].x += __shfl_xor_sync(0xffffffffu, dg[v].x, o); dg[v].y += __shfl_xor_sync(0xffffffffu, dg[v].y, o);
      dg[v].z += __shfl_xor_sync(0xffffffffu, dg[v].z, o); dg[v].w += __shfl_xor_sync(0xffffffffu, dg[v].w, o);
      db[v].x += __shfl_xor_sync(0xffffffffu, db[v].x, o); db[v].y += __shfl_xor_sync(0xffffffffu, db[v].y, o);
      db[v].z += __shfl_xor_sync(0xffffffffu, db[v].z, o); db[v].w += __shfl_xor_sync(0xffffffffu, db[v].w, o);
    }
    if (sub == 0) {
      const int c = 4 * (t + T * v);
      atomicAdd(red + c + 0, dg[v].x); atomicAdd(red + c + 1, dg[v].y);
      atomicAdd(red + c + 2, dg[v].z); atomicAdd(red + c + 3, dg[v].w);
      atomicAdd(red + C + c + 0, db[v].x); atomicAdd(red + C + c + 1, db[v].y);
      atomicAdd(red + C + c + 2, db[v].z); atomicAdd(red + C + c + 3, db[v].w);
    }
    if (kPreBias) {
      for (int o = T; o < 32; o <<= 1) {
        dpb[v].x += __shfl_xor_sync(0xffffffffu, dpb[v].x, o); dpb[v].y += __shfl_xor_sync(0xffffffffu, dpb[v].y, o);
        dpb[v].z += __shfl_xor_sync(0xffffffffu, dpb[v].z, o); dpb[v].w += __shfl_xor_sync(0xffffffffu, dpb[v].w, o);
      }
      if (sub == 0) {
        const int c = 4 * (t + T * v);
        atomicAdd(red + 2 * C + c + 0, dpb[v].x); atomicAdd(red + 2 * C + c + 1, dpb[v].y);
        atomicAdd(red + 2 * C + c + 2, dpb[v].z); atomicAdd(red + 2 * C + c + 3, dpb[v].w);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += kThreads) {
    if (dgamma) atomicAdd(dgamma + i, red[i]);
    if (dbeta) atomicAdd(dbeta + i, red[C + i]);
    if (kPreBias && dpre_bias) atomicAdd(dpre_bias + i, red[2 * C + i]);
  }
}


// Generic fallback for channel counts the vector kernels do not cover (C % 4 != 0 or an odd factor > 3, e.g. the
// reference's own embed_dim = 2 test config): one warp per row, scalar accesses.  Correct, not fast.
__global__ void __launch_bounds__(kThreads)
ln_fwd_generic_kernel(const float* __restrict__ x, const float* __restrict__ pre_bias, const float* __restrict__ res,
                      const float* __restrict__ gamma,
                      const float* __restrict__ beta, float* __restrict__ y, float* __restrict__ mean_out,
                      float* __restrict__ rstd_out, long long rows, int C, float eps, const LnExtra ex) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const long long stride = (long long)gridDim.x * (kThreads / 32);
  for (long long row = warp0; row < rows; row += stride) {
    const float* xr = x + row * C;
    const uint32_t dkey = ex.drop_thresh ? hs::drop_row_key(ex.seed, row) : 0u;
    const float osc = ex.row_scale ? ex.row_scale[row / ex.rows_per_scale] : 1.0f;
    auto xin = [&](int c) {
      const float v = xr[c] + (pre_bias ? pre_bias[c] : 0.f);
      if (!ex.drop_thresh) return v;
      return hs::drop_keep_elem(dkey, c, ex.drop_thresh) ? v * ex.drop_scale : 0.f;
    };
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += xin(c);
    const float mu = group_sum(s, 32) / (float)C;
    float q = 0.f;
    for (int c = lane; c < C; c += 32) q += (xin(c) - mu) * (xin(c) - mu);
    const float rs = rsqrtf(group_sum(q, 32) / (float)C + eps);
    for (int c = lane; c < C; c += 32) {
      float o = fmaf((xin(c) - mu) * rs, gamma[c], beta[c]) * osc;
      if (res) o += res[row * C + c];
      y[row * C + c] = o;
    }
    if (lane == 0 && mean_out) {
      mean_out[row] = mu;
      rstd_out[row] = rs;
    }
  }
}

__global__ void __launch_bounds__(kThreads)
ln_bwd_generic_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ pre_bias,
                      const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                      float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                      float* __restrict__ dpre_bias, long long rows, int C, const LnExtra ex) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const long long stride = (long long)gridDim.x * (kThreads / 32);
  for (long long row = warp0; row < rows; row += stride) {
    const float mu = mean[row], rs = rstd[row];
    const uint32_t dkey = ex.drop_thresh ? hs::drop_row_key(ex.seed, row) : 0u;
    const float osc = ex.row_scale ? ex.row_scale[row / ex.rows_per_scale] : 1.0f;
    auto keepf = [&](int c) { return !ex.drop_thresh ? 1.0f : (hs::drop_keep_elem(dkey, c, ex.drop_thresh) ? ex.drop_scale : 0.f); };
    auto xin = [&](int c) { return (x[row * C + c] + (pre_bias ? pre_bias[c] : 0.f)) * keepf(c); };
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float xh = (xin(c) - mu) * rs, w = dy[row * C + c] * osc * gamma[c];
      s1 += w;
      s2 += w * xh;
    }
    s1 = group_sum(s1, 32) / (float)C;
    s2 = group_sum(s2, 32) / (float)C;
    for (int c = lane; c < C; c += 32) {
      const float d = dy[row * C + c] * osc, xh = (xin(c) - mu) * rs;
      const float o = rs * (d * gamma[c] - s1 - xh * s2) * keepf(c);
      dx[row * C + c] = o;
      if (dgamma) atomicAdd(dgamma + c, d * xh);
      if (dbeta) atomicAdd(dbeta + c, d);
      if (dpre_bias) atomicAdd(dpre_bias + c, o);
    }
  }
}

// C/4 = T * V with T a power of two <= 32 and V in {1, 2, 3, 4, 6, 8, 12}
bool pick_shape(int C, int* T, int* V) {
  if (C <= 0 || (C & 3)) return false;
  const int C4 = C / 4;
  const int vs[] = {1, 2, 3, 4, 6, 8, 12};
  int bestT = 0, bestV = 0;
  for (int v : vs) {
    if (C4 % v) continue;
    const int t = C4 / v;
    if (t > 32 || (t & (t - 1))) continue;
    if (t > bestT) { bestT = t; bestV = v; }  // widest row group first: most lanes per row, fewest registers
  }
  if (!bestT) return false;
  *T = bestT; *V = bestV;
  return true;
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

template <typename K>
int grid_for(K kernel, size_t smem, long long rows, int T) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  const long long rows_per_block = (long long)(kThreads / 32) * (32 / T);
  long long want = (rows + rows_per_block - 1) / rows_per_block;
  const long long cap = (long long)num_sms() * per_sm;
  if (want > cap) want = cap;
  return (int)(want < 1 ? 1 : want);
}

#define HS_LN_DISPATCH(V_, ...)           \
  switch (V_) {                           \
    case 1: { constexpr int VV = 1; __VA_ARGS__; } break;   \
    case 2: { constexpr int VV = 2; __VA_ARGS__; } break;   \
    case 3: { constexpr int VV = 3; __VA_ARGS__; } break;   \
    case 4: { constexpr int VV = 4; __VA_ARGS__; } break;   \
    case 6: { constexpr int VV = 6; __VA_ARGS__; } break;   \
    case 8: { constexpr int VV = 8; __VA_ARGS__; } break;   \
    default: { constexpr int VV = 12; __VA_ARGS__; } break; \
  }

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

extern "C" {

static int make_extra(const char* fn, const float* row_scale, int64_t rows_per_scale, float in_drop, uint64_t seed,
                      LnExtra* ex) {
  HS_REQUIRE(in_drop >= 0.f && in_drop < 1.f, "%s: in_drop must be in [0, 1), got %f", fn, in_drop);
  HS_REQUIRE(!row_scale || rows_per_scale > 0, "%s: row_scale needs rows_per_scale > 0", fn);
  ex->row_scale = row_scale;
  ex->rows_per_scale = rows_per_scale > 0 ? rows_per_scale : 1;
  ex->drop_thresh = in_drop > 0.f ? hs::drop_thresh(in_drop) : 0u;
  ex->drop_scale = 1.0f / (1.0f - in_drop);
  ex->seed = seed;
  return HS_OK;
}

int hs_layernorm_fwd(const float* x, const float* pre_bias, const float* residual, const float* gamma, const float* beta,
                     const float* row_scale, int64_t rows_per_scale, float in_drop, uint64_t seed, float* y, float* mean,
                     float* rstd, int64_t rows, int C, float eps, void* stream) {
  LnExtra ex;
  if (int rc = make_extra("hs_layernorm_fwd", row_scale, rows_per_scale, in_drop, seed, &ex)) return rc;
  HS_REQUIRE(x && gamma && beta && y, "hs_layernorm_fwd: null pointer");
  HS_REQUIRE((mean == nullptr) == (rstd == nullptr), "hs_layernorm_fwd: mean and rstd go together");
  HS_REQUIRE(rows > 0, "hs_layernorm_fwd: rows must be positive");
  HS_REQUIRE(C > 0, "hs_layernorm_fwd: C must be positive");
  int T, V;
  const bool vec = pick_shape(C, &T, &V) && aligned16(x) && aligned16(y) && aligned16(gamma) && aligned16(beta) &&
                   (!residual || aligned16(residual)) && (!pre_bias || aligned16(pre_bias));
  if (!vec) {
    long long blocks = (rows + kThreads / 32 - 1) / (kThreads / 32);
    if (blocks > (long long)num_sms() * 8) blocks = (long long)num_sms() * 8;
    ln_fwd_generic_kernel<<<(int)blocks, kThreads, 0, (cudaStream_t)stream>>>(x, pre_bias, residual, gamma, beta, y,
                                                                              mean, rstd, rows, C, eps, ex);
    HS_LAUNCH_CHECK();
    return HS_OK;
  }
#define HS_LN_FWD_LAUNCH(DROP)                                                                                   \
  HS_LN_DISPATCH(V, {                                                                                            \
    const int grid = grid_for(ln_fwd_kernel<VV, DROP>, 0, rows, T);                                              \
    ln_fwd_kernel<VV, DROP><<<grid, kThreads, 0, (cudaStream_t)stream>>>(                                        \
        reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(pre_bias),                           \
        reinterpret_cast<const float4*>(residual), reinterpret_cast<const float4*>(gamma),                       \
        reinterpret_cast<const float4*>(beta), reinterpret_cast<float4*>(y), mean, rstd, rows, T, eps, ex);      \
  })
  if (ex.drop_thresh) {
    HS_LN_FWD_LAUNCH(true);
  } else {
    HS_LN_FWD_LAUNCH(false);
  }
#undef HS_LN_FWD_LAUNCH
  HS_LAUNCH_CHECK();
  return HS_OK;
}

int hs_layernorm_bwd(const float* dy, const float* x, const float* pre_bias, const float* mean, const float* rstd,
                     const float* gamma, const float* row_scale, int64_t rows_per_scale, float in_drop, uint64_t seed,
                     float* dx, float* dgamma, float* dbeta, float* dpre_bias, int64_t rows, int C, void* stream) {
  LnExtra ex;
  if (int rc = make_extra("hs_layernorm_bwd", row_scale, rows_per_scale, in_drop, seed, &ex)) return rc;
  HS_REQUIRE(dy && x && mean && rstd && gamma && dx, "hs_layernorm_bwd: null pointer");
  HS_REQUIRE(rows > 0, "hs_layernorm_bwd: rows must be positive");
  HS_REQUIRE(C > 0, "hs_layernorm_bwd: C must be positive");
  int T, V;
  HS_REQUIRE(pre_bias || !dpre_bias, "hs_layernorm_bwd: dpre_bias without pre_bias");
  // the pre-bias variant keeps V more float4 accumulators: vector path for V <= 6 (C <= 768), generic beyond
  const bool vec = pick_shape(C, &T, &V) && aligned16(dy) && aligned16(x) && aligned16(dx) && aligned16(gamma) &&
                   (!pre_bias || (aligned16(pre_bias) && V <= 6));
  if (!vec) {
    long long blocks = (rows + kThreads / 32 - 1) / (kThreads / 32);
    if (blocks > (long long)num_sms() * 8) blocks = (long long)num_sms() * 8;
    ln_bwd_generic_kernel<<<(int)blocks, kThreads, 0, (cudaStream_t)stream>>>(dy, x, pre_bias, mean, rstd, gamma, dx,
                                                                              dgamma, dbeta, dpre_bias, rows, C, ex);
    HS_LAUNCH_CHECK();
    return HS_OK;
  }
  const size_t smem = (pre_bias ? 3 : 2) * (size_t)C * sizeof(float);
#define HS_LN_BWD_LAUNCH(PB, DROP)                                                                               \
  HS_LN_DISPATCH(V, {                                                                                            \
    constexpr int VB = (PB && VV > 6) ? 6 : VV; /* the pre-bias variant is only dispatched for V <= 6 */         \
    const int grid = grid_for(ln_bwd_kernel<VB, PB, DROP>, smem, rows, T);                                       \
    ln_bwd_kernel<VB, PB, DROP><<<grid, kThreads, smem, (cudaStream_t)stream>>>(                                 \
        reinterpret_cast<const float4*>(dy), reinterpret_cast<const float4*>(x),                                 \
        reinterpret_cast<const float4*>(pre_bias), mean, rstd, reinterpret_cast<const float4*>(gamma),           \
        reinterpret_cast<float4*>(dx), dgamma, dbeta, dpre_bias, rows, T, ex);                                   \
  })
  if (pre_bias) {
    if (ex.drop_thresh) { HS_LN_BWD_LAUNCH(true, true); } else { HS_LN_BWD_LAUNCH(true, false); }
  } else {
    if (ex.drop_thresh) { HS_LN_BWD_LAUNCH(false, true); } else { HS_LN_BWD_LAUNCH(false, false); }
  }
#undef HS_LN_BWD_LAUNCH
  HS_LAUNCH_CHECK();
  return HS_OK;
}

}  // extern "C"
