// Tail of the decoder, fused:  logits = Conv1d_{1x1}(LayerNorm(x))  in one pass over the (B * N_pix, C) activation
// (FinalPatchExpand_X4.norm followed by SwinHPTransformerSys.output, swin_hp_transformer.py:450 and :781-786 / 945), and
// its backward in one pass.  The normalised activation (2.4 GB at the bench size) is never written or re-read: the
// forward reads x and writes the f_out-channel logits directly in the (B, f_out, N_pix) layout the loss consumes; the
// backward reads x and d(logits), rebuilds d(LN output) = d(logits) . W on the fly, writes dx, and accumulates
//     S[k][c] = sum_rows d(logits)[row][k] * xhat[row][c],     G[k] = sum_rows d(logits)[row][k]
// from which every parameter gradient follows (host side, tiny):  dW = gamma * S + beta * G,  dgamma = sum_k W * S,
// dbeta = sum_k W * G,  dbias = G.
//
// Layout as in hs_layernorm.cu: a row of C = 32 V channels is owned by 8 lanes (V float4 each), a warp covers 4 R
// consecutive rows per iteration.  Forward: gamma is folded into the weights staged in shared memory
// (logit_k = rstd * sum_c (x_c - mean) gamma_c W_kc + sum_c beta_c W_kc).  Backward: the S accumulators are spread over
// the whole warp -- lane (sub, t) keeps its 4 V channels for the classes k = sub, sub + 4, ... -- and the lanes exchange
// xhat and d(logits) of the warp's rows through shared memory, so S costs 4 V * KQ registers instead of 4 V * K.
#include "hs_common.h"

namespace {

constexpr int kT = 8;  // lanes per row

__device__ __forceinline__ float group_sum8(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, acc))));
}

// One reduce-scatter step over the lane pair (lane, lane ^ mask): the N values are split into a lower and an upper half,
// "upper" lanes keep the upper half and send the lower one.  After the three steps (masks 4, 2, 1) of an 8-lane group
// every lane holds the complete sums of ceil(N / 8) values: 11 shuffles for 12 values instead of 36.
template <int N>
__device__ __forceinline__ void reduce_scatter_step(const float (&a)[N], float (&o)[(N + 1) / 2], bool upper, int mask) {
  constexpr int H = (N + 1) / 2;
#pragma unroll
  for (int i = 0; i < H; ++i) {
    const float lo = a[i], hi = (i + H < N) ? a[i + H] : 0.f;
    const float send = upper ? lo : hi, keep = upper ? hi : lo;
    o[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
  }
}

// (sample, pixel) of a row; 32-bit division when the row count allows it
__device__ __forceinline__ void split_row(long long row, long long P, bool small, long long& b, long long& p) {
  if (small) {
    const uint32_t bb = (uint32_t)row / (uint32_t)P;
    b = bb;
    p = (long long)((uint32_t)row - bb * (uint32_t)P);
  } else {
    b = row / P;
    p = row - b * P;
  }
}

template <int V, int KQ, int R, int TH>
__global__ void __launch_bounds__(TH, 2)
ln_head_fwd_kernel(const float4* __restrict__ x, const float4* __restrict__ gamma, const float* __restrict__ beta,
                   const float* __restrict__ w, const float* __restrict__ hbias, float* __restrict__ logits,
                   float* __restrict__ mean_out, float* __restrict__ rstd_out, long long rows, long long P, int K, float eps) {
  constexpr int KP = 4 * KQ, C4 = kT * V, C = 4 * C4;
  __shared__ float4 wg[KP][C4];  // gamma * W, zero rows for k >= K
  __shared__ float b0[KP];       // sum_c beta_c W_kc + bias_k
  for (int i = threadIdx.x; i < KP * C4; i += TH) {
    const int k = i / C4, c4 = i - k * C4;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < K) {
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(w) + (size_t)k * C4 + c4), g4 = __ldg(gamma + c4);
      o = make_float4(w4.x * g4.x, w4.y * g4.y, w4.z * g4.z, w4.w * g4.w);
    }
    wg[k][c4] = o;
  }
  if (threadIdx.x < KP) {
    const int k = threadIdx.x;
    float s = 0.f;
    if (k < K) {
      for (int c = 0; c < C; ++c) s = fmaf(__ldg(beta + c), __ldg(w + (size_t)k * C + c), s);
      if (hbias) s += __ldg(hbias + k);
    }
    b0[k] = s;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, t = lane & (kT - 1), sub = lane >> 3;
  const float invC = 1.0f / (float)C;
  const bool small = rows < (1LL << 31);
  // classes this lane ends up with after the reduce-scatter: kmine .. kmine + nmine - 1
  constexpr int H0 = (KP + 1) / 2, H1 = (H0 + 1) / 2, H2 = (H1 + 1) / 2;
  const bool up4 = (t & 4) != 0, up2 = (t & 2) != 0, up1 = (t & 1) != 0;
  const int kmine = (up4 ? H0 : 0) + (up2 ? H1 : 0) + (up1 ? H2 : 0);
  const int nmine = up1 ? H1 - H2 : H2;
  // the trip count is the same for every thread (rows beyond the end are predicated off), which lets the compiler
  // prove the warp converged at the shuffles: a thread-dependent loop bound wraps each of them in WARPSYNC.COLLECTIVE
  const long long warp0 = (long long)blockIdx.x * (TH / 32) + (threadIdx.x >> 5);
  const long long stride = (long long)gridDim.x * (TH / 32) * (4 * R);
  const long long iters = (rows + stride - 1) / stride;
  // software pipeline: the rows of iteration it + 1 are requested before those of iteration it are processed (16 warps
  // per SM do not hide the HBM latency of a load-then-compute loop: ncu showed long-scoreboard stalls first)
  float4 nx[R][V];
  auto fetch = [&](long long base) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long row = base + 4 * r + sub;
#pragma unroll
      for (int v = 0; v < V; ++v)
        nx[r][v] = row < rows ? __ldcs(x + row * C4 + t + kT * v) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  fetch(warp0 * (4 * R));
  for (long long it = 0; it < iters; ++it) {
    const long long base = warp0 * (4 * R) + it * stride;
    float4 a[R][V];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int v = 0; v < V; ++v) a[r][v] = nx[r][v];
    if (it + 1 < iters) fetch(base + stride);
    float rs[R], acc[R][KP];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long row = base + 4 * r + sub;
      const bool ok = row < rows;
      float s = 0.f;
#pragma unroll
      for (int v = 0; v < V; ++v) s += (a[r][v].x + a[r][v].y) + (a[r][v].z + a[r][v].w);
      const float mu = group_sum8(s) * invC;
      float q = 0.f;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        a[r][v].x -= mu; a[r][v].y -= mu; a[r][v].z -= mu; a[r][v].w -= mu;
        q += (a[r][v].x * a[r][v].x + a[r][v].y * a[r][v].y) + (a[r][v].z * a[r][v].z + a[r][v].w * a[r][v].w);
      }
      rs[r] = rsqrtf(group_sum8(q) * invC + eps);
      if (ok && t == 0) {
        mean_out[row] = mu;
        rstd_out[row] = rs[r];
      }
#pragma unroll
      for (int k = 0; k < KP; ++k) acc[r][k] = 0.f;
    }
#pragma unroll
    for (int v = 0; v < V; ++v)
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        const float4 w4 = wg[k][t + kT * v];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r][k] = dot4(a[r][v], w4, acc[r][k]);
      }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float r1[H0], r2[H1], r3[H2];
      reduce_scatter_step<KP>(acc[r], r1, up4, 4);
      reduce_scatter_step<H0>(r1, r2, up2, 2);
      reduce_scatter_step<H1>(r2, r3, up1, 1);
      const long long row = base + 4 * r + sub;
      if (row < rows) {
        long long b, p;
        split_row(row, P, small, b, p);
        float* out = logits + (b * K) * P + p;
#pragma unroll
        for (int i = 0; i < H2; ++i)
          if (i < nmine && kmine + i < K) out[(long long)(kmine + i) * P] = fmaf(rs[r], r3[i], b0[kmine + i]);
      }
    }
  }
}

constexpr int kThreadsB = 192;  // backward: 2 x 6 warps per SM leave 170 registers per thread for the prefetch

template <int V, int KQ, int R>
__global__ void __launch_bounds__(kThreadsB, 2)
ln_head_bwd_kernel(const float* __restrict__ dlogits, const float4* __restrict__ x, const float* __restrict__ mean,
                   const float* __restrict__ rstd, const float4* __restrict__ gamma, const float* __restrict__ w,
                   float4* __restrict__ dx, float* __restrict__ s_acc, float* __restrict__ g_acc, long long rows,
                   long long P, int K) {
  constexpr int KP = 4 * KQ, C4 = kT * V, C = 4 * C4, NR = 4 * R, kWarps = kThreadsB / 32;
  __shared__ float4 ws[KP][C4];             // W, zero rows for k >= K
  __shared__ float4 xs[kWarps][NR][C4];     // xhat of the warp's rows
  __shared__ float4 g_rk[kWarps][NR][KQ];   // d(logits) [row][k]
  __shared__ float4 g_kr[kWarps][KP][R];    // d(logits) [k][row]
  __shared__ float red_s[KP * C];
  __shared__ float red_g[KP];
  for (int i = threadIdx.x; i < KP * C4; i += kThreadsB) {
    const int k = i / C4, c4 = i - k * C4;
    ws[k][c4] = k < K ? __ldg(reinterpret_cast<const float4*>(w) + (size_t)k * C4 + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int i = threadIdx.x; i < KP * C; i += kThreadsB) red_s[i] = 0.f;
  if (threadIdx.x < KP) red_g[threadIdx.x] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, t = lane & (kT - 1), sub = lane >> 3, wid = threadIdx.x >> 5;
  const float invC = 1.0f / (float)C;
  const bool small = rows < (1LL << 31);
  float4 S[KQ][V];
  float G[KQ];
#pragma unroll
  for (int j = 0; j < KQ; ++j) {
    G[j] = 0.f;
#pragma unroll
    for (int v = 0; v < V; ++v) S[j][v] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float4 gam[V];
#pragma unroll
  for (int v = 0; v < V; ++v) gam[v] = __ldg(gamma + t + kT * v);
  float* grk = reinterpret_cast<float*>(&g_rk[wid][0][0]);  // [NR][KP]
  float* gkr = reinterpret_cast<float*>(&g_kr[wid][0][0]);  // [KP][NR]
  const long long warp0 = (long long)blockIdx.x * kWarps + wid;
  const long long stride = (long long)gridDim.x * kWarps * NR;
  const long long iters = (rows + stride - 1) / stride;  // uniform trip count: see the forward kernel
  // software pipeline (see the forward kernel): x, the statistics and d(logits) of iteration it + 1 are requested before
  // iteration it is processed
  float4 nx[R][V];
  float nmu[R], nrs[R], ng[R][(KP + kT - 1) / kT];
  auto fetch = [&](long long base) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long row = base + 4 * r + sub;
      const bool ok = row < rows;
      nmu[r] = ok ? __ldg(mean + row) : 0.f;
      nrs[r] = ok ? __ldg(rstd + row) : 0.f;
#pragma unroll
      for (int v = 0; v < V; ++v) nx[r][v] = ok ? __ldcs(x + row * C4 + t + kT * v) : make_float4(0.f, 0.f, 0.f, 0.f);
      long long b, p;
      split_row(ok ? row : 0, P, small, b, p);
      const float* gp = dlogits + (b * K) * P + p;
#pragma unroll
      for (int jj = 0; jj < (KP + kT - 1) / kT; ++jj) {  // lane t fetches classes t, t + 8
        const int k = t + kT * jj;
        ng[r][jj] = (ok && k < K) ? __ldcs(gp + (long long)k * P) : 0.f;
      }
    }
  };
  fetch(warp0 * NR);
  for (long long it = 0; it < iters; ++it) {
    const long long base = warp0 * NR + it * stride;
    float4 xh[R][V];
    float rs[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float mu = nmu[r];
      rs[r] = nrs[r];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const float4 xv = nx[r][v];
        xh[r][v] = make_float4((xv.x - mu) * rs[r], (xv.y - mu) * rs[r], (xv.z - mu) * rs[r], (xv.w - mu) * rs[r]);
        xs[wid][4 * r + sub][t + kT * v] = xh[r][v];
      }
#pragma unroll
      for (int jj = 0; jj < (KP + kT - 1) / kT; ++jj) {
        const int k = t + kT * jj;
        if (k < KP) {
          grk[(4 * r + sub) * KP + k] = ng[r][jj];
          gkr[k * NR + 4 * r + sub] = ng[r][jj];
        }
      }
    }
    if (it + 1 < iters) fetch(base + stride);
    __syncwarp();
    // d(LN output) of the own rows, LayerNorm backward, dx
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float gk[KP];
#pragma unroll
      for (int j = 0; j < KQ; ++j) {
        const float4 g4 = g_rk[wid][4 * r + sub][j];
        gk[4 * j + 0] = g4.x; gk[4 * j + 1] = g4.y; gk[4 * j + 2] = g4.z; gk[4 * j + 3] = g4.w;
      }
      float4 wv[V];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < KP; ++k) {
          const float4 w4 = ws[k][t + kT * v];
          d.x = fmaf(gk[k], w4.x, d.x); d.y = fmaf(gk[k], w4.y, d.y);
          d.z = fmaf(gk[k], w4.z, d.z); d.w = fmaf(gk[k], w4.w, d.w);
        }
        wv[v] = make_float4(d.x * gam[v].x, d.y * gam[v].y, d.z * gam[v].z, d.w * gam[v].w);
        s1 += (wv[v].x + wv[v].y) + (wv[v].z + wv[v].w);
        s2 += (wv[v].x * xh[r][v].x + wv[v].y * xh[r][v].y) + (wv[v].z * xh[r][v].z + wv[v].w * xh[r][v].w);
      }
      s1 = group_sum8(s1) * invC;
      s2 = group_sum8(s2) * invC;
      const long long row = base + 4 * r + sub;
      if (row < rows) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          float4 o;
          o.x = rs[r] * (wv[v].x - s1 - xh[r][v].x * s2);
          o.y = rs[r] * (wv[v].y - s1 - xh[r][v].y * s2);
          o.z = rs[r] * (wv[v].z - s1 - xh[r][v].z * s2);
          o.w = rs[r] * (wv[v].w - s1 - xh[r][v].w * s2);
          dx[row * C4 + t + kT * v] = o;
        }
      }
    }
    // S[k][c] += g[row][k] * xhat[row][c] over the warp's NR rows; this lane: classes sub + 4 j, channels of t
    float gq[KQ][NR];
#pragma unroll
    for (int j = 0; j < KQ; ++j)
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float4 g4 = g_kr[wid][sub + 4 * j][r];
        gq[j][4 * r + 0] = g4.x; gq[j][4 * r + 1] = g4.y; gq[j][4 * r + 2] = g4.z; gq[j][4 * r + 3] = g4.w;
      }
#pragma unroll
    for (int rr = 0; rr < NR; ++rr) {
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const float4 x4 = xs[wid][rr][t + kT * v];
#pragma unroll
        for (int j = 0; j < KQ; ++j) {
          S[j][v].x = fmaf(gq[j][rr], x4.x, S[j][v].x); S[j][v].y = fmaf(gq[j][rr], x4.y, S[j][v].y);
          S[j][v].z = fmaf(gq[j][rr], x4.z, S[j][v].z); S[j][v].w = fmaf(gq[j][rr], x4.w, S[j][v].w);
        }
      }
#pragma unroll
      for (int j = 0; j < KQ; ++j) G[j] += gq[j][rr];
    }
    __syncwarp();  // the exchange buffers are rewritten by the next iteration
  }
#pragma unroll
  for (int j = 0; j < KQ; ++j) {
    const int k = sub + 4 * j;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const int c = 4 * (t + kT * v);
      atomicAdd(red_s + k * C + c + 0, S[j][v].x); atomicAdd(red_s + k * C + c + 1, S[j][v].y);
      atomicAdd(red_s + k * C + c + 2, S[j][v].z); atomicAdd(red_s + k * C + c + 3, S[j][v].w);
    }
    if (t == 0) atomicAdd(red_g + k, G[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * C; i += kThreadsB) atomicAdd(s_acc + i, red_s[i]);
  if (threadIdx.x < K) atomicAdd(g_acc + threadIdx.x, red_g[threadIdx.x]);
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

constexpr int kRF = 2, kThreadsF = 192;  // forward: rows per lane and iteration, block size (2 x 6 warps per SM: 170 registers)

constexpr int kRB = 1;  // backward: rows per lane and iteration (2 needs 255 registers and was slower)

#define HS_LH_DISPATCH_KQ(V_, ...)                                         \
  switch ((K + 3) / 4) {                                                   \
    case 1: { constexpr int V = V_, KQ = 1; __VA_ARGS__; break; }          \
    case 2: { constexpr int V = V_, KQ = 2; __VA_ARGS__; break; }          \
    case 3: { constexpr int V = V_, KQ = 3; __VA_ARGS__; break; }          \
    default: { constexpr int V = V_, KQ = 4; __VA_ARGS__; break; }         \
  }
#define HS_LH_DISPATCH(...)                                                \
  switch (C / 32) {                                                        \
    case 1: HS_LH_DISPATCH_KQ(1, __VA_ARGS__) break;                       \
    case 2: HS_LH_DISPATCH_KQ(2, __VA_ARGS__) break;                       \
    case 3: HS_LH_DISPATCH_KQ(3, __VA_ARGS__) break;                       \
    default: HS_LH_DISPATCH_KQ(4, __VA_ARGS__) break;                      \
  }

}  // namespace

extern "C" {

int hs_ln_head_supported(int64_t rows, int C, int K) {
  return (rows > 0 && (C == 32 || C == 64 || C == 96 || C == 128) && K >= 1 && K <= 16) ? 1 : 0;
}

int hs_ln_head_fwd(const float* x, const float* gamma, const float* beta, const float* w, const float* head_bias,
                   float* logits, float* mean, float* rstd, int64_t rows, int64_t rows_per_sample, int C, int K, float eps,
                   void* stream) {
  HS_REQUIRE(x && gamma && beta && w && logits && mean && rstd, "hs_ln_head_fwd: null pointer");
  HS_REQUIRE(rows_per_sample > 0 && rows % rows_per_sample == 0, "hs_ln_head_fwd: rows (%lld) is not a multiple of "
             "rows_per_sample (%lld)", (long long)rows, (long long)rows_per_sample);
  if (!hs_ln_head_supported(rows, C, K))
    return hs::fail(HS_ERR_UNSUPPORTED, "hs_ln_head_fwd: C=%d K=%d is not covered (C in {32, 64, 96, 128}, K <= 16)", C, K);
  HS_REQUIRE(aligned16(x) && aligned16(gamma) && aligned16(w), "hs_ln_head_fwd: x, gamma and w must be 16-byte aligned");
  const long long per_block = (kThreadsF / 32) * 4 * kRF;
  long long blocks = (rows + per_block - 1) / per_block;
  const long long cap = (long long)num_sms() * 2;
  if (blocks > cap) blocks = cap;
  HS_LH_DISPATCH((ln_head_fwd_kernel<V, KQ, kRF, kThreadsF><<<(unsigned)blocks, kThreadsF, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(gamma), beta, w, head_bias, logits, mean, rstd,
      rows, rows_per_sample, K, eps)));
  HS_LAUNCH_CHECK();
  return HS_OK;
}

int hs_ln_head_bwd(const float* dlogits, const float* x, const float* mean, const float* rstd, const float* gamma,
                   const float* w, float* dx, float* s_acc, float* g_acc, int64_t rows, int64_t rows_per_sample, int C,
                   int K, void* stream) {
  HS_REQUIRE(dlogits && x && mean && rstd && gamma && w && dx && s_acc && g_acc, "hs_ln_head_bwd: null pointer");
  HS_REQUIRE(rows_per_sample > 0 && rows % rows_per_sample == 0, "hs_ln_head_bwd: rows (%lld) is not a multiple of "
             "rows_per_sample (%lld)", (long long)rows, (long long)rows_per_sample);
  if (!hs_ln_head_supported(rows, C, K))
    return hs::fail(HS_ERR_UNSUPPORTED, "hs_ln_head_bwd: C=%d K=%d is not covered (C in {32, 64, 96, 128}, K <= 16)", C, K);
  HS_REQUIRE(aligned16(x) && aligned16(gamma) && aligned16(w) && aligned16(dx),
             "hs_ln_head_bwd: x, gamma, w and dx must be 16-byte aligned");
  const long long per_block = (kThreadsB / 32) * 4 * kRB;
  long long blocks = (rows + per_block - 1) / per_block;
  const long long cap = (long long)num_sms() * 2;
  if (blocks > cap) blocks = cap;
  HS_LH_DISPATCH((ln_head_bwd_kernel<V, KQ, kRB><<<(unsigned)blocks, kThreadsB, 0, (cudaStream_t)stream>>>(
      dlogits, reinterpret_cast<const float4*>(x), mean, rstd, reinterpret_cast<const float4*>(gamma), w,
      reinterpret_cast<float4*>(dx), s_acc, g_acc, rows, rows_per_sample, K)));
  HS_LAUNCH_CHECK();
  return HS_OK;
}

}  // extern "C"
