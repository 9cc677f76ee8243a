// Generic (any window size dividing N -- the power-of-two rule of hp_windowing.py:16 is enforced by the HEALPix
// modules, the flat twin has e.g. 4x6 or 7x7 windows -- and any head_dim) fp32 CUDA-core implementation of the windowed
// attention core with the HEALPix shift / window partition / reverse folded into its loads and
// stores.  This is the exact-arithmetic path used for shapes the tcgen05 kernel does not cover
// (e.g. BASELINE config 1: window 16, head_dim 16; the reference's own test config: window 4,
// head_dim 2) and the cross-check for the tensor-core kernel.
//
// Replaces swin_hp_transformer.py:136-171 (+ :319-330 shift/partition/reverse) -- see
// include/healswin_b200.h for the exact contract.
#include <cfloat>

#include "hs_common.h"
#include "hs_kernels.h"

namespace {

constexpr int kThreads = 256;
constexpr float kLogitScaleMax = 4.605170185988092f;  // log(1/0.01), swin_hp_transformer.py:144-146
constexpr float kNormEps = 1e-12f;                    // F.normalize eps
constexpr float kMaskFill = -100.0f;                  // hp_shifting.py:25

struct AttnArgs {
  const float* qkv;
  const float* dout;  // bwd only
  const int32_t* src;
  const uint8_t* groups;
  const float* mask;
  const float* bias;
  const float* logit_scale;
  float scale;
  float* out;    // fwd
  float* lse;    // fwd, optional: (H, B*N) log2-domain log-sum-exp per (head, token)
  float* dqkv;   // bwd
  float* dbias;  // bwd
  float* dlogit; // bwd
  int B;
  long long N;
  int C, H, ws, D;
  int nW;        // windows per sample
  int cos;
  uint32_t drop_thresh;  // 0 = no attention dropout
  float drop_scale;      // 1 / (1 - p)
  uint64_t seed;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// shared-memory carve-up (floats): Q[ws][D+1] K[ws][D+1] V[ws][D+1] S[ws][ws+1] rows[ws] (int) grp[ws]
struct Smem {
  float *q, *k, *v, *s;
  int* row;
  int* grp;
  float* inv;  // [2*ws] 1/max(|q|,eps), 1/max(|k|,eps) (cos only)
  int ldq, lds;
};

__device__ __forceinline__ Smem carve(float* base, int ws, int D) {
  Smem m;
  m.ldq = D + 1;
  m.lds = ws + 1;
  m.q = base;
  m.k = m.q + ws * m.ldq;
  m.v = m.k + ws * m.ldq;
  m.s = m.v + ws * m.ldq;
  m.inv = m.s + ws * m.lds;
  m.row = reinterpret_cast<int*>(m.inv + 2 * ws);
  m.grp = m.row + ws;
  return m;
}

size_t smem_bytes_fwd(int ws, int D) { return sizeof(float) * (3 * ws * (D + 1) + ws * (ws + 1) + 2 * ws) + 2 * sizeof(int) * ws; }

// loads the (window, head) tile; normalises q/k rows for cos attention and folds the logit scale
// (or the dot-product scale) into q.  Returns the effective scale that was folded into q.
__device__ float load_tile(const AttnArgs& a, const Smem& m, long long wb, int h, bool keep_raw_q) {
  const int ws = a.ws, D = a.D;
  const int b = (int)(wb / a.nW), w = (int)(wb % a.nW);
  for (int j = threadIdx.x; j < ws; j += blockDim.x) {
    const long long slot = (long long)w * ws + j;
    m.row[j] = a.src ? a.src[slot] : (int)slot;
    m.grp[j] = a.groups ? (int)a.groups[slot] : 0;
  }
  __syncthreads();
  const float* base = a.qkv + (long long)b * a.N * 3 * a.C + (long long)h * D;
  for (int idx = threadIdx.x; idx < ws * D; idx += blockDim.x) {
    const int j = idx / D, dd = idx - j * D;
    const float* r = base + (long long)m.row[j] * 3 * a.C + dd;
    m.q[j * m.ldq + dd] = r[0];
    m.k[j * m.ldq + dd] = r[a.C];
    m.v[j * m.ldq + dd] = r[2 * a.C];
  }
  __syncthreads();
  float eff = a.scale;
  if (a.cos) {
    eff = expf(fminf(a.logit_scale[h], kLogitScaleMax));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    for (int r = warp; r < 2 * ws; r += nwarp) {
      const float* p = (r < ws) ? (m.q + r * m.ldq) : (m.k + (r - ws) * m.ldq);
      float ss = 0.f;
      for (int dd = lane; dd < D; dd += 32) ss += p[dd] * p[dd];
      ss = warp_sum(ss);
      if (lane == 0) {
        const float iv = 1.0f / fmaxf(sqrtf(ss), kNormEps);
        m.inv[r] = iv;
        if (a.lse) {  // planes 1 (1/|q|) and 2 (1/|k|) of the statistics buffer (cos attention)
          const long long plane = (long long)a.H * a.B * a.N;
          const int j = r < ws ? r : r - ws;
          a.lse[(r < ws ? 1 : 2) * plane + ((long long)h * a.B + b) * a.N + m.row[j]] = iv;
        }
      }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < ws * D; idx += blockDim.x) {
      const int j = idx / D, dd = idx - j * D;
      m.q[j * m.ldq + dd] *= m.inv[j] * (keep_raw_q ? 1.0f : eff);
      m.k[j * m.ldq + dd] *= m.inv[ws + j];
    }
  } else if (!keep_raw_q) {
    for (int idx = threadIdx.x; idx < ws * D; idx += blockDim.x) {
      const int j = idx / D, dd = idx - j * D;
      m.q[j * m.ldq + dd] *= eff;
    }
  }
  __syncthreads();
  return eff;
}

// S = (q_eff k^T) * mult + bias + mask, then row softmax in place.  4x4 register tiles when ws % 4 == 0.
__device__ void logits_softmax(const AttnArgs& a, const Smem& m, long long wb, int h, float mult,
                               float* lse_out = nullptr /* base of this (head, sample): indexed by token */,
                               bool drop_in_place = false) {
  const int ws = a.ws, D = a.D;
  const int w = (int)(wb % a.nW);
  const float* bias = a.bias ? a.bias + (long long)h * ws * ws : nullptr;
  const float* mask = a.mask ? a.mask + (long long)w * ws * ws : nullptr;
  if (ws % 4 == 0) {
    const int T = ws / 4;
    for (int tile = threadIdx.x; tile < T * T; tile += blockDim.x) {
      const int ti = tile / T, tj = tile - ti * T;
      float acc[4][4];
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = 0.f;
      for (int dd = 0; dd < D; ++dd) {
        float qa[4], kb[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) qa[x] = m.q[(ti * 4 + x) * m.ldq + dd];
#pragma unroll
        for (int y = 0; y < 4; ++y) kb[y] = m.k[(tj + y * T) * m.ldq + dd];
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
          for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(qa[x], kb[y], acc[x][y]);
      }
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) {
          const int i = ti * 4 + x, j = tj + y * T;
          float v = acc[x][y] * mult;
          if (bias) v += bias[i * ws + j];
          if (m.grp[i] != m.grp[j]) v += kMaskFill;
          if (mask) v += mask[i * ws + j];
          m.s[i * m.lds + j] = v;
        }
    }
  } else {
    for (int idx = threadIdx.x; idx < ws * ws; idx += blockDim.x) {
      const int i = idx / ws, j = idx - i * ws;
      float v = 0.f;
      for (int dd = 0; dd < D; ++dd) v = fmaf(m.q[i * m.ldq + dd], m.k[j * m.ldq + dd], v);
      v *= mult;
      if (bias) v += bias[i * ws + j];
      if (m.grp[i] != m.grp[j]) v += kMaskFill;
      if (mask) v += mask[i * ws + j];
      m.s[i * m.lds + j] = v;
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int i = warp; i < ws; i += nwarp) {
    float* r = m.s + i * m.lds;
    float mx = -FLT_MAX;
    for (int j = lane; j < ws; j += 32) mx = fmaxf(mx, r[j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < ws; j += 32) {
      const float e = expf(r[j] - mx);
      r[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int j = lane; j < ws; j += 32) r[j] *= inv;
    if (lse_out && lane == 0) lse_out[m.row[i]] = (mx + logf(sum)) * 1.4426950408889634f;
    if (drop_in_place) {  // forward: attn = dropout(softmax(.))   [swin_hp_transformer.py:165-169]
      const uint32_t key = hs::drop_unit_key(a.seed, wb, h, a.H);
      for (int j = lane; j < ws; j += 32) r[j] = hs::drop_keep(key, i, j, ws, a.drop_thresh) ? r[j] * a.drop_scale : 0.f;
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kThreads) attn_fwd_kernel(AttnArgs a) {
  extern __shared__ float smem_f[];
  const Smem m = carve(smem_f, a.ws, a.D);
  const int h = blockIdx.y;
  const long long total = (long long)a.B * a.nW;
  const int ws = a.ws, D = a.D;
  for (long long wb = blockIdx.x; wb < total; wb += gridDim.x) {
    load_tile(a, m, wb, h, false);
    const int b = (int)(wb / a.nW);
    logits_softmax(a, m, wb, h, 1.0f, a.lse ? a.lse + ((long long)h * a.B + b) * a.N : nullptr, a.drop_thresh != 0);
    float* obase = a.out + (long long)b * a.N * a.C + (long long)h * D;
    for (int idx = threadIdx.x; idx < ws * D; idx += blockDim.x) {
      const int i = idx / D, dd = idx - i * D;
      const float* p = m.s + i * m.lds;
      float acc = 0.f;
      for (int j = 0; j < ws; ++j) acc = fmaf(p[j], m.v[j * m.ldq + dd], acc);
      obase[(long long)m.row[i] * a.C + dd] = acc;
    }
    __syncthreads();
  }
}

// backward shared memory: fwd layout + dO[ws][D+1] + dS accumulator for dbias [ws][ws] (optional)
size_t smem_bytes_bwd(int ws, int D, bool with_dbias) {
  return smem_bytes_fwd(ws, D) + sizeof(float) * (ws * (D + 1) + (with_dbias ? ws * ws : 0) + 32);
}

__global__ void __launch_bounds__(kThreads) attn_bwd_kernel(AttnArgs a) {
  extern __shared__ float smem_f[];
  const Smem m = carve(smem_f, a.ws, a.D);
  const int ws = a.ws, D = a.D;
  float* dO = reinterpret_cast<float*>(m.grp + ws);
  float* red = dO + ws * m.ldq;            // 32 floats of block-reduce scratch
  float* dB = a.dbias ? red + 32 : nullptr;  // [ws][ws] running sum of dS over this CTA's windows
  const int h = blockIdx.y;
  const long long total = (long long)a.B * a.nW;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  if (dB)
    for (int idx = threadIdx.x; idx < ws * ws; idx += blockDim.x) dB[idx] = 0.f;
  float dscale_acc = 0.f;  // per-thread partial of sum dS * (q_hat . k_hat)
  for (long long wb = blockIdx.x; wb < total; wb += gridDim.x) {
    // q_hat (unscaled), k_hat, v in smem; eff = logit scale (cos) or dot scale
    const float eff = load_tile(a, m, wb, h, true);
    const int b = (int)(wb / a.nW);
    const float* dob = a.dout + (long long)b * a.N * a.C + (long long)h * D;
    for (int idx = threadIdx.x; idx < ws * D; idx += blockDim.x) {
      const int i = idx / D, dd = idx - i * D;
      dO[i * m.ldq + dd] = dob[(long long)m.row[i] * a.C + dd];
    }
    logits_softmax(a, m, wb, h, eff);  // m.s = P  (syncs inside also publish dO)
    const uint32_t dkey = hs::drop_unit_key(a.seed, wb, h, a.H);
    auto dmask = [&](int i, int j) -> float {  // dropout multiplier of P[i][j]: 0 or 1 / (1 - p)
      return (a.drop_thresh == 0) ? 1.0f : (hs::drop_keep(dkey, i, j, ws, a.drop_thresh) ? a.drop_scale : 0.f);
    };
    float* dqkv_b = a.dqkv + (long long)b * a.N * 3 * a.C + (long long)h * D;
    // dV[j] = sum_i P[i][j] dO[i]
    for (int idx = threadIdx.x; idx < ws * D; idx += blockDim.x) {
      const int j = idx / D, dd = idx - j * D;
      float acc = 0.f;
      for (int i = 0; i < ws; ++i) acc = fmaf(m.s[i * m.lds + j] * dmask(i, j), dO[i * m.ldq + dd], acc);
      dqkv_b[(long long)m.row[j] * 3 * a.C + 2 * a.C + dd] = acc;
    }
    __syncthreads();
    // dS = P * (dP - rowsum(P*dP)), dP[i][j] = dO[i].V[j]; one warp per row; overwrite m.s with dS
    for (int i = warp; i < ws; i += nwarp) {
      float* pr = m.s + i * m.lds;
      float dot = 0.f;
      for (int j = lane; j < ws; j += 32) {
        float dp = 0.f;
        for (int dd = 0; dd < D; ++dd) dp = fmaf(dO[i * m.ldq + dd], m.v[j * m.ldq + dd], dp);
        dot = fmaf(pr[j], dp * dmask(i, j), dot);
      }
      dot = warp_sum(dot);
      for (int j = lane; j < ws; j += 32) {
        float dp = 0.f;
        for (int dd = 0; dd < D; ++dd) dp = fmaf(dO[i * m.ldq + dd], m.v[j * m.ldq + dd], dp);
        const float ds = pr[j] * (dp * dmask(i, j) - dot);
        pr[j] = ds;
        if (dB) dB[i * ws + j] += ds;  // (i, j) owned by exactly one thread of this CTA
      }
    }
    __syncthreads();
    // d q_hat[i] = eff * sum_j dS[i][j] k_hat[j];  d k_hat[j] = eff * sum_i dS[i][j] q_hat[i]
    // stash results in dO (dq) and v (dk): both are dead now.
    for (int idx = threadIdx.x; idx < ws * D; idx += blockDim.x) {
      const int i = idx / D, dd = idx - i * D;
      float aq = 0.f, ak = 0.f;
      for (int j = 0; j < ws; ++j) {
        aq = fmaf(m.s[i * m.lds + j], m.k[j * m.ldq + dd], aq);
        ak = fmaf(m.s[j * m.lds + i], m.q[j * m.ldq + dd], ak);
      }
      dO[i * m.ldq + dd] = aq * eff;
      m.v[i * m.ldq + dd] = ak * eff;
    }
    __syncthreads();
    if (a.cos) {
      // through F.normalize: dq = (dqh - qh (qh.dqh)) / max(|q|, eps); also dscale += qh . dqh / eff
      for (int r = warp; r < 2 * ws; r += nwarp) {
        const bool isq = r < ws;
        const int i = isq ? r : r - ws;
        const float* hat = (isq ? m.q : m.k) + i * m.ldq;
        float* g = (isq ? dO : m.v) + i * m.ldq;
        float dot = 0.f;
        for (int dd = lane; dd < D; dd += 32) dot = fmaf(hat[dd], g[dd], dot);
        dot = warp_sum(dot);
        if (isq && lane == 0) dscale_acc += dot / eff;  // sum_j dS[i][j] cos_ij
        const float inv = m.inv[r];
        // F.normalize divides by max(|x|, eps): when |x| < eps the denominator is the constant eps
        const bool clamped = inv >= 1.0f / kNormEps;
        for (int dd = lane; dd < D; dd += 32) g[dd] = (g[dd] - (clamped ? 0.f : hat[dd] * dot)) * inv;
      }
      __syncthreads();
    }
    for (int idx = threadIdx.x; idx < ws * D; idx += blockDim.x) {
      const int i = idx / D, dd = idx - i * D;
      float* r = dqkv_b + (long long)m.row[i] * 3 * a.C + dd;
      r[0] = dO[i * m.ldq + dd];
      r[a.C] = m.v[i * m.ldq + dd];
    }
    __syncthreads();
  }
  if (dB) {
    float* g = a.dbias + (long long)h * ws * ws;
    for (int idx = threadIdx.x; idx < ws * ws; idx += blockDim.x) atomicAdd(g + idx, dB[idx]);
  }
  if (a.cos && a.dlogit) {
    // d logit_scale = dscale * scale, zero where the clamp is active (torch.clamp backward)
    float v = warp_sum(dscale_acc);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int i = 0; i < nwarp; ++i) t += red[i];
      const float ls = a.logit_scale[h];
      if (ls <= kLogitScaleMax) atomicAdd(a.dlogit + h, t * expf(ls));
    }
  }
}

__global__ void bias_expand_kernel(const float* table, const int32_t* index, float* bias, int T, int H, int n2) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= H * n2) return;
  const int h = idx / n2, ij = idx - h * n2;
  bias[idx] = table[(long long)index[ij] * H + h];
}
__global__ void bias_reduce_kernel(const float* dbias, const int32_t* index, float* dtable, int T, int H, int n2) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= H * n2) return;
  const int h = idx / n2, ij = idx - h * n2;
  atomicAdd(dtable + (long long)index[ij] * H + h, dbias[idx]);
}

int validate(const char* fn, const void* qkv, int B, int64_t N, int C, int H, int ws) {
  HS_REQUIRE(qkv != nullptr, "%s: null qkv", fn);
  HS_REQUIRE(B > 0 && N > 0 && C > 0 && H > 0 && ws > 0, "%s: non-positive dimension", fn);
  HS_REQUIRE(C % H == 0, "%s: dim %d not divisible by num_heads %d", fn, C, H);
  HS_REQUIRE(N % ws == 0, "%s: window_size %d does not divide N=%lld", fn, ws, (long long)N);
  HS_REQUIRE(N * 3 * C < (int64_t)1 << 40, "%s: tensor too large", fn);
  return HS_OK;
}

int pick_grid_x(long long total, int H, int num_sms, int ctas_per_sm) {
  long long want = ((long long)num_sms * ctas_per_sm + H - 1) / H;
  if (want < 1) want = 1;
  return (int)(total < want ? total : want);
}

int sm_count() {
  // per device (a process may drive several GPUs): an immutable cache, filled on first use
  static int cached[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return 148;
  if (!cached[dev]) cudaDeviceGetAttribute(&cached[dev], cudaDevAttrMultiProcessorCount, dev);
  return cached[dev] ? cached[dev] : 148;
}

}  // namespace

namespace hs {

int window_attn_fwd_simt(const float* qkv, const int32_t* src, const uint8_t* groups, const float* mask,
                         const float* bias, const float* logit_scale, float scale, DropCfg drop, float* out, float* lse,
                         int B, int64_t N, int C, int H, int ws, uint32_t flags, cudaStream_t stream) {
  int rc = validate("hs_window_attn_fwd", qkv, B, N, C, H, ws);
  if (rc) return rc;
  HS_REQUIRE(out != nullptr, "hs_window_attn_fwd: null out");
  HS_REQUIRE(!(flags & HS_ATTN_COS) || logit_scale, "hs_window_attn_fwd: cos attention needs logit_scale");
  AttnArgs a{};
  a.qkv = qkv; a.src = src; a.groups = groups; a.mask = mask; a.bias = bias; a.logit_scale = logit_scale;
  a.scale = scale; a.out = out; a.lse = lse; a.B = B; a.N = N; a.C = C; a.H = H; a.ws = ws; a.D = C / H;
  a.nW = (int)(N / ws); a.cos = (flags & HS_ATTN_COS) ? 1 : 0;
  a.drop_thresh = drop.p > 0.f ? hs::drop_thresh(drop.p) : 0u; a.drop_scale = 1.0f / (1.0f - drop.p); a.seed = drop.seed;
  const size_t smem = smem_bytes_fwd(ws, a.D);
  if (smem > 200 * 1024)
    return hs::fail(HS_ERR_UNSUPPORTED, "hs_window_attn_fwd: window %d x head_dim %d needs %zu B of shared memory", ws, a.D, smem);
  HS_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long total = (long long)B * a.nW;
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm > 8) per_sm = 8;
  if (per_sm < 1) per_sm = 1;
  dim3 grid(pick_grid_x(total, H, sm_count(), per_sm * 4), H);
  attn_fwd_kernel<<<grid, kThreads, smem, stream>>>(a);
  HS_LAUNCH_CHECK();
  return HS_OK;
}

int window_attn_bwd_simt(const float* qkv, const float* dout, const int32_t* src, const uint8_t* groups,
                         const float* mask, const float* bias, const float* logit_scale, float scale, DropCfg drop,
                         float* dqkv, float* dbias, float* dlogit, int B, int64_t N, int C, int H, int ws,
                         uint32_t flags, cudaStream_t stream) {
  int rc = validate("hs_window_attn_bwd", qkv, B, N, C, H, ws);
  if (rc) return rc;
  HS_REQUIRE(dout && dqkv, "hs_window_attn_bwd: null dout/dqkv");
  HS_REQUIRE(!(flags & HS_ATTN_COS) || logit_scale, "hs_window_attn_bwd: cos attention needs logit_scale");
  AttnArgs a{};
  a.qkv = qkv; a.dout = dout; a.src = src; a.groups = groups; a.mask = mask; a.bias = bias;
  a.logit_scale = logit_scale; a.scale = scale; a.dqkv = dqkv; a.dbias = dbias; a.dlogit = dlogit;
  a.B = B; a.N = N; a.C = C; a.H = H; a.ws = ws; a.D = C / H; a.nW = (int)(N / ws);
  a.cos = (flags & HS_ATTN_COS) ? 1 : 0;
  a.drop_thresh = drop.p > 0.f ? hs::drop_thresh(drop.p) : 0u; a.drop_scale = 1.0f / (1.0f - drop.p); a.seed = drop.seed;
  const size_t smem = smem_bytes_bwd(ws, a.D, dbias != nullptr);
  if (smem > 200 * 1024)
    return hs::fail(HS_ERR_UNSUPPORTED, "hs_window_attn_bwd: window %d x head_dim %d needs %zu B of shared memory", ws, a.D, smem);
  HS_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long total = (long long)B * a.nW;
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm > 8) per_sm = 8;
  if (per_sm < 1) per_sm = 1;
  // persistent CTAs: each keeps a private dbias accumulator, so the grid is bounded by the SM count
  dim3 grid(pick_grid_x(total, H, sm_count(), per_sm), H);
  attn_bwd_kernel<<<grid, kThreads, smem, stream>>>(a);
  HS_LAUNCH_CHECK();
  return HS_OK;
}

}  // namespace hs

extern "C" {

int hs_rel_bias_expand(const float* table, const int32_t* index, float* bias, int T, int H, int ws, void* stream) {
  HS_REQUIRE(table && index && bias && T > 0 && H > 0 && ws > 0, "hs_rel_bias_expand: bad arguments");
  const int n = H * ws * ws;
  bias_expand_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(table, index, bias, T, H, ws * ws);
  HS_LAUNCH_CHECK();
  return HS_OK;
}

int hs_rel_bias_reduce(const float* dbias, const int32_t* index, float* dtable, int T, int H, int ws, void* stream) {
  HS_REQUIRE(dbias && index && dtable && T > 0 && H > 0 && ws > 0, "hs_rel_bias_reduce: bad arguments");
  const int n = H * ws * ws;
  bias_reduce_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(dbias, index, dtable, T, H, ws * ws);
  HS_LAUNCH_CHECK();
  return HS_OK;
}

int hs_device_info(int* sm_major, int* sm_minor, int* num_sms, char* name, int name_len) {
  int dev = 0;
  HS_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  HS_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm_major) *sm_major = p.major;
  if (sm_minor) *sm_minor = p.minor;
  if (num_sms) *num_sms = p.multiProcessorCount;
  if (name && name_len > 0) {
    strncpy(name, p.name, (size_t)name_len - 1);
    name[name_len - 1] = 0;
  }
  if (p.major != 10) return hs::fail(HS_ERR_UNSUPPORTED, "device %s is sm_%d%d, this library is built for sm_100a only", p.name, p.major, p.minor);
  return HS_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Row gather: out[b, p, :] = x[b, idx[p], :]  -- the standalone form of shifter.shift / shift_back
// (hp_shifting.py:69-73, 302-306, 400-404) for callers that use the shifter objects directly.
// 128-bit vectorised, one warp-contiguous row segment per iteration; HBM-bound.
namespace {
__global__ void gather_rows_kernel(const float4* __restrict__ x, const int32_t* __restrict__ idx,
                                   float4* __restrict__ out, int B, long long N, int C4) {
  const long long total = (long long)B * N * C4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / C4;
    const int c = (int)(i - row * C4);
    const long long b = row / N, p = row - b * N;
    out[i] = x[(b * N + idx[p]) * C4 + c];
  }
}
__global__ void gather_rows_scalar_kernel(const float* __restrict__ x, const int32_t* __restrict__ idx,
                                          float* __restrict__ out, int B, long long N, int C) {
  const long long total = (long long)B * N * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / C;
    const int c = (int)(i - row * C);
    const long long b = row / N, p = row - b * N;
    out[i] = x[(b * N + idx[p]) * C + c];
  }
}
}  // namespace

extern "C" int hs_gather_rows(const float* x, const int32_t* idx, float* out, int B, int64_t N, int C, void* stream) {
  HS_REQUIRE(x && idx && out && B > 0 && N > 0 && C > 0, "hs_gather_rows: bad arguments");
  const long long total = (long long)B * N * C;
  const bool vec = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) % 16 == 0);
  const long long work = vec ? total / 4 : total;
  long long blocks = (work + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (vec)
    gather_rows_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(x), idx, reinterpret_cast<float4*>(out), B, N, C / 4);
  else
    gather_rows_scalar_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, idx, out, B, N, C);
  HS_LAUNCH_CHECK();
  return HS_OK;
}
