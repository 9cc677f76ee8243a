// Cross-entropy of the segmentation head, forward and gradient in one pass over the logits
// (nn.CrossEntropyLoss at heal_swin/models_lightning/segmentation/model_lightning_swin_hp.py:45, applied to the network
// output (B, K, P) at :109 -- class planes of P pixels, K = f_out classes).  torch runs log_softmax forward, nll forward,
// nll backward and log_softmax backward as four kernels with two (B, K, P) temporaries; here one thread owns one pixel,
// reads its K logits once (coalesced across pixels within each class plane), and writes the UNSCALED gradient
// softmax(logits) - onehot(target); the loss sum and the number of counted pixels are block-reduced and added atomically.
// The mean (1 / count) and the upstream gradient are applied by the caller.
#include "hs_common.h"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxK = 32;

template <typename TT>
__global__ void __launch_bounds__(kThreads)
ce_fwd_bwd_kernel(const float* __restrict__ logits, const TT* __restrict__ target, float* __restrict__ dlogits,
                  float* __restrict__ acc /* [0] loss sum, [1] counted pixels */, int K, long long P, long long total,
                  long long ignore_index) {
  float loss = 0.f, cnt = 0.f;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
    const long long b = i / P, p = i - b * P;
    const float* lp = logits + b * K * P + p;
    float v[kMaxK];
    float mx = -3.4e38f;
    // fully unrolled over the maximum class count with guards: v[] stays in registers
#pragma unroll
    for (int k = 0; k < kMaxK; ++k)
      if (k < K) {
        v[k] = __ldcs(lp + (long long)k * P);
        mx = fmaxf(mx, v[k]);
      }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxK; ++k)
      if (k < K) {
        v[k] = __expf(v[k] - mx);
        s += v[k];
      }
    const long long t = (long long)target[i];
    const bool counted = t != ignore_index && t >= 0 && t < K;
    const float inv = 1.0f / s;
    float* dp = dlogits + b * K * P + p;
#pragma unroll
    for (int k = 0; k < kMaxK; ++k)
      if (k < K) {
        const float sm = v[k] * inv;
        __stcs(dp + (long long)k * P, counted ? (sm - (k == t ? 1.0f : 0.0f)) : 0.0f);
        if (counted && k == t) loss -= __logf(fmaxf(sm, 1e-38f));
      }
    cnt += counted ? 1.0f : 0.0f;
  }
  // block reduction, one atomic pair per block
  __shared__ float red[2][kThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    loss += __shfl_xor_sync(0xffffffffu, loss, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = loss;
    red[1][threadIdx.x >> 5] = cnt;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    loss = threadIdx.x < kThreads / 32 ? red[0][threadIdx.x] : 0.f;
    cnt = threadIdx.x < kThreads / 32 ? red[1][threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      loss += __shfl_xor_sync(0xffffffffu, loss, o);
      cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if (threadIdx.x == 0) {
      atomicAdd(acc, loss);
      atomicAdd(acc + 1, cnt);
    }
  }
}

}  // namespace

extern "C" {

int hs_cross_entropy_supported(int K) { return (K >= 1 && K <= kMaxK) ? 1 : 0; }

int hs_cross_entropy(const float* logits, const void* target, int target_bytes, float* dlogits, float* acc, int B, int K,
                     int64_t P, int64_t ignore_index, void* stream) {
  HS_REQUIRE(logits && target && dlogits && acc && B > 0 && P > 0, "hs_cross_entropy: bad arguments");
  HS_REQUIRE(target_bytes == 1 || target_bytes == 8, "hs_cross_entropy: targets must be uint8 or int64 (got %d bytes)",
             target_bytes);
  if (!hs_cross_entropy_supported(K))
    return hs::fail(HS_ERR_UNSUPPORTED, "hs_cross_entropy: K=%d classes not covered (1..%d)", K, kMaxK);
  const long long total = (long long)B * P;
  long long blocks = (total + kThreads - 1) / kThreads;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
  if (target_bytes == 1)
    ce_fwd_bwd_kernel<uint8_t><<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(
        logits, static_cast<const uint8_t*>(target), dlogits, acc, K, P, total, ignore_index);
  else
    ce_fwd_bwd_kernel<long long><<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(
        logits, static_cast<const long long*>(target), dlogits, acc, K, P, total, ignore_index);
  HS_LAUNCH_CHECK();
  return HS_OK;
}

}  // extern "C"
