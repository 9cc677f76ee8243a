// Internal kernel-family entry points (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace hs {

// attention-probability dropout: p = 0 disables it (see hs_common.h for the mask definition)
struct DropCfg {
  float p;
  uint64_t seed;
};

int window_attn_fwd_simt(const float* qkv, const int32_t* src, const uint8_t* groups, const float* mask,
                         const float* bias, const float* logit_scale, float scale, DropCfg drop, float* out, float* lse,
                         int B, int64_t N, int C, int H, int ws, uint32_t flags, cudaStream_t stream);
int window_attn_bwd_simt(const float* qkv, const float* dout, const int32_t* src, const uint8_t* groups,
                         const float* mask, const float* bias, const float* logit_scale, float scale, DropCfg drop,
                         float* dqkv, float* dbias, float* dlogit, int B, int64_t N, int C, int H, int ws,
                         uint32_t flags, cudaStream_t stream);

// tcgen05 / TMA path (window 64, head_dim 32, no dense mask); see hs_attn_tc.cu / hs_attn_bwd_tc.cu
bool window_attn_tc_supported(const float* qkv, const float* out, const float* mask, int B, int64_t N, int C, int H,
                              int ws);
int window_attn_fwd_tc(const float* qkv, const int32_t* src, const uint8_t* groups, const float* bias,
                       const float* logit_scale, float scale, DropCfg drop, float* out, float* lse, int B, int64_t N,
                       int C, int H, uint32_t flags, cudaStream_t stream);
int window_attn_bwd_tc(const float* qkv, const float* out, const float* lse, const float* dout, const int32_t* src,
                       const uint8_t* groups, const float* bias, const float* logit_scale, float scale, DropCfg drop,
                       float* dqkv, float* dbias, float* dlogit, int B, int64_t N, int C, int H, uint32_t flags,
                       cudaStream_t stream);

}  // namespace hs
