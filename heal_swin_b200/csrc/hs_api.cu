// C-ABI entry points of the attention kernels: argument validation + kernel-family dispatch.
#include "hs_common.h"
#include "hs_kernels.h"

extern "C" {

int hs_window_attn_fwd(const float* qkv, const int32_t* src, const uint8_t* groups, const float* mask,
                       const float* bias, const float* logit_scale, float scale, float attn_drop, uint64_t seed,
                       float* out, float* lse, int B, int64_t N, int C, int H, int ws, uint32_t flags, void* stream) {
  HS_REQUIRE(attn_drop >= 0.f && attn_drop < 1.f, "hs_window_attn_fwd: attn_drop must be in [0, 1), got %f", attn_drop);
  const hs::DropCfg drop{attn_drop, seed};
  if (!(flags & HS_ATTN_NO_TC) && hs::window_attn_tc_supported(qkv, out, mask, B, N, C, H, ws))
    return hs::window_attn_fwd_tc(qkv, src, groups, bias, logit_scale, scale, drop, out, lse, B, N, C, H, flags,
                                  (cudaStream_t)stream);
  return hs::window_attn_fwd_simt(qkv, src, groups, mask, bias, logit_scale, scale, drop, out, lse, B, N, C, H, ws,
                                  flags, (cudaStream_t)stream);
}

int hs_window_attn_bwd(const float* qkv, const float* out, const float* lse, const float* dout, const int32_t* src,
                       const uint8_t* groups, const float* mask, const float* bias, const float* logit_scale,
                       float scale, float attn_drop, uint64_t seed, float* dqkv, float* dbias, float* dlogit_scale,
                       int B, int64_t N, int C, int H, int ws, uint32_t flags, void* stream) {
  HS_REQUIRE(attn_drop >= 0.f && attn_drop < 1.f, "hs_window_attn_bwd: attn_drop must be in [0, 1), got %f", attn_drop);
  const hs::DropCfg drop{attn_drop, seed};
  if (!(flags & HS_ATTN_NO_TC) && out && lse && hs::window_attn_tc_supported(qkv, dqkv, mask, B, N, C, H, ws) &&
      !((reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(out)) & 15))
    return hs::window_attn_bwd_tc(qkv, out, lse, dout, src, groups, bias, logit_scale, scale, drop, dqkv, dbias,
                                  dlogit_scale, B, N, C, H, flags, (cudaStream_t)stream);
  return hs::window_attn_bwd_simt(qkv, dout, src, groups, mask, bias, logit_scale, scale, drop, dqkv, dbias,
                                  dlogit_scale, B, N, C, H, ws, flags, (cudaStream_t)stream);
}

}  // extern "C"
