// Pieces shared by the tcgen05 attention kernels (forward: hs_attn_tc.cu, backward: hs_attn_bwd_tc.cu):
// tile constants, the shared-memory swizzle address maps and the host-side TMA descriptor builder.
#pragma once
#include "hs_common.h"
#include "hs_sm100.cuh"

namespace hs {
namespace tc {

constexpr int kWS = 64;              // tokens per window
constexpr int kD = 32;               // head dim
constexpr int kTile = kWS * kD * 4;  // 8192 B: one 64 x 32 fp32 tile
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLogitScaleMax = 4.605170185988092f;  // log(1/0.01), swin_hp_transformer.py:144-146
constexpr float kNormEps = 1e-12f;                    // F.normalize eps
constexpr float kMaskFill = -100.0f;                  // hp_shifting.py:25

// tcgen05 kind::tf32 reads the top 19 bits of each fp32 operand word, i.e. it TRUNCATES the mantissa to 10 bits.  A
// truncated operand is on average (1 - 2^-11 * E[1/m]) of its value, E[1/m] = 1/(2 ln 2) = 0.7213 for a log-uniform
// mantissa m in [1, 2): a systematic shrink of 3.52e-4 per truncated operand, on top of a zero-mean error with the same
// variance as round-to-nearest.  The kernels multiply each product by the inverse of that mean shrink (once per
// truncated operand) so that the TF32 products are unbiased; operands the kernels write themselves (P, dS) are
// rounded to nearest and need no correction.
constexpr float kTruncFix1 = 1.0003522f;  // one truncated operand
constexpr float kTruncFix2 = 1.0007045f;  // two truncated operands

constexpr int kFlagContig = 1, kFlagUniform = 2, kFlagValid = 4;

// byte offset of 16-byte chunk c16 of row r inside a tile stored with the TMA 128B swizzle (K-major operand)
__device__ __forceinline__ uint32_t sw128_off(int r, int c16) { return (uint32_t)(r * 128 + ((c16 ^ (r & 7)) << 4)); }
// same for CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B (MN-major 32-bit operand)
__device__ __forceinline__ uint32_t sw128b32_off(int r, int c16) {
  return (uint32_t)(r * 128 + ((((c16 >> 1) ^ (r & 3)) << 5) | ((c16 & 1) << 4)));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// 2-D fp32 tensor (rows x cols, dense) with a (box_cols x box_rows) box
// (pitch_cols: row pitch in elements when the tensor is a column block of a wider one; 0 = dense)
inline int make_map(CUtensorMap* m, const float* base, long long rows, int cols, CUtensorMapSwizzle sw,
                    int box_cols = kD, int box_rows = kWS, long long pitch_cols = 0) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return hs::fail(HS_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)(pitch_cols ? pitch_cols : cols) * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return hs::fail(HS_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return HS_OK;
}

inline int sm_count() {
  int dev = 0, n = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n > 0 ? n : 148;
}

}  // namespace tc
}  // namespace hs
