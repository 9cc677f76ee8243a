// Shared helpers for the healswin_b200 C-ABI library (error reporting, launch checks).
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/healswin_b200.h"

namespace hs {

// thread-local last-error message (hs_last_error)
char* error_buffer();
int fail(int code, const char* fmt, ...);

}  // namespace hs

// ---- attention-probability dropout (swin_hp_transformer.py:167, nn.Dropout(attn_drop) on the softmax output) ----
// Counter-based: whether P[i][j] of (window-in-batch wb, head h) is kept is a pure function of (seed, wb, h, i, j), so
// the forward, the backward and both orientations of the tensor-core backward regenerate identical masks without
// storing them.  lowbias32 integer mix; an entry is dropped when its hash is below p * 2^32.
#if defined(__CUDACC__)
#define HS_HD __host__ __device__ __forceinline__
#else
#define HS_HD inline
#endif
namespace hs {
HS_HD uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
// INDIRECT seeds.  A seed with bit 63 set is not a value but a reference: bits 0-47 hold the address of a uint64 counter in
// device memory, bits 48-62 a call id; the effective seed is a mix of the counter's CURRENT value and the id.  This is how
// a captured CUDA graph gets fresh masks on every replay (the graph bumps the counter once per replay; the host-side
// integers baked into the kernel nodes stay the same): heal_swin_b200/graph.py.  Forward and backward of an op pass the
// same reference and read the same counter value within one replay.  Device code only.
HS_HD uint64_t resolve_seed(uint64_t seed) {
#if defined(__CUDA_ARCH__)
  if (seed >> 63) {
    const unsigned long long base = *reinterpret_cast<const unsigned long long*>(seed & 0xFFFFFFFFFFFFull);
    const unsigned long long id = (seed >> 48) & 0x7FFFull;
    unsigned long long x = (base + 0x9E3779B97F4A7C15ull) * (2 * id + 1);
    x ^= x >> 31; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 29;
    return x;
  }
#endif
  return seed;
}
HS_HD uint32_t drop_unit_key(uint64_t seed, long long wb, int h, int H) {
  seed = resolve_seed(seed);
  return mix32((uint32_t)seed ^ mix32((uint32_t)(seed >> 32) + (uint32_t)(wb * H + h)));
}
HS_HD bool drop_keep(uint32_t unit_key, int i, int j, int ws, uint32_t thresh) {
  return mix32(unit_key ^ ((uint32_t)(i * ws + j) * 0x9E3779B9u)) >= thresh;
}
// element-wise dropout of a (rows, C) activation (proj_drop / Mlp.drop, swin_hp_transformer.py:38-43, 173): the mask of
// element (row, col) is a pure function of (seed, row, col)
HS_HD uint32_t drop_row_key(uint64_t seed, long long row) {
  seed = resolve_seed(seed);
  return mix32((uint32_t)seed ^ mix32((uint32_t)(seed >> 32) + (uint32_t)row) ^ (uint32_t)((unsigned long long)row >> 32));
}
HS_HD bool drop_keep_elem(uint32_t row_key, int col, uint32_t thresh) {
  return mix32(row_key ^ ((uint32_t)col * 0x9E3779B9u)) >= thresh;
}
HS_HD uint32_t drop_thresh(float p) {
  const double t = (double)p * 4294967296.0;
  return t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
}
}  // namespace hs

#define HS_REQUIRE(cond, ...)                                  \
  do {                                                         \
    if (!(cond)) return hs::fail(HS_ERR_ARG, __VA_ARGS__);     \
  } while (0)

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define HS_CUDA(expr)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      return hs::fail(HS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                      __FILE__, __LINE__);                                                 \
  } while (0)
#define HS_LAUNCH_CHECK() HS_CUDA(cudaGetLastError())
#endif
