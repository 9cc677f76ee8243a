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
