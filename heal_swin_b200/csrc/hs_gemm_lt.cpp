// Library GEMMs through cuBLASLt: the forward of nn.Linear (y = x W^T + b) and its input gradient with the residual-
// shortcut gradient folded in (dx = dy @ W + c).
// cuBLASLt is the one library that takes C and D as different buffers, which is what removes the separate gradient
// accumulation pass of every residual connection (autograd of  x + branch(x)  at swin_hp_transformer.py:333-338): the
// shortcut's gradient c is read once by the GEMM epilogue instead of being copied or added in a pass of its own.
// cuBLASLt is bound at run time (dlopen of the soname the process already has through PyTorch, or the CUDA toolkit's):
// the library keeps no link-time dependency on it.
#include <cublasLt.h>
#include <dlfcn.h>

#include <map>
#include <mutex>
#include <tuple>

#include "hs_common.h"

namespace {

struct LtApi {
  void* so = nullptr;
  decltype(&cublasLtCreate) create = nullptr;
  decltype(&cublasLtMatmulDescCreate) desc_create = nullptr;
  decltype(&cublasLtMatmulDescDestroy) desc_destroy = nullptr;
  decltype(&cublasLtMatmulDescSetAttribute) desc_set = nullptr;
  decltype(&cublasLtMatrixLayoutCreate) layout_create = nullptr;
  decltype(&cublasLtMatrixLayoutDestroy) layout_destroy = nullptr;
  decltype(&cublasLtMatmulPreferenceCreate) pref_create = nullptr;
  decltype(&cublasLtMatmulPreferenceSetAttribute) pref_set = nullptr;
  decltype(&cublasLtMatmulPreferenceDestroy) pref_destroy = nullptr;
  decltype(&cublasLtMatmulAlgoGetHeuristic) heuristic = nullptr;
  decltype(&cublasLtMatmul) matmul = nullptr;
  bool ok = false;
};

template <typename F>
bool bind(void* so, const char* name, F* fn) {
  *fn = reinterpret_cast<F>(dlsym(so, name));
  return *fn != nullptr;
}

const LtApi& lt_api() {
  static LtApi api = [] {
    LtApi a;
    for (const char* name : {"libcublasLt.so.12", "libcublasLt.so.13", "libcublasLt.so"}) {
      a.so = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (a.so) break;
    }
    if (!a.so) return a;
    a.ok = bind(a.so, "cublasLtCreate", &a.create) && bind(a.so, "cublasLtMatmulDescCreate", &a.desc_create) &&
           bind(a.so, "cublasLtMatmulDescDestroy", &a.desc_destroy) &&
           bind(a.so, "cublasLtMatmulDescSetAttribute", &a.desc_set) &&
           bind(a.so, "cublasLtMatrixLayoutCreate", &a.layout_create) &&
           bind(a.so, "cublasLtMatrixLayoutDestroy", &a.layout_destroy) &&
           bind(a.so, "cublasLtMatmulPreferenceCreate", &a.pref_create) &&
           bind(a.so, "cublasLtMatmulPreferenceSetAttribute", &a.pref_set) &&
           bind(a.so, "cublasLtMatmulPreferenceDestroy", &a.pref_destroy) &&
           bind(a.so, "cublasLtMatmulAlgoGetHeuristic", &a.heuristic) && bind(a.so, "cublasLtMatmul", &a.matmul);
    return a;
  }();
  return api;
}

// one plan (descriptors + the heuristic's first algorithm) per (device, kind, T, N, K, workspace size)
struct Plan {
  cublasLtMatmulDesc_t op = nullptr;
  cublasLtMatrixLayout_t a = nullptr, b = nullptr, c = nullptr;
  cublasLtMatmulAlgo_t algo;
  size_t workspace = 0;
};

std::mutex g_mu;
std::map<int, cublasLtHandle_t> g_handles;
std::map<std::tuple<int, int, long long, int, int, size_t>, Plan> g_plans;

enum Kind { kDgrad = 0, kForward = 1, kForwardBias = 2 };

// kDgrad:    out (T, K) = in (T, N) @ w (N, K) [+ c]     column-major: out^T (K, T) = w^T (K, N) . in^T (N, T)
// kForward*: out (T, N) = in (T, K) @ w (N, K)^T [+ bias] column-major: out^T (N, T) = (w^T (K, N))^T . in^T (K, T)
int lt_gemm(Kind kind, const float* in, const float* w, const float* c, const float* bias, float* out, int64_t T, int N,
            int K, void* workspace, uint64_t workspace_bytes, void* stream, const char* who) {
  const LtApi& lt = lt_api();
  if (!lt.ok) return hs::fail(HS_ERR_CUDA, "%s: cuBLASLt (libcublasLt.so.12) could not be loaded", who);
  int dev = 0;
  HS_CUDA(cudaGetDevice(&dev));
  cublasLtHandle_t handle;
  Plan plan;
  {
    std::lock_guard<std::mutex> lock(g_mu);
    auto h = g_handles.find(dev);
    if (h == g_handles.end()) {
      cublasLtHandle_t nh;
      if (lt.create(&nh) != CUBLAS_STATUS_SUCCESS) return hs::fail(HS_ERR_CUDA, "%s: cublasLtCreate failed", who);
      h = g_handles.emplace(dev, nh).first;
    }
    handle = h->second;
    const size_t ws = workspace ? (size_t)workspace_bytes : 0;
    const auto key = std::make_tuple(dev, (int)kind, (long long)T, N, K, ws);
    auto p = g_plans.find(key);
    if (p == g_plans.end()) {
      Plan np;
      const uint64_t m = kind == kDgrad ? K : N, kk = kind == kDgrad ? N : K;  // column-major D is (m x T), contraction kk
      cublasStatus_t st = lt.desc_create(&np.op, CUBLAS_COMPUTE_32F_FAST_TF32, CUDA_R_32F);
      if (kind != kDgrad && st == CUBLAS_STATUS_SUCCESS) {
        const cublasOperation_t tr = CUBLAS_OP_T;
        st = lt.desc_set(np.op, CUBLASLT_MATMUL_DESC_TRANSA, &tr, sizeof(tr));
      }
      if (kind == kForwardBias && st == CUBLAS_STATUS_SUCCESS) {
        const cublasLtEpilogue_t epi = CUBLASLT_EPILOGUE_BIAS;
        st = lt.desc_set(np.op, CUBLASLT_MATMUL_DESC_EPILOGUE, &epi, sizeof(epi));
      }
      // A = w as stored: column-major (K x N), leading dimension K
      if (st == CUBLAS_STATUS_SUCCESS) st = lt.layout_create(&np.a, CUDA_R_32F, (uint64_t)K, (uint64_t)N, (int64_t)K);
      if (st == CUBLAS_STATUS_SUCCESS) st = lt.layout_create(&np.b, CUDA_R_32F, kk, (uint64_t)T, (int64_t)kk);
      if (st == CUBLAS_STATUS_SUCCESS) st = lt.layout_create(&np.c, CUDA_R_32F, m, (uint64_t)T, (int64_t)m);
      cublasLtMatmulPreference_t pref = nullptr;
      if (st == CUBLAS_STATUS_SUCCESS) st = lt.pref_create(&pref);
      if (st == CUBLAS_STATUS_SUCCESS) st = lt.pref_set(pref, CUBLASLT_MATMUL_PREF_MAX_WORKSPACE_BYTES, &ws, sizeof(ws));
      cublasLtMatmulHeuristicResult_t res;
      int found = 0;
      if (kind == kForwardBias && st == CUBLAS_STATUS_SUCCESS)  // the heuristic wants to see a bias pointer
        st = lt.desc_set(np.op, CUBLASLT_MATMUL_DESC_BIAS_POINTER, &bias, sizeof(bias));
      if (st == CUBLAS_STATUS_SUCCESS) st = lt.heuristic(handle, np.op, np.a, np.b, np.c, np.c, pref, 1, &res, &found);
      if (pref) lt.pref_destroy(pref);
      if (st != CUBLAS_STATUS_SUCCESS || found == 0)
        return hs::fail(HS_ERR_CUDA, "%s: cuBLASLt has no algorithm for T=%lld N=%d K=%d (status %d)", who, (long long)T, N,
                        K, (int)st);
      np.algo = res.algo;
      np.workspace = res.workspaceSize;
      p = g_plans.emplace(key, np).first;
    }
    plan = p->second;
    if (kind == kForwardBias) {
      // the bias pointer is part of the (shared) descriptor: set and launch under the lock
      if (lt.desc_set(plan.op, CUBLASLT_MATMUL_DESC_BIAS_POINTER, &bias, sizeof(bias)) != CUBLAS_STATUS_SUCCESS)
        return hs::fail(HS_ERR_CUDA, "%s: could not set the bias pointer", who);
      const float alpha = 1.0f, beta = 0.0f;
      cublasStatus_t st = lt.matmul(handle, plan.op, &alpha, w, plan.a, in, plan.b, &beta, out, plan.c, out, plan.c,
                                    &plan.algo, workspace, plan.workspace, (cudaStream_t)stream);
      if (st != CUBLAS_STATUS_SUCCESS) return hs::fail(HS_ERR_CUDA, "%s: cublasLtMatmul failed with status %d", who, (int)st);
      return HS_OK;
    }
  }
  const float alpha = 1.0f, beta = c ? 1.0f : 0.0f;
  cublasStatus_t st = lt.matmul(handle, plan.op, &alpha, w, plan.a, in, plan.b, &beta, c ? c : out, plan.c, out, plan.c,
                                &plan.algo, workspace, plan.workspace, (cudaStream_t)stream);
  if (st != CUBLAS_STATUS_SUCCESS) return hs::fail(HS_ERR_CUDA, "%s: cublasLtMatmul failed with status %d", who, (int)st);
  return HS_OK;
}

}  // namespace

extern "C" int hs_linear_dgrad_acc(const float* dy, const float* w, const float* c, float* dx, int64_t T, int N, int K,
                                   void* workspace, uint64_t workspace_bytes, void* stream) {
  HS_REQUIRE(dy && w && dx && T > 0 && N > 0 && K > 0, "hs_linear_dgrad_acc: bad arguments");
  return lt_gemm(kDgrad, dy, w, c, nullptr, dx, T, N, K, workspace, workspace_bytes, stream, "hs_linear_dgrad_acc");
}

extern "C" int hs_linear_fwd(const float* x, const float* w, const float* bias, float* y, int64_t T, int N, int K,
                             void* workspace, uint64_t workspace_bytes, void* stream) {
  HS_REQUIRE(x && w && y && T > 0 && N > 0 && K > 0, "hs_linear_fwd: bad arguments");
  return lt_gemm(bias ? kForwardBias : kForward, x, w, nullptr, bias, y, T, N, K, workspace, workspace_bytes, stream,
                 "hs_linear_fwd");
}
