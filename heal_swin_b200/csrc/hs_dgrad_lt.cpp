// Input gradient of nn.Linear with the residual-shortcut gradient folded in:   dx = dy @ W + c     (library GEMM)
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

// one plan (descriptors + the heuristic's first algorithm) per (device, T, N, K, beta != 0, workspace size)
struct Plan {
  cublasLtMatmulDesc_t op = nullptr;
  cublasLtMatrixLayout_t a = nullptr, b = nullptr, c = nullptr;
  cublasLtMatmulAlgo_t algo;
  size_t workspace = 0;
};

std::mutex g_mu;
std::map<int, cublasLtHandle_t> g_handles;
std::map<std::tuple<int, long long, int, int, int, size_t>, Plan> g_plans;

}  // namespace

extern "C" int hs_linear_dgrad_acc(const float* dy, const float* w, const float* c, float* dx, int64_t T, int N, int K,
                                   void* workspace, uint64_t workspace_bytes, void* stream) {
  HS_REQUIRE(dy && w && dx && T > 0 && N > 0 && K > 0, "hs_linear_dgrad_acc: bad arguments");
  const LtApi& lt = lt_api();
  if (!lt.ok) return hs::fail(HS_ERR_CUDA, "hs_linear_dgrad_acc: cuBLASLt (libcublasLt.so.12) could not be loaded");
  int dev = 0;
  HS_CUDA(cudaGetDevice(&dev));
  cublasLtHandle_t handle;
  Plan plan;
  {
    std::lock_guard<std::mutex> lock(g_mu);
    auto h = g_handles.find(dev);
    if (h == g_handles.end()) {
      cublasLtHandle_t nh;
      if (lt.create(&nh) != CUBLAS_STATUS_SUCCESS) return hs::fail(HS_ERR_CUDA, "cublasLtCreate failed");
      h = g_handles.emplace(dev, nh).first;
    }
    handle = h->second;
    const auto key = std::make_tuple(dev, (long long)T, N, K, c ? 1 : 0, (size_t)workspace_bytes);
    auto p = g_plans.find(key);
    if (p == g_plans.end()) {
      // row-major dx (T, K) = dy (T, N) @ w (N, K)   <=>   column-major dx^T (K, T) = w^T (K, N) @ dy^T (N, T)
      Plan np;
      cublasStatus_t st = lt.desc_create(&np.op, CUBLAS_COMPUTE_32F_FAST_TF32, CUDA_R_32F);
      if (st == CUBLAS_STATUS_SUCCESS) st = lt.layout_create(&np.a, CUDA_R_32F, (uint64_t)K, (uint64_t)N, (int64_t)K);
      if (st == CUBLAS_STATUS_SUCCESS) st = lt.layout_create(&np.b, CUDA_R_32F, (uint64_t)N, (uint64_t)T, (int64_t)N);
      if (st == CUBLAS_STATUS_SUCCESS) st = lt.layout_create(&np.c, CUDA_R_32F, (uint64_t)K, (uint64_t)T, (int64_t)K);
      cublasLtMatmulPreference_t pref = nullptr;
      if (st == CUBLAS_STATUS_SUCCESS) st = lt.pref_create(&pref);
      size_t ws = workspace ? (size_t)workspace_bytes : 0;
      if (st == CUBLAS_STATUS_SUCCESS)
        st = lt.pref_set(pref, CUBLASLT_MATMUL_PREF_MAX_WORKSPACE_BYTES, &ws, sizeof(ws));
      cublasLtMatmulHeuristicResult_t res;
      int found = 0;
      if (st == CUBLAS_STATUS_SUCCESS) st = lt.heuristic(handle, np.op, np.a, np.b, np.c, np.c, pref, 1, &res, &found);
      if (pref) lt.pref_destroy(pref);
      if (st != CUBLAS_STATUS_SUCCESS || found == 0)
        return hs::fail(HS_ERR_CUDA, "hs_linear_dgrad_acc: cuBLASLt has no algorithm for T=%lld N=%d K=%d (status %d)",
                        (long long)T, N, K, (int)st);
      np.algo = res.algo;
      np.workspace = res.workspaceSize;
      p = g_plans.emplace(key, np).first;
    }
    plan = p->second;
  }
  const float alpha = 1.0f, beta = c ? 1.0f : 0.0f;
  cublasStatus_t st = lt.matmul(handle, plan.op, &alpha, w, plan.a, dy, plan.b, &beta, c ? c : dx, plan.c, dx, plan.c,
                                &plan.algo, workspace, plan.workspace, (cudaStream_t)stream);
  if (st != CUBLAS_STATUS_SUCCESS)
    return hs::fail(HS_ERR_CUDA, "hs_linear_dgrad_acc: cublasLtMatmul failed with status %d", (int)st);
  return HS_OK;
}
