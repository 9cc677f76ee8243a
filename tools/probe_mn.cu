// Probe: tcgen05.mma kind::tf32 with BOTH operands MN-major (SWIZZLE_128B_BASE32B) spanning several 32-element slabs
// along M / N -- the operand form a token-contraction (weight-gradient) GEMM needs:  D[m][n] = sum_t A[t][m] * B[t][n].
// A: T tokens x 128 (4 slabs of 32), B: T tokens x 64 (2 slabs); a slab is T rows of 128 B in the layout TMA writes with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  Tries (LBO, SBO) per variant; run each variant in its own process.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/probe_mn.cu -o tools/probe_mn
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>

#include "../heal_swin_b200/csrc/hs_sm100.cuh"

using namespace hs::sm100;

constexpr int T = 64, MA = 128, NB = 64;
constexpr int SLAB = T * 128;  // bytes

__device__ __forceinline__ uint32_t sw128b32_off(int r, int c16) {
  return (uint32_t)(r * 128 + ((((c16 >> 1) ^ (r & 3)) << 5) | ((c16 & 1) << 4)));
}

__global__ void probe(const float* A, const float* B, float* D, uint32_t lbo_a, uint32_t sbo_a, uint32_t lbo_b, uint32_t sbo_b,
                      uint32_t step) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = sm;                 // 4 slabs
  uint8_t* sb = sm + 4 * SLAB;      // 2 slabs
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  // fill: A[t][m] -> slab m/32, row t, 16B chunk (m%32)/4
  for (int i = threadIdx.x; i < T * MA; i += blockDim.x) {
    const int t = i / MA, m = i % MA;
    *reinterpret_cast<float*>(sa + (m / 32) * SLAB + sw128b32_off(t, (m % 32) / 4) + (m % 4) * 4) = A[i];
  }
  for (int i = threadIdx.x; i < T * NB; i += blockDim.x) {
    const int t = i / NB, n = i % NB;
    *reinterpret_cast<float*>(sb + (n / 32) * SLAB + sw128b32_off(t, (n % 32) / 4) + (n % 4) * 4) = B[i];
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tmem_base, 64);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_tf32(128, NB, 1, 1);
    for (int s = 0; s < T / 8; ++s) {
      const uint64_t da = umma_desc_at(umma_smem_desc(lbo_a, sbo_a, kLayoutSw128B32), smem_u32(sa) + s * step);
      const uint64_t db = umma_desc_at(umma_smem_desc(lbo_b, sbo_b, kLayoutSw128B32), smem_u32(sb) + s * step);
      umma_tf32_ss(tm, da, db, idesc, s > 0);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  if (threadIdx.x < 128) {
    const int warp = threadIdx.x >> 5;
    uint32_t r[32];
    for (int half = 0; half < 2; ++half) {
      tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + half * 32, r);
      tmem_wait_ld();
      for (int c = 0; c < 32; ++c) D[threadIdx.x * NB + half * 32 + c] = __uint_as_float(r[c]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 64);
}

int main(int argc, char** argv) {
  const int v = argc > 1 ? atoi(argv[1]) : 0;
  struct Var { uint32_t la, sa, lb, sb, step; const char* name; };
  const Var vars[] = {
      {SLAB, 512, SLAB, 512, 1024, "lbo=slab sbo=512 step=1024"},
      {512, SLAB, 512, SLAB, 1024, "lbo=512 sbo=slab step=1024"},
      {SLAB, 1024, SLAB, 1024, 1024, "lbo=slab sbo=1024 step=1024"},
      {1024, SLAB, 1024, SLAB, 1024, "lbo=1024 sbo=slab step=1024"},
      {SLAB, 256, SLAB, 256, 1024, "lbo=slab sbo=256 step=1024"},
      {1024, 512, 1024, 512, 1024, "lbo=1024 sbo=512 step=1024 (single-slab setting)"},
  };
  const Var& V = vars[v];
  std::vector<float> A(T * MA), B(T * NB), D(MA * NB), R(MA * NB, 0.f);
  srand(1);
  auto rnd = [] { return (float)((rand() % 17) - 8) / 8.0f; };  // exactly representable in tf32
  for (auto& x : A) x = rnd();
  for (auto& x : B) x = rnd();
  for (int t = 0; t < T; ++t)
    for (int m = 0; m < MA; ++m)
      for (int n = 0; n < NB; ++n) R[m * NB + n] += A[t * MA + m] * B[t * NB + n];
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, D.size() * 4);
  const size_t smem = 6 * SLAB + 2048;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<<<1, 128, smem>>>(dA, dB, dD, V.la, V.sa, V.lb, V.sb, V.step);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("variant %d (%s): kernel error: %s\n", v, V.name, cudaGetErrorString(e)); return 2; }
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0; int bad = 0;
  int bad_by_block[4][2] = {};
  for (int m = 0; m < MA; ++m)
    for (int n = 0; n < NB; ++n) {
      const double d = fabs(D[m * NB + n] - R[m * NB + n]);
      if (d > maxerr) maxerr = d;
      if (d > 1e-3) { ++bad; ++bad_by_block[m / 32][n / 32]; }
    }
  printf("variant %d (%s): max_err %.3g bad %d -> %s   bad per (m-slab, n-slab):", v, V.name, maxerr, bad, bad ? "FAIL" : "PASS");
  for (int i = 0; i < 4; ++i) printf(" [%d %d]", bad_by_block[i][0], bad_by_block[i][1]);
  printf("\n");
  return 0;
}
