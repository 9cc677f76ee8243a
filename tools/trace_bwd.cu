// Diagnostics: phase timeline of the tcgen05 attention backward (CTA (0,0), first units) at the BASELINE stage-0
// shape.  Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DHS_BWD_TRACE tools/trace_bwd.cu \
//        heal_swin_b200/csrc/hs_error.cpp -o tools/trace_bwd      (run on the GPU box)
#include "../heal_swin_b200/csrc/hs_attn_bwd_tc.cu"

#include <cstdio>
#include <cstdlib>
#include <vector>

int main(int argc, char** argv) {
  const int B = 8, H = 3, C = 96;
  const long long N = 196608;
  const int cos = argc > 1 ? atoi(argv[1]) : 1;
  const size_t nq = (size_t)B * N * 3 * C, no = (size_t)B * N * C;
  float *qkv, *dout, *dqkv, *bias, *ls, *dbias, *dls, *outp, *lse;
  cudaMalloc(&qkv, nq * 4); cudaMalloc(&dout, no * 4); cudaMalloc(&dqkv, nq * 4);
  cudaMalloc(&outp, no * 4); cudaMalloc(&lse, (size_t)3 * H * B * N * 4); cudaMemset(lse, 0, (size_t)3 * H * B * N * 4);
  cudaMalloc(&bias, H * 64 * 64 * 4); cudaMalloc(&ls, H * 4); cudaMalloc(&dbias, H * 64 * 64 * 4); cudaMalloc(&dls, H * 4);
  std::vector<float> h(1 << 22);
  for (auto& v : h) v = (float)rand() / RAND_MAX - 0.5f;
  for (size_t off = 0; off < nq; off += h.size()) cudaMemcpy(qkv + off, h.data(), std::min(h.size(), nq - off) * 4, cudaMemcpyHostToDevice);
  for (size_t off = 0; off < no; off += h.size()) cudaMemcpy(dout + off, h.data(), std::min(h.size(), no - off) * 4, cudaMemcpyHostToDevice);
  for (size_t off = 0; off < no; off += h.size()) cudaMemcpy(outp + off, h.data(), std::min(h.size(), no - off) * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(bias, h.data(), H * 64 * 64 * 4, cudaMemcpyHostToDevice);
  float hls[3] = {2.3f, 2.1f, 2.5f};
  cudaMemcpy(ls, hls, 12, cudaMemcpyHostToDevice);
  cudaMemset(dbias, 0, H * 64 * 64 * 4); cudaMemset(dls, 0, H * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int it = 0; it < 3; ++it) {
    cudaEventRecord(e0);
    int rc = hs::window_attn_bwd_tc(qkv, outp, lse, dout, nullptr, nullptr, bias, cos ? ls : nullptr, 0.1767f, hs::DropCfg{0.f, 0}, dqkv, dbias, dls, B, N, C, H,
                                    cos ? HS_ATTN_COS : 0, 0);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("run %d rc=%d err=%s %.3f ms\n", it, rc, cudaGetErrorString(err), ms);
  }
#ifdef HS_BWD_TRACE
  static long long tr[kTraceRoles * kTraceUnits * kTracePoints];
  cudaMemcpyFromSymbol(tr, g_trace, sizeof(tr));
  auto at = [&](int role, int n, int k) { return tr[(role * kTraceUnits + n) * kTracePoints + k]; };
  const long long t0 = at(5, 0, 0);
  const char* names[8] = {"wg0.nat", "wg0.tr ", "wg1.nat", "wg1.tr ", "mma    ", "prod   ", "epi.nat", "epi.tr "};
  printf("cycles relative to the producer's first stamp; elementwise points: full, s_ready, ds_arrive\n");
  printf("mma points: [0] waits done, [1] scores issued, [2] ds_ready(n), [3] x_ready(n), [4] dK issued; prod: [0] slot free, [1] issued\n");
  printf("epilogue points: loop top, o_ready, tmem_ld done (stage_free), store issued, slot released\n");
  for (int n = 16; n < 28; ++n) {
    for (int role = 0; role < 8; ++role) {
      if (!kCoop && role < 4 && (n & 1) != (role >> 1)) continue;
      printf("unit %2d %s:", n, names[role]);
      const int np = role < 4 ? 3 : (role == 4 ? 5 : (role == 5 ? 2 : 5));
      for (int k = 0; k < np; ++k) printf(" %8lld", at(role, n, k) - t0);
      printf("\n");
    }
  }
  // average per-phase durations over units 8..47
  {
    double d[4][4] = {};
    int cnt[4] = {};
    const int step = kCoop ? 1 : 2;
    for (int n = 8; n + step < kTraceUnits; ++n)
      for (int role = 0; role < 4; ++role) {
        if (!kCoop && (n & 1) != (role >> 1)) continue;
        d[role][0] += (double)(at(role, n, 7) - at(role, n, 0));      // statistics wait
        d[role][1] += (double)(at(role, n, 1) - at(role, n, 7));      // s_ready wait
        d[role][2] += (double)(at(role, n, 2) - at(role, n, 1));      // sweep
        d[role][3] += (double)(at(role, n + step, 0) - at(role, n, 2));  // to the next unit's full
        cnt[role]++;
      }
    for (int role = 0; role < 4; ++role)
      printf("%s avg cycles: stats wait %.0f | s_ready wait %.0f | sweep %.0f | to next full %.0f\n", names[role],
             d[role][0] / cnt[role], d[role][1] / cnt[role], d[role][2] / cnt[role], d[role][3] / cnt[role]);
    double e[2][5] = {};
    int c = 0;
    for (int n = 8; n + 1 < kTraceUnits; ++n, ++c)
      for (int role = 6; role < 8; ++role) {
        for (int k = 0; k < 4; ++k) e[role - 6][k] += (double)(at(role, n, k + 1) - at(role, n, k));
        e[role - 6][4] += (double)(at(role, n + 1, 0) - at(role, n, 4));
      }
    for (int role = 0; role < 2; ++role)
      printf("%s avg cycles: wait o_ready %.0f | tmem_ld %.0f | math+stage+store %.0f | store read %.0f | loop %.0f\n", names[6 + role],
             e[role][0] / c, e[role][1] / c, e[role][2] / c, e[role][3] / c, e[role][4] / c);
  }
  printf("unit period (cycles): %.0f\n", (double)(at(6, 46, 0) - at(6, 8, 0)) / 38.0);
#endif
  return 0;
}
