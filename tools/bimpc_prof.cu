// Developer tool: cycle breakdown of one BiMPC station solve (csrc/bimpc_solve.cuh built with
// -DBIMPC_PROFILE).  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DBIMPC_PROFILE \
//   -Iincentive-design-mpc_b200/csrc -Iinclude tools/bimpc_prof.cu -o /tmp/bimpc_prof && /tmp/bimpc_prof
#include <cstdio>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "bimpc_solve.cuh"

int main() {
  const int N = 24, P = 12, S = 1;
  bimpc::BiConsts c{N, P, 1e3, 1.0, 1.0, 0.3, 0.3, 2, 10.0, 50.0, 0.25, 0.15};
  std::vector<double> om(N), Mp(P, 500.0 / 12 / 30000.0), beta(P, 0.02), gam(P), x0(1, 0.05), dem(N);
  for (int k = 0; k < N; ++k) om[k] = std::pow(5.0, k - N + 1), dem[k] = 0.6 + 0.15 * std::sin(0.26 * k);
  for (int p = 0; p < P; ++p) gam[p] = 0.9 - (0.3 + 0.05 * p) - 0.02;
  auto up = [](const std::vector<double>& v) { double* d; cudaMalloc(&d, v.size() * 8); cudaMemcpy(d, v.data(), v.size() * 8, cudaMemcpyHostToDevice); return d; };
  double *d_om = up(om), *d_mp = up(Mp), *d_b = up(beta), *d_g = up(gam), *d_x0 = up(x0), *d_dem = up(dem);
  double *ws, *wl, *ug, *obj; int32_t *st, *it; long long* prof;
  cudaMalloc(&ws, P * N * 8); cudaMalloc(&wl, P * N * 8); cudaMalloc(&ug, N * 8); cudaMalloc(&obj, 8);
  cudaMalloc(&st, 4); cudaMalloc(&it, 4); cudaMalloc(&prof, 64); cudaMemset(prof, 0, 64);
  double* li; cudaMalloc(&li, bimpc::global_scratch_doubles(N, P) * 8);
  bimpc::BiArgs a{S, d_om, d_mp, d_mp, d_b, d_b, d_g, d_g, d_x0, d_dem, ws, wl, ug, st, it, obj, 1e-9, 100, prof, li};
  const size_t smem = bimpc::scratch_doubles(N, P, bimpc::kThreads) * 8;
  cudaFuncSetAttribute(bimpc::bimpc_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemset(prof, 0, 64);
    cudaEventRecord(e0);
    bimpc::bimpc_solve_kernel<<<1, bimpc::kThreads, smem>>>(c, a);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
  }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long hp[8]; int hst, hit;
  cudaMemcpy(hp, prof, 64, cudaMemcpyDeviceToHost); cudaMemcpy(&hst, st, 4, cudaMemcpyDeviceToHost); cudaMemcpy(&hit, it, 4, cudaMemcpyDeviceToHost);
  printf("status %d iters %d  %.3f ms  smem %zu B  err %s\n", hst, hit, ms, smem, cudaGetErrorString(cudaGetLastError()));
  printf("cycles: residuals %lld  factor %lld  solves %lld  predictor+corrector total %lld\n", hp[0], hp[1], hp[2], hp[3]);
  printf("factor: la/tl %lld  schur %lld  elimination %lld  scaling %lld\n", hp[4], hp[5], hp[6], hp[7]);
  return 0;
}
