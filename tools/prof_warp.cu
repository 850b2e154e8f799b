// Per-phase cycle profile of the warp-cooperative K1 (csrc/lompc_solve_warp.cuh, compiled with
// LOMPC_WARP_PROF) on the bench's workload shape, plus the floor of the timing harness (an empty kernel
// between two events after an L2 flush).   nvcc ... -DLOMPC_WARP_PROF -I include -I .../csrc tools/prof_warp.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "lompc_solve_warp.cuh"

namespace lompc_detail {
void count_launch() {}
int cuda_fail(cudaError_t e, const char* what) { fprintf(stderr, "%s: %s\n", what, cudaGetErrorString(e)); return -3; }
}  // namespace lompc_detail

__global__ void empty_kernel() {}

static lompc::Consts make_consts(int N, bool large) {
  lompc::Consts cs;
  memset(&cs, 0, sizeof(cs));
  const double delta = large ? 0.025 : 0.05, theta = large ? 50.0 : 10.0, y_max = 0.9, w_max = large ? 0.15 : 0.25;
  cs.N = N; cs.large = large; cs.delta = delta; cs.theta = theta; cs.w_max = w_max; cs.y_max = y_max;
  cs.c = 2.0 * delta * theta * theta; cs.q_scale = 3.0 * theta / (4.0 * w_max); cs.theta2 = theta * theta;
  if (!large) {
    cs.d_base = 2.0 * theta * theta / 0.81; cs.nseg = 1; cs.brk[0] = 0.0; cs.brk[1] = w_max; cs.slope[0] = 0.0;
  } else {
    cs.nseg = 4;
    const double br[5] = {0.0, 0.125, 0.5, 0.75, 1.0}, sl[4] = {0.0, 1.0, 1.5, 2.0};
    const double per_w = (theta * w_max) * (theta * w_max) / w_max;
    for (int i = 0; i <= 4; ++i) cs.brk[i] = br[i] * w_max;
    for (int j = 0; j < 4; ++j) cs.slope[j] = per_w * sl[j];
  }
  return cs;
}

// rotate = 0: L2 flushed (256 MiB memset) before every launch -- data AND code come from DRAM;
// rotate = 1: no flush, the launches walk over R distinct input copies of > 126 MB in total (the L2 size), so
//             the data is cold but the kernel's code stays cached.
template <int N, int SPL>
void run(int B, bool large, int reps, int rotate = 0) {
  constexpr int QPW = 32 / (N / SPL);
  lompc::Consts cs = make_consts(N, large);
  std::vector<double> lm((size_t)B * 3 * N), lr(B), gam(B);
  srand(2 + large);
  auto u = []() { return rand() / (RAND_MAX + 1.0); };
  for (auto& x : lm) x = cs.theta * u();
  for (auto& x : lr) x = 3 * N * cs.delta * u();
  for (auto& x : gam) x = cs.y_max * u();
  double *d_lm, *d_lr, *d_g, *d_w, *d_c;
  int32_t *d_st, *d_it;
  const size_t per_copy = (size_t)B * (4 * N + 3) * 8;
  const int R = rotate ? (int)((size_t)140e6 / per_copy + 1) : 1;
  cudaMalloc(&d_lm, lm.size() * 8 * R); cudaMalloc(&d_lr, (size_t)B * 8 * R); cudaMalloc(&d_g, (size_t)B * 8 * R);
  cudaMalloc(&d_w, (size_t)B * N * 8 * R); cudaMalloc(&d_c, (size_t)B * 8 * R); cudaMalloc(&d_st, B * 4); cudaMalloc(&d_it, B * 4);
  for (int r = 0; r < R; ++r) {
    cudaMemcpy(d_lm + lm.size() * r, lm.data(), lm.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(d_lr + (size_t)B * r, lr.data(), B * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(d_g + (size_t)B * r, gam.data(), B * 8, cudaMemcpyHostToDevice);
  }
  lompc::WarpArgs wa;
  memset(&wa, 0, sizeof(wa));
  wa.nsegs = 1;
  wa.total_warps = (B + QPW - 1) / QPW;
  wa.seg[0].cs = cs;
  lompc::SolveArgs& a = wa.seg[0].a;
  a.B = B; a.lmbd = d_lm; a.lmbd_stride = 3 * N; a.lmbd_r = d_lr; a.lmbd_r_stride = 1; a.gamma = d_g;
  a.w_out = d_w; a.cost_out = d_c; a.status = d_st; a.iters = d_it; a.max_iter = 200; a.tol = 1e-11;
  void* flush;
  cudaMalloc(&flush, 256 << 20);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  unsigned long long zero[8] = {0};
  float best = 1e9f, best_empty = 1e9f;
  if (rotate) reps = 2 * R;
  for (int r = 0; r < reps + 2; ++r) {
    if (!rotate) cudaMemsetAsync(flush, r, 256 << 20);
    cudaMemcpyToSymbol(lompc::g_warp_prof, zero, sizeof(zero));
    cudaDeviceSynchronize();
    {
      const int q = r % R;
      a.lmbd = d_lm + lm.size() * q; a.lmbd_r = d_lr + (size_t)B * q; a.gamma = d_g + (size_t)B * q;
      a.w_out = d_w + (size_t)B * N * q; a.cost_out = d_c + (size_t)B * q;
    }
    cudaEventRecord(e0);
    lompc::lompc_solve_warp_kernel<N, SPL><<<wa.total_warps, 32>>>(wa);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r >= 2 && ms < best) best = ms;
    if (!rotate) cudaMemsetAsync(flush, r, 256 << 20);
    cudaEventRecord(e0);
    empty_kernel<<<wa.total_warps, 32>>>();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    if (r >= 2 && ms < best_empty) best_empty = ms;
  }
  unsigned long long p[8];
  cudaMemcpyFromSymbol(p, lompc::g_warp_prof, sizeof(p));
  std::vector<int32_t> st(B), it(B);
  cudaMemcpy(st.data(), d_st, B * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(it.data(), d_it, B * 4, cudaMemcpyDeviceToHost);
  int bad = 0, itmax = 0; double itsum = 0;
  for (int i = 0; i < B; ++i) { bad += st[i] != 0; itmax = it[i] > itmax ? it[i] : itmax; itsum += it[i]; }
  const double w = (double)p[5];
  printf("%s N=%d SPL=%d %s B=%d: kernel %.2f us (empty launch %.2f us), iters mean %.2f max %d, bad %d | per warp: setup %.0f "
         "A %.0f B %.0f C %.0f out %.0f cycles; loop trips %.2f, C passes %.2f (%.2f per trip) | per trip: A %.0f B %.0f C %.0f\n",
         rotate ? "[rotating inputs, no flush]" : "[L2 flushed]", N, SPL, large ? "large" : "small", B, best * 1e3, best_empty * 1e3, itsum / B, itmax, bad, p[0] / w, p[1] / w, p[2] / w,
         p[3] / w, p[4] / w, p[6] / w, p[7] / w, (double)p[7] / (p[6] - w), p[1] / (double)p[6], p[2] / (double)(p[6] - w),
         p[3] / (double)(p[6] - w));
  cudaFree(d_lm); cudaFree(d_lr); cudaFree(d_g); cudaFree(d_w); cudaFree(d_c); cudaFree(d_st); cudaFree(d_it); cudaFree(flush);
}

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 512;
  for (int large = 0; large < 2; ++large) {
    run<24, 3>(B, large, 10, 1);
    run<24, 3>(B, large, 10);
    run<12, 3>(B, large, 10);
    run<48, 3>(B, large, 10);
    run<96, 3>(B, large, 10);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
