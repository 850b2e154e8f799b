// Does the FP64 pipe of sm_100a issue a warp with <= 16 active lanes in one pass instead of two?
// One warp per SMSP (4 warps per CTA, one CTA per SM), 8 independent DFMA chains per thread,
// timed with clock64 for active-lane counts 32 / 16 / 8 and for lanes spread over both half-warps.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k(double* out, long long* cyc, unsigned mask, int iters) {
  const int lane = threadIdx.x & 31;
  double a0 = 1.0 + lane, a1 = 1.1, a2 = 1.2, a3 = 1.3, a4 = 1.4, a5 = 1.5, a6 = 1.6, a7 = 1.7;
  const double b = 1.0000001, c = 1e-9;
  long long t0 = 0, t1 = 0;
  if ((mask >> lane) & 1) {
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
      a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 128 * 8); cudaMallocManaged(&cyc, 8);
  const int iters = 4096;
  unsigned masks[] = {0xffffffffu, 0x0000ffffu, 0x000000ffu, 0x00ff00ffu, 0x55555555u};
  const char* names[] = {"32 lanes", "lower 16", "lower 8", "8+8 across halves", "16 interleaved"};
  for (int w = 0; w < 2; ++w)
    for (int m = 0; m < 5; ++m) {
      k<<<148, 128>>>(out, cyc, masks[m], iters);
      cudaDeviceSynchronize();
      if (w) printf("%-20s %.2f cycles per DFMA warp-instruction (one warp per SMSP)\n", names[m], (double)*cyc / (iters * 8.0));
    }
  return 0;
}
