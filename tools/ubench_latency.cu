// Dependent-issue latencies that decide the shape of the warp-cooperative K1 (sm_100a): DFMA, 64-bit SHFL,
// a scan step (SHFL + DFMA), REDUX, LDS broadcast; and the SM-wide SHFL throughput with 1..8 warps per SM.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void lat(double* out, long long* cyc, int iters) {
  const int lane = threadIdx.x & 31;
  double a = 1.0 + lane * 1e-3;
  const double b = 1.0000001, c = 1e-9;
  long long t0, t1;
  __shared__ double sm[64];
  sm[lane] = a; sm[32 + lane] = b;
  __syncwarp();
  // 0: DFMA chain
  t0 = clock64();
  for (int i = 0; i < iters; ++i) a = fma(a, b, c);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  // 1: 64-bit shuffle chain
  t0 = clock64();
  for (int i = 0; i < iters; ++i) a = __shfl_xor_sync(0xffffffffu, a, 1);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[1] = t1 - t0;
  // 2: scan step = shuffle + DFMA
  t0 = clock64();
  for (int i = 0; i < iters; ++i) a = fma(__shfl_xor_sync(0xffffffffu, a, 1), b, c);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[2] = t1 - t0;
  // 3: REDUX (int max)
  int v = lane + (int)a;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) v = __reduce_max_sync(0xffffffffu, v) + lane;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[3] = t1 - t0;
  // 4: LDS chain (address depends on the loaded value)
  int idx = lane;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) idx = ((int)sm[idx & 63]) & 63;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[4] = t1 - t0;
  // 5: DSETP + select chain (max of two doubles)
  double m = a;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { const double x = m + c; m = x > b ? x : b; }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[5] = t1 - t0;
  // 6: DADD chain
  t0 = clock64();
  for (int i = 0; i < iters; ++i) m = m + c;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[6] = t1 - t0;
  // 7: rcp.approx + Newton
  double r = a + 2.0;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(r));
    const double e = fma(-r, y, 1.0); const double t = fma(e, e, e); r = fma(y, t, y) + 2.0;
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[7] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + v + idx + m + r;
}

__global__ void shfl_tp(double* out, long long* cyc, int iters) {
  double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    a0 = __shfl_xor_sync(0xffffffffu, a0, 1); a1 = __shfl_xor_sync(0xffffffffu, a1, 2);
    a2 = __shfl_xor_sync(0xffffffffu, a2, 4); a3 = __shfl_xor_sync(0xffffffffu, a3, 8);
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMallocManaged(&cyc, 64);
  const int iters = 4096;
  const char* names[] = {"DFMA", "SHFL.64 (2 x SHFL)", "SHFL.64 + DFMA", "REDUX.MAX + IADD", "LDS + cvt", "DADD + DSETP + SEL",
                         "DADD", "MUFU.RCP64H + 3 DFMA + DADD"};
  for (int w = 0; w < 2; ++w) { lat<<<1, 32>>>(out, cyc, iters); cudaDeviceSynchronize(); }
  for (int i = 0; i < 8; ++i) printf("latency %-28s %.1f cycles\n", names[i], (double)cyc[i] / iters);
  for (int warps = 1; warps <= 16; warps *= 2) {
    for (int w = 0; w < 2; ++w) { shfl_tp<<<148, 32 * warps>>>(out, cyc, iters); cudaDeviceSynchronize(); }
    printf("SHFL throughput, %2d warps per SM: %.2f cycles per 32-bit SHFL warp-instruction per SM\n", warps,
           (double)*cyc / (iters * 8.0 * warps));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
