// FP64 roofline denominator: MEASURED_PEAKS.json holds no FP64 figure, so the
// DFMA peak of the device is measured in the same run as the kernels
// (SURVEY.md section 8d).  Register-resident chains, 8 independent per thread.
#include <cuda_runtime.h>

#include "lompc_b200.h"

namespace {

__global__ void __launch_bounds__(256) dfma_chain_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
  double x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456) out[0] = s;  // keeps the chains live, never true in practice
}

}  // namespace

extern "C" int lompc_measure_fp64_peak(int device, int iters, double* tflops_out, double* ms_out) {
  if (!tflops_out || iters < 1) return LOMPC_ERR_ARG;
  if (lompc_device_count() <= device || device < 0) return LOMPC_ERR_NO_DEVICE;
  if (cudaSetDevice(device) != cudaSuccess) return LOMPC_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return LOMPC_ERR_CUDA;
  double* d_out = nullptr;
  if (cudaMalloc(&d_out, 8) != cudaSuccess) return LOMPC_ERR_CUDA;
  const int threads = 256, blocks = prop.multiProcessorCount * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    dfma_chain_kernel<<<blocks, threads>>>(d_out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  if (cudaGetLastError() != cudaSuccess) return LOMPC_ERR_CUDA;
  const double flops = 2.0 * 64.0 * (double)iters * (double)threads * (double)blocks;
  *tflops_out = flops / (best * 1e-3) / 1e12;
  if (ms_out) *ms_out = best;
  return LOMPC_OK;
}
