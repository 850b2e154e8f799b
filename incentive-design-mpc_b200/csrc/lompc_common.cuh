// Shared definitions of the LoMPC kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "lompc_b200.h"

// Shared by the translation units of the library (defined in lompc_api.cu).
namespace lompc_detail {
void count_launch();                              // feeds lompc_launch_count()
int cuda_fail(cudaError_t e, const char* what);   // records the message, returns LOMPC_ERR_CUDA
}  // namespace lompc_detail

namespace lompc {
struct Consts;
struct SolveArgs;
struct WarpArgs;
}  // namespace lompc

namespace lompc_detail {
// What the other translation units need to know of a handle (defined in lompc_api.cu).
struct HandleView {
  const lompc::Consts* cs;
  int device;
  int max_iter;
  double tol;
  int variant;
};
HandleView handle_view(const lompc_t* h);
// K1 with the handle's automatic kernel choice (lompc_api.cu).
int launch_k1(const lompc_t* h, const lompc::SolveArgs& a, cudaStream_t s);
// Warp-cooperative K1 (lompc_warp_api.cu).  spl = stages per lane (3 or 6); fills warp_begin / total_warps.
bool warp_kernel_supports(int N, int spl);
int launch_k1_warp(int device, int N, int spl, lompc::WarpArgs& wa, cudaStream_t s);
}  // namespace lompc_detail

namespace lompc {

constexpr int kMaxSeg = 4;  // pieces of the large-EV pwl (lompc.py:111-115)

// Per-handle constants, passed to kernels by value.
struct Consts {
  int N;
  int large;         // 0 small EV, 1 large EV
  double delta;      // relative weight of the charging cost
  double theta;      // battery capacity
  double w_max;      // box upper bound (lompc.py:93)
  double y_max;      // gamma <= y_max (lompc.py:87)
  double c;          // 2*delta*theta^2 = strong convexity modulus m (lompc.py:71)
  double q_scale;    // 3*theta/(4*w_max) (lompc.py:67)
  double theta2;     // theta^2
  double d_base;     // small EV: 2*theta^2/0.81 (lompc.py:107); large: 0
  // Breakpoints b[0..nseg] of the separable term and the slope (per unit w) of
  // the pwl on each segment; small EV: one segment [0, w_max] with slope 0.
  int nseg;
  double brk[kMaxSeg + 1];
  double slope[kMaxSeg];
};

// Reciprocal of a positive, well-scaled double (Riccati pivots): MUFU.RCP64H seed + one cubic
// Newton step, off the IEEE division's slow path.
__device__ __forceinline__ double fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));  // ~2^-23 relative error
  const double e = fma(-x, y, 1.0);
  const double t = fma(e, e, e);  // one cubic step: error e^3 ~ 2^-69
  return fma(y, t, y);
}

// The two halves of fast_rcp, for callers that want other work between the MUFU and the refinement.
__device__ __forceinline__ double rcp_seed(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
}
__device__ __forceinline__ double rcp_refine(double x, double y) {
  const double e = fma(-x, y, 1.0);
  const double t = fma(e, e, e);
  return fma(y, t, y);
}

// max / min of two non-NaN doubles: one DSETP + a select.  (fmax / fmin quiet NaNs, which costs
// ~8 SASS instructions per call in FP64; a NaN here propagates and ends as LOMPC_ST_MAXITER.)
__device__ __forceinline__ double dmax2(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double dmin2(double a, double b) { return a < b ? a : b; }
// max(x, 0) on the integer pipe: clears every bit when the sign bit is set (-0.0 -> +0.0, NaN stays NaN).
// `x > 0.0 ? x : 0.0` is canonicalised to max.f64 by the compiler, which ptxas expands to DSETP.MAX + selects
// + a NaN fix-up (8 instructions, one of them on the FP64 pipe); this is a shift and two ANDs.
__device__ __forceinline__ double dpos(double x) {
  const int hi = __double2hiint(x), lo = __double2loint(x);
  const int keep = ~(hi >> 31);
  return __hiloint2double(hi & keep, lo & keep);
}

struct SolveArgs {
  int64_t B;
  const double* lmbd;
  int64_t lmbd_stride;
  const double* lmbd_r;
  int64_t lmbd_r_stride;
  const double* gamma;
  double* w_out;
  double* cost_out;
  int32_t* status;
  int32_t* iters;
  double* kkt_res;
  int max_iter;
  double tol;
  // ---- group mode (price loop, price_solver.py:196-214,272-285); all nullable ----
  const int32_t* group_of;  // [B] row of lmbd / lmbd_r / w_ref used by QP b (NULL: row = b)
  const int32_t* skip;      // [rows] QPs whose row has skip != 0 are not solved (converged groups)
  const double* w_ref;      // [rows, N] reference trajectories for err_out
  double* err_out;          // [B] sqrt((w-w_ref)' (A'A + lmbd_r/delta I) (w-w_ref)), price_solver.py:207
  double* w0_out;           // [B] first-step charge w[0]
  double* price0_out;       // [B] LoMPC.get_price0 (lompc.py:164-170)
  const double* w_init;     // [B,N] feasible starting points (the previous solutions of the price loop;
                            // register kernel only) or NULL: start from w = 0
  int vec16;                // register kernel: lmbd rows and w_out rows are 16-byte aligned (set by the launcher)
};

// ---- warp-cooperative K1 (lompc_solve_warp.cuh): one launch serves up to kMaxWarpSegs segments (EV types) ----
constexpr int kMaxWarpSegs = 4;  // == LOMPC_SET_MAX_SEGMENTS

struct WarpSeg {
  Consts cs;
  SolveArgs a;
  int warp_begin;  // first warp of the launch that belongs to this segment
};

struct WarpArgs {
  int nsegs;
  int total_warps;
  const unsigned long long* epoch_src;  // device word holding the call's epoch (NULL: epoch 0)
  unsigned long long* summary;          // atomicMax(epoch * 4 + worst status of the launch); NULL: no summary
  WarpSeg seg[kMaxWarpSegs];
};

}  // namespace lompc
