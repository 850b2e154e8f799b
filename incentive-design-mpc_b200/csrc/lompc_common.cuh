// Shared definitions of the LoMPC kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "lompc_b200.h"

namespace lompc {

constexpr int kMaxSeg = 4;  // pieces of the large-EV pwl (lompc.py:111-115)

// Per-handle constants, passed to kernels by value.
struct Consts {
  int N;
  int large;         // 0 small EV, 1 large EV
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
};

}  // namespace lompc
