// K1: batched lower-level MPC QP solve, one QP per thread (sm_100a, fp64).
//
// Replaces LoMPC.solve_lompc (reference lompc.py:137-156) for a whole batch.
// The QP (lompc.py:92-135) is, per EV,
//
//   min  sum_k [ 1/2 d_k w_k^2 + g_k w_k + psi(w_k) ] + c/2 sum_k (s_k - gamma)^2
//   s.t. s_k = s_{k-1} + w_k  (s = A w, A = tril(ones), lompc.py:69),  0 <= w_k <= w_max
//
// with d_k = 2(lmbd_r theta^2 + q lmbd3_k) [+ 2 theta^2/0.81 small EV],
// g_k = theta (lmbd1_k - lmbd2_k), psi = the large-EV pwl (convex, 4 pieces).
// That is a scalar-state optimal-control problem, so every linear solve is an
// O(N) scalar Riccati recursion instead of a dense factorisation.
//
// Algorithm (exact, finite): an active-set Newton method built from two sweeps
// per iteration.
//   backward sweep  k = N-1..0 : costate p_k = c sum_{j>=k}(s_j - gamma), gradient
//       q_k = d_k w_k + g_k + p_k, per-coordinate KKT test (which also picks
//       the working segment / binding set), Riccati recursion of the
//       equality-constrained problem on that working set -> gains (K, kappa).
//   forward sweep   k = 0..N-1 : stage-optimal rollout -- each w_k is the exact
//       minimiser of  stage cost + quadratic cost-to-go  over [0, w_max]
//       (a fused clip for the box, a min/max ladder for the pwl kinks), and the
//       objective of the rollout is accumulated on the fly.
// The rollout is accepted iff it does not raise the objective (to fp64 resolution).  A
// rejected rollout is retried with a proximal (Levenberg-Marquardt) term mu/2 |w - w_cur|^2
// added to every stage (d_k -> d_k + mu, g_k -> g_k - mu w_k in the Riccati model only):
// mu starts at 4c, grows x4 per rejection and is reset to 0 by an acceptance.  For mu
// above the Lipschitz constant the regularised rollout is a descent step, so the
// iteration cannot cycle, and accepted and rejected iterations run the SAME two sweeps
// (no divergent fallback path inside a warp).  It stops when the KKT residual of the
// backward sweep is below tol * scale, i.e. at the exact optimum of the active face.
#pragma once
#include "lompc_common.cuh"

namespace lompc {

// Shared-memory layout: per-thread columns, element k of thread t at [k*T + t]
// (conflict-free: a warp touches 32 consecutive doubles).
template <int NSEG>
struct SmemLayout {
  static constexpr int kArrays = (NSEG > 1) ? 7 : 6;  // D,G,KK,KAP,W0,W1 (+INV)
  __host__ __device__ static size_t bytes(int N, int T) {
    return (size_t)kArrays * N * T * sizeof(double);
  }
};

template <int NSEG>
__device__ __forceinline__ double pwl_value(const Consts& cs, double x) {
  double psi = 0.0;
#pragma unroll
  for (int j = 1; j < NSEG; ++j) psi += (cs.slope[j] - cs.slope[j - 1]) * fmax(x - cs.brk[j], 0.0);
  return psi;
}

template <int NSEG>
__global__ void __launch_bounds__(128) lompc_solve_kernel(const Consts cs, const SolveArgs a) {
  extern __shared__ double smem[];
  const int T = blockDim.x;
  const int t = threadIdx.x;
  const int N = cs.N;
  const int64_t b = (int64_t)blockIdx.x * T + t;
  const bool live = b < a.B;

  double* D = smem + t;
  double* G = D + (size_t)N * T;
  double* KK = G + (size_t)N * T;
  double* KAP = KK + (size_t)N * T;
  double* WA = KAP + (size_t)N * T;
  double* WB = WA + (size_t)N * T;
  double* INV = WB + (size_t)N * T;  // only touched when NSEG > 1
  if (!live) return;  // no block-level sync below

  // ---- problem data (lompc.py:101-135 restated) -------------------------------
  const int64_t row = a.group_of ? (int64_t)a.group_of[b] : b;
  if (a.skip && a.skip[row]) return;
  const double* lm = a.lmbd + row * a.lmbd_stride;
  const double lr = a.lmbd_r[row * a.lmbd_r_stride];
  const double gam = a.gamma[b];
  int st = LOMPC_ST_OK;
  if (!(gam >= 0.0) || !(lr >= 0.0)) st = LOMPC_ST_NEGATIVE;  // nonneg parameters, lompc.py:78-82 (NaN is not nonneg)
  double l2sum = 0.0, gmax = 0.0;
  for (int k = 0; k < N; ++k) {
    const double l1 = lm[k], l2 = lm[N + k], l3 = lm[2 * N + k];
    if (!(l1 >= 0.0) || !(l2 >= 0.0) || !(l3 >= 0.0)) st = LOMPC_ST_NEGATIVE;
    const double g = cs.theta * (l1 - l2);
    const double d = 2.0 * (lr * cs.theta2 + cs.q_scale * l3) + cs.d_base;
    D[k * T] = d;
    G[k * T] = g;
    WA[k * T] = 0.0;
    gmax = fmax(gmax, fabs(g));
    l2sum += l2;
  }
  if (gam > cs.y_max) st = LOMPC_ST_BAD_GAMMA;  // lompc.py:87 (asserted before the parameters are set)
  const double c = cs.c;
  const double wmax = cs.w_max;
  const double gscale = fmax(1.0, gmax + c * N * cs.y_max);
  const double tq = a.tol * gscale;
  const int tqh = __double2hiint(tq);
  const double cg = c * gam;
  // A coordinate within `band` of a breakpoint is treated as sitting on it (the
  // epsilon-binding set of projected Newton: without it a coordinate 1 ulp off a
  // bound, pushed towards it, would stay "free" and stall the search).
  const double band = 1e-9 * wmax;
  // Objective values closer than ~1e-15 of the summed magnitudes cannot be ordered in fp64 (fixed part for
  // the linear / tracking / pwl terms, 1e-15 (|f| + |fn|) for the quadratic ones; see lompc_solve_reg.cuh).
  const double fbase = 1e-15 * (c * N * cs.y_max * cs.y_max + N * wmax * (gmax + cs.slope[NSEG - 1]));
  double brk[NSEG + 1], slope[NSEG], blo[NSEG + 1], bhi[NSEG + 1];
#pragma unroll
  for (int i = 0; i <= NSEG; ++i) {
    brk[i] = cs.brk[i];
    blo[i] = brk[i] - band;
    bhi[i] = brk[i] + band;
  }
#pragma unroll
  for (int j = 0; j < NSEG; ++j) slope[j] = cs.slope[j];

  double* W = WA;   // current feasible iterate
  double* WN = WB;  // candidate
  double sN = 0.0;  // s_{N-1} of the current iterate
  double f = 0.5 * c * N * gam * gam;  // objective at w = 0 (without kappa0)
  int vh = 0;       // high word of the largest KKT violation of the last backward sweep
  double mu = 0.0;  // proximal weight of the safeguard
  int it = 0;
  bool converged = (st != LOMPC_ST_OK);  // invalid input: report, output zeros

  // The two sweeps are those of lompc_solve_reg.cuh (same formulas, see the comments there: Riccati recursion in
  // homogeneous form with the reciprocal off the chain, branch-free KKT test, violation maximum on the integer
  // pipe, tree-shaped pwl minimiser, negated gains) with the per-stage vectors in shared memory and run-time N.
  for (; !converged && it < a.max_iter; ++it) {
    // ---------------- backward sweep ----------------
    double pa = 0.0, pb = 1.0, pr = 0.0, p = 0.0, e = sN - gam;
    vh = 0;
    for (int k = N - 1; k >= 0; --k) {
      const double wk = W[k * T], dk = D[k * T], gk = G[k * T];
      p = fma(c, e, p);
      const double q = fma(dk, wk, gk) + p;
      const double mq = -q;
      double s_hi = slope[0], s_lo = slope[0];
      bool atbp = false;
#pragma unroll
      for (int j = 1; j < NSEG; ++j) {
        const bool ge = wk >= blo[j], gt = wk > bhi[j];
        if (ge) s_hi = slope[j];
        if (gt) s_lo = slope[j];
        atbp |= (ge != gt);
      }
      const bool top = wk >= blo[NSEG], bot = wk <= band;
      if (top) s_hi = 1e300;
      if (bot) s_lo = -1e300;
      atbp |= top | bot;
      const double va = mq - s_hi, vb = s_lo - mq;
      const bool right = va > tq, left = vb > tq;
      const bool binding = atbp && !right && !left;
      const double sl = left ? s_lo : s_hi;
      vh = max(vh, max(__double2hiint(va), __double2hiint(vb)));
      const double dm = dk + mu;  // proximal model: d + mu, g - mu w
      const double gm = fma(-mu, wk, gk);
      const double tq_ = fma(c, pb, pa);
      const double tu = fma(-cg, pb, pr);
      const double bn = fma(dm, pb, tq_);
      const double ib = fast_rcp(bn);
      KK[k * T] = -(tq_ * ib);
      KAP[k * T] = -(fma(gm, pb, tu) * ib);
      if (NSEG > 1) INV[k * T] = pb * ib;
      if (binding) {
        pa = tq_;
        pr = fma(tq_, wk, tu);
      } else {
        pa = dm * tq_;
        pr = fma(dm, tu, -tq_ * (gm + sl));
        pb = bn;
      }
      if (pb > 0x1p600) {  // keep the homogeneous triple in range (exact rescale)
        pa *= 0x1p-600;
        pb *= 0x1p-600;
        pr *= 0x1p-600;
      }
      e -= wk;
    }
    if (vh < tqh) {
      converged = true;
      break;
    }
    // ---------------- forward sweep: stage-optimal rollout ----------------
    double fn = 0.0, s = 0.0;
    for (int k = 0; k < N; ++k) {
      const double kk = KK[k * T], kap = KAP[k * T];
      double x;
      if (NSEG > 1) {
        const double inv = INV[k * T];
        double m[NSEG];
#pragma unroll
        for (int j = 0; j < NSEG; ++j) m[j] = dmin2(fma(kk, s, fma(-slope[j], inv, kap)), brk[j + 1]);
#pragma unroll
        for (int h = 1; h < NSEG; h *= 2) {
#pragma unroll
          for (int j = 0; j + h < NSEG; j += 2 * h) m[j] = dmax2(m[j], m[j + h]);
        }
        x = dpos(m[0]);
      } else {
        const double x0 = fma(kk, s, kap);
        x = x0 > wmax ? wmax : x0;
        x = x0 < 0.0 ? 0.0 : x;
      }
      WN[k * T] = x;
      s += x;
      const double ee = s - gam;
      fn += x * fma(0.5 * D[k * T], x, G[k * T]) + 0.5 * c * ee * ee;
      if (NSEG > 1) {
#pragma unroll
        for (int j = 1; j < NSEG; ++j) fn += (slope[j] - slope[j - 1]) * dpos(x - brk[j]);
      }
    }
    if (fn <= f + (fbase + 1e-15 * (fabs(f) + fabs(fn)))) {
      double* tmp = W;
      W = WN;
      WN = tmp;
      f = dmin2(f, fn);
      sN = s;
      mu = 0.0;
    } else {
      mu = fmax(4.0 * c, 4.0 * mu);
      if (mu > 1e30) break;  // no representable descent step left: report MAXITER below
    }
  }
  if (!converged && st == LOMPC_ST_OK) st = LOMPC_ST_MAXITER;

  // ---- outputs: w, cost in the reference's form (lompc.py:155) ----------------
  double cost = cs.theta * wmax * l2sum;  // theta * lmbd2 @ w_max, lompc.py:130
  double s = 0.0;
  double* wo = a.w_out ? a.w_out + b * (int64_t)N : nullptr;
  for (int k = 0; k < N; ++k) {
    const double x = W[k * T];
    if (wo) wo[k] = x;
    s += x;
    cost += x * fma(0.5 * D[k * T], x, G[k * T]) + 0.5 * c * s * (s - 2.0 * gam);
    if (NSEG > 1) cost += pwl_value<NSEG>(cs, x);
  }
  if (a.cost_out) a.cost_out[b] = cost;
  if (a.err_out) {  // ||w - w_ref|| in the A_bar metric: v'A'Av = ||cumsum(v)||^2 (price_solver.py:188-194,207)
    const double* wr = a.w_ref + row * (int64_t)N;
    const double kap = lr / cs.delta;
    double cum = 0.0, e2 = 0.0;
    for (int k = 0; k < N; ++k) {
      const double v = W[k * T] - wr[k];
      cum += v;
      e2 += cum * cum + kap * v * v;
    }
    a.err_out[b] = sqrt(e2);
  }
  const double w0 = W[0];
  if (a.w0_out) a.w0_out[b] = w0;
  if (a.price0_out)  // lompc.py:164-170
    a.price0_out[b] = cs.theta * (w0 * lm[0] + (wmax - w0) * lm[N]) + cs.q_scale * w0 * w0 * lm[2 * N] +
                      cs.theta2 * w0 * w0 * lr;
  if (a.status) a.status[b] = st;
  if (a.iters) a.iters[b] = it;
  if (a.kkt_res) a.kkt_res[b] = __hiloint2double(vh, vh ? -1 : 0) / gscale;  // upper bound (2^-20 relative)
}

}  // namespace lompc
