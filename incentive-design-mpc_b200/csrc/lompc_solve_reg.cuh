// K1, register-resident variant for compile-time horizons (N = 12, 24): same algorithm
// as lompc_solve.cuh (see the header there), restructured for the B200 SM:
//   * the iterate w, the diagonal d and the linear term g live in REGISTERS (fully
//     unrolled sweeps, no address arithmetic); only the Riccati gains (K, kappa[, inv], stored
//     negated) and the parked previous iterate live in shared memory (column-per-thread
//     [k][T] layout, bank-conflict free): 8 warps per SM;
//   * the FP64 pipe issues one warp instruction every 2 cycles per scheduler whatever the number of
//     active lanes (tools/ubench_fp64_halfwarp.cu), and with 2 warps per scheduler it is the binding
//     resource: everything that is not arithmetic is kept off it (violation maximum as an integer,
//     max(x, 0) by bit masking, breakpoint flag from predicates) and the sweeps are branch-free, so
//     that ptxas schedules across the 24 unrolled stages (tools/sass_sched.py shows the result);
//   * the reciprocal of the Riccati pivot is MUFU.RCP64H + one cubic Newton step instead of the
//     IEEE division sequence, and the gains of stage k are formed while stage k-1 runs;
//   * the thread's own 3N-double price row is read with 16-byte loads (the rows of a warp are 576 B
//     apart, every load touches 32 lines whatever its width) and the result leaves the same way.
#pragma once
#include <type_traits>

#include "lompc_common.cuh"
#include "lompc_solve.cuh"

namespace lompc {

// Keeps the shared-memory loads of one unrolled stage inside that stage: without it the
// compiler hoists all 3N loads of a sweep to its top and spills ~60 doubles per thread.
#define LOMPC_STAGE_FENCE() asm volatile("" ::: "memory")

// T = threads per CTA, MINB = CTAs per SM asked of the register allocator, GREG = keep the
// linear term g in registers as well (fewer shared-memory arrays, more registers).
template <int N, int NSEG, int T, bool GREG>
struct RegSmem {
  static constexpr int kArrays = 3 + (GREG ? 0 : 1) + (NSEG > 1 ? 1 : 0);  // KK, KAP, WN [, G] [, INV]
  // + the CTA's table of subdifferentials [s_lo, s_hi] per piece code (piece_table), after the arrays
  static constexpr int kTab = 2 * (2 * NSEG + 1);
  static constexpr size_t bytes = ((size_t)kArrays * N * T + kTab) * sizeof(double);
};

// Piece code of a coordinate w_k: 2i = sitting on breakpoint i (0: the lower box end, NSEG: the upper one),
// 2j + 1 = inside piece j.  The table holds (s_lo, s_hi) = the subdifferential of the separable term there
// (+-1e300 at the box ends; both = slope_j inside a piece), one double2 per code.  Filled by the first
// 2 NSEG + 1 threads of the CTA; the caller synchronises.
template <int NSEG>
__device__ __forceinline__ void piece_table(const Consts& cs, double* tab, int t) {
  if (t <= 2 * NSEG) {
    const int i = t >> 1;
    double lo, hi;
    if (t & 1) {
      lo = hi = cs.slope[i];
    } else {
      lo = i > 0 ? cs.slope[i - 1] : -1e300;
      hi = i < NSEG ? cs.slope[i] : 1e300;
    }
    tab[2 * t] = lo;
    tab[2 * t + 1] = hi;
  }
}

// The solve itself: one QP per thread, iterate / diagonal / linear term in the caller's
// registers.  `smem_t` = this thread's column of the CTA's shared-memory arrays (base + t).
// `lm` may point to global or shared memory (the fused price loop keeps the group's price
// row in shared memory).  `warm`: W holds a feasible starting point on entry (else W = 0).
// SYNC_SETUP: `lm` points into a staging area that ALIASES the shared-memory arrays (the bulk-copy kernel
// below): the whole CTA passes a barrier between the last read of the price rows and the first write of the gains.
template <int N, int NSEG, int T, bool GREG, bool OPT = true, bool SYNC_SETUP = false>
__device__ __forceinline__ void solve_reg(const Consts& cs, const double* lm, const double lr, const double gam,
                                          const double tol, const int max_iter, const bool warm, const bool vec,
                                          double* smem_t, const double* tab, double (&W)[N],
                                          double (&D)[N], double (&GR)[GREG ? N : 1], double& l2sum_out,
                                          double& gscale_out, double& viol_out, int& st_out, int& it_out) {
  double* KK = smem_t;
  double* KAP = KK + N * T;
  double* WN = KAP + N * T;
  double* GS = WN + N * T;                    // only when !GREG
  double* INV = GS + (GREG ? 0 : N * T);      // only when NSEG > 1
#define LOMPC_G(k) (GREG ? GR[GREG ? (k) : 0] : GS[(k) * T])
  int st = LOMPC_ST_OK;
  if (!(gam >= 0.0) || !(lr >= 0.0)) st = LOMPC_ST_NEGATIVE;  // (NaN is not nonneg either)
  double l2sum = 0.0, gmax = 0.0;
  int dmin_hi = 0x7ff00000;  // high word of min_k d_k (d_k >= 0)
#define LOMPC_STAGE_DATA(k, l1, l2, l3)                                      \
  {                                                                         \
    if (!((l1) >= 0.0) || !((l2) >= 0.0) || !((l3) >= 0.0)) st = LOMPC_ST_NEGATIVE; \
    const double g_ = cs.theta * ((l1) - (l2));                             \
    if (GREG) GR[GREG ? (k) : 0] = g_; else GS[(k) * T] = g_;               \
    D[k] = 2.0 * (lr * cs.theta2 + cs.q_scale * (l3)) + cs.d_base;          \
    dmin_hi = min(dmin_hi, __double2hiint(D[k]));                           \
    if (!warm) W[k] = 0.0;                                                  \
    gmax = dmax2(gmax, fabs(g_));                                           \
    l2sum += (l2);                                                          \
  }
  if (vec && N % 2 == 0) {
    // 16-byte loads of the thread's own price row (the rows of a warp are 3N doubles apart, so every load
    // instruction touches 32 lines whatever its width: half as many instructions, half the LSU wavefronts)
    const double2* lm2 = reinterpret_cast<const double2*>(lm);
#pragma unroll
    for (int k = 0; k < N; k += 2) {
      const double2 a1 = lm2[k / 2], a2 = lm2[(N + k) / 2], a3 = lm2[(2 * N + k) / 2];
      LOMPC_STAGE_DATA(k, a1.x, a2.x, a3.x);
      LOMPC_STAGE_DATA(k + 1, a1.y, a2.y, a3.y);
      LOMPC_STAGE_FENCE();
    }
  } else {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const double l1 = lm[k], l2 = lm[N + k], l3 = lm[2 * N + k];
      LOMPC_STAGE_DATA(k, l1, l2, l3);
      LOMPC_STAGE_FENCE();
    }
  }
#undef LOMPC_STAGE_DATA
  if (SYNC_SETUP) __syncthreads();
  if (gam > cs.y_max) st = LOMPC_ST_BAD_GAMMA;
  const double c = cs.c, wmax = cs.w_max;
  const double gscale = fmax(1.0, gmax + c * N * cs.y_max);
  const double tq = tol * gscale;
  const int tqh = __double2hiint(tq);
  const double cg = c * gam;
  const double band = 1e-9 * wmax;
  // Objective values closer than ~1e-15 of the magnitudes that were summed cannot be ordered in fp64:
  // a fixed part for the linear / tracking / pwl terms plus 1e-15 (|f| + |fn|) for the quadratic ones (NOT
  // a worst-case bound with max_k d_k: the closed-form regulariser can return lmbd3_k ~ 1e11 where w_k ~ 0,
  // and a tolerance scaled by it would accept ascent steps and let the iteration wander).
  const double fbase = 1e-15 * (c * N * cs.y_max * cs.y_max + N * wmax * (gmax + cs.slope[NSEG - 1]));
  double brk[NSEG + 1], slope[NSEG];
#pragma unroll
  for (int i = 0; i <= NSEG; ++i) brk[i] = cs.brk[i];
#pragma unroll
  for (int j = 0; j < NSEG; ++j) slope[j] = cs.slope[j];
  double blo[NSEG + 1], bhi[NSEG + 1];  // breakpoints -+ band
#pragma unroll
  for (int i = 0; i <= NSEG; ++i) {
    blo[i] = brk[i] - band;
    bhi[i] = brk[i] + band;
  }

  double sN = 0.0, f = 0.0, mu = 0.0;
  int vh = 0;
  // Piece codes of the iterate (see piece_table), 4 bits per stage.  They are a by-product of the forward sweep
  // (which branch of its min/max tree produced x_k says where x_k sits), so that the backward sweep's KKT test
  // reads the subdifferential from the table instead of comparing w_k with every breakpoint again; only a
  // caller-provided starting point is classified by comparisons (a coordinate within `band` of a breakpoint
  // counts as sitting on it).  W = 0: every coordinate sits on the lower box end (code 0).
  // (Small EV, one piece: two comparisons per stage are cheaper than the bookkeeping, so no codes there.)
  constexpr bool kCodes = NSEG > 1;
  constexpr int kCW = kCodes ? (4 * N + 31) / 32 : 1;
  unsigned codes[kCW], codes_old[kCW];  // of W, of the iterate parked in WN
#pragma unroll
  for (int i = 0; i < kCW; ++i) codes[i] = codes_old[i] = 0u;
  if (warm) {
    // start from the caller's feasible W (the solution at the previous prices of the price
    // loop): the active set is usually already right and one verification sweep remains
    double s0 = 0.0, f0 = 0.0;
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const double x = dmin2(dpos(W[k]), wmax);
      W[k] = x;
      s0 += x;
      if (kCodes) {
        int cd = 1;  // nge + ngt + 1: 2i on breakpoint i (nge = i, ngt = i - 1), 2j + 1 inside piece j
#pragma unroll
        for (int j = 1; j < NSEG; ++j) cd += (x >= blo[j] ? 1 : 0) + (x > bhi[j] ? 1 : 0);
        if (x >= blo[NSEG]) cd = 2 * NSEG;
        if (x <= band) cd = 0;
        codes[(kCodes ? k : 0) / 8] |= (unsigned)cd << (4 * (k % 8));
      }
      if (!OPT) {  // no optimistic phase: the safeguarded loop starts here and needs the objective of W
        const double e = s0 - gam;
        f0 += x * fma(0.5 * D[k], x, LOMPC_G(k)) + 0.5 * c * e * e;
        if (NSEG > 1) {
#pragma unroll
          for (int j = 1; j < NSEG; ++j) f0 += (slope[j] - slope[j - 1]) * dpos(x - brk[j]);
        }
        LOMPC_STAGE_FENCE();
      }
    }
    sN = s0;
    f = f0;
  } else if (!OPT) {
    f = 0.5 * c * N * gam * gam;  // objective of W = 0
  }
  int it = 0;
  bool converged = (st != LOMPC_ST_OK);

  // ---------------- backward sweep: KKT test + Riccati gains; returns true when W is optimal ----------------
  // Riccati recursion in homogeneous form: P = pa/pb, r = pr/pb.  The numerators and
  // the denominator obey a LINEAR recurrence (2 dependent FMAs per stage); the one
  // reciprocal per stage (1/pb_new, needed only for the gains) is off the dependency chain.
  auto backward = [&](auto prox_tag) -> bool {
    constexpr bool PROX = decltype(prox_tag)::value;
    double pa = 0.0, pb = 1.0, pr = 0.0, p = 0.0, e = sN - gam;  // e = s_k - gamma
    double dl_y = 0.0, dl_bn = 1.0, dl_tq = 0.0, dl_kn = 0.0, dl_pb = 0.0;  // stage k+1's deferred gain data
    vh = 0;  // high word of the largest KKT violation (non-negative doubles order like their high words)
#pragma unroll
    for (int k = N - 1; k >= 0; --k) {
      const double wk = W[k], dk = D[k], gk = LOMPC_G(k);
      p = fma(c, e, p);  // costate: c * sum_{j>=k} (s_j - gamma)
      const double q = fma(dk, wk, gk) + p;
      // Subdifferential [s_lo, s_hi] of the separable term at w_k, a coordinate within `band` of
      // a breakpoint counting as sitting on it (+-1e300 at the box ends).  Interior of a piece:
      // s_lo == s_hi = its slope.  -q outside the interval (by more than the tolerance) moves
      // the coordinate onto the neighbouring piece, inside it the coordinate stays put (binding).
      const double mq = -q;
      double s_lo, s_hi;
      bool atbp;  // sitting on a breakpoint or a box end  (<=> s_lo < s_hi)
      if (kCodes) {
        const unsigned cd = (codes[(kCodes ? k : 0) / 8] >> (4 * (k % 8))) & 15u;
        const double2 sub = reinterpret_cast<const double2*>(tab)[cd];
        s_lo = sub.x;
        s_hi = sub.y;
        atbp = (cd & 1u) == 0u;
      } else {
        const bool top = wk >= blo[NSEG], bot = wk <= band;
        s_hi = top ? 1e300 : slope[0];
        s_lo = bot ? -1e300 : slope[0];
        atbp = top | bot;
      }
      // distance of -q from [s_lo, s_hi]: at most one of va, vb is positive.  Everything below is branch-free
      // (a ladder of ?: compiles to a divergent DSETP -> BRA chain per stage that also splits the sweep into
      // basic blocks), and the running maximum is kept on the integer pipe.
      const double va = mq - s_hi, vb = s_lo - mq;
      const bool right = va > tq, left = vb > tq;
      const bool binding = atbp && !right && !left;
      const double sl = left ? s_lo : s_hi;  // slope of the working piece (unused when binding)
      vh = max(vh, max(__double2hiint(va), __double2hiint(vb)));
      // proximal model of the safeguard: d + mu, g - mu w  (mu == 0 throughout the optimistic phase)
      const double dm = PROX ? dk + mu : dk;
      const double gm = PROX ? fma(-mu, wk, gk) : gk;
      const double tq_ = fma(c, pb, pa);     // Q * pb,  Q = c + P
      const double tu = fma(-cg, pb, pr);    // (r - c gamma) * pb
      const double bn = fma(dm, pb, tq_);    // (dm + Q) * pb
      // Gains of stage k: Q/(dm+Q), (r - c gamma + gm)/(dm+Q), 1/(dm+Q).  Only the seed of the reciprocal is
      // issued here; its Newton step, the three products and the stores are written one stage LATER (software
      // pipelining: the MUFU latency and the 4-deep FMA chain of the refinement hide behind the next stage).
      if (k < N - 1) {
        const double ib = rcp_refine(dl_bn, dl_y);
        KK[(k + 1) * T] = -(dl_tq * ib);   // the gains are stored NEGATED: the rollout is x0 = KK s + KAP
        KAP[(k + 1) * T] = -(dl_kn * ib);  // (a separate negation would sit on the forward chain)
        if (NSEG > 1) INV[(k + 1) * T] = dl_pb * ib;
      }
      dl_y = rcp_seed(bn);
      dl_bn = bn;
      dl_tq = tq_;
      dl_kn = fma(gm, pb, tu);
      dl_pb = pb;
      if (binding) {  // P <- Q, r <- Q w + r - c gamma (denominator unchanged)
        pa = tq_;
        pr = fma(tq_, wk, tu);
      } else {        // P <- Q dm/(dm+Q), r <- (dm (r - c gamma) - Q h)/(dm+Q)
        pa = dm * tq_;
        pr = fma(dm, tu, -tq_ * (gm + sl));
        pb = bn;
      }
      if ((k & 7) == 0 && pb > 0x1p600) {  // keep the homogeneous triple in range (exact rescale)
        pa *= 0x1p-600;
        pb *= 0x1p-600;
        pr *= 0x1p-600;
      }
      e -= wk;
      LOMPC_STAGE_FENCE();
    }
    if (vh < tqh) return true;  // => violation < tol * scale (the comparison of the high words is the stricter one)
    {  // gains of stage 0 (deferred)
      const double ib = rcp_refine(dl_bn, dl_y);
      KK[0] = -(dl_tq * ib);
      KAP[0] = -(dl_kn * ib);
      if (NSEG > 1) INV[0] = dl_pb * ib;
    }
    return false;
  };
  // ---------------- forward sweep: stage-optimal rollout into W (the old iterate is parked in WN) ----------------
  // Returns the objective of the rollout when OBJ, else 0; `s_end` = its final state.
  auto forward = [&](auto obj_tag, double& s_end) -> double {
    constexpr bool OBJ = decltype(obj_tag)::value;
    double fn = 0.0, s = 0.0;
    unsigned ncodes[kCW];
#pragma unroll
    for (int i = 0; i < kCW; ++i) {
      ncodes[i] = 0u;
      codes_old[i] = codes[i];
    }
    // gains of stage k are loaded one stage ahead: the chain of stage k+1 can start as soon as s is known and
    // the objective terms of stage k fill its latencies
    double kk_n = KK[0], kap_n = KAP[0], inv_n = (NSEG > 1) ? INV[0] : 0.0;
#pragma unroll
    for (int k = 0; k < N; ++k) {
      // The state s is the loop-carried dependency of the sweep; everything that does not depend on it is
      // kept off that chain.
      const double kk = kk_n, kap = kap_n;
      const double inv = inv_n;
      if (k + 1 < N) {
        kk_n = KK[(k + 1) * T];
        kap_n = KAP[(k + 1) * T];
        if (NSEG > 1) inv_n = INV[(k + 1) * T];
      }
      double x;
      int cd_new;
      if (NSEG > 1) {
        // minimiser of  stage cost + cost-to-go  over [0, w_max]: with c_j = kk s + kap - slope_j inv (kk, kap: negated gains) the
        // stationary point of piece j,  x = max(0, max_j min(c_j, brk[j+1]))  (brk[NSEG] = w_max): one FMA per
        // candidate on the chain, the mins are independent and the max is a tree.
        // The winner of the tree tells where x sits: candidate c_j itself -> inside piece j (code 2j + 1), its
        // cap brk[j+1] -> on that breakpoint (code 2j + 2), a negative maximum -> clipped to the lower end (0).
        double m[NSEG];
        int mc[NSEG];
#pragma unroll
        for (int j = 0; j < NSEG; ++j) {
          const double cj = fma(kk, s, fma(-slope[j], inv, kap));
          const bool below = cj < brk[j + 1];
          m[j] = below ? cj : brk[j + 1];
          mc[j] = below ? 2 * j + 1 : 2 * j + 2;
        }
#pragma unroll
        for (int h = 1; h < NSEG; h *= 2) {
#pragma unroll
          for (int j = 0; j + h < NSEG; j += 2 * h) {
            const bool first = m[j] > m[j + h];
            m[j] = first ? m[j] : m[j + h];
            mc[j] = first ? mc[j] : mc[j + h];
          }
        }
        const bool neg = __double2hiint(m[0]) < 0;
        x = dpos(m[0]);
        cd_new = neg ? 0 : mc[0];
      } else {
        // clip with both comparisons on x0 (in parallel) instead of a min(max()) chain
        const double x0 = fma(kk, s, kap);
        const bool over = x0 > wmax, under = x0 < 0.0;
        x = over ? wmax : x0;
        x = under ? 0.0 : x;
        cd_new = under ? 0 : (over ? 2 : 1);
      }
      if (kCodes) ncodes[(kCodes ? k : 0) / 8] |= (unsigned)cd_new << (4 * (k % 8));
      WN[k * T] = W[k];  // the current iterate is parked (restored only if the rollout is rejected)
      W[k] = x;
      s += x;
      if (OBJ) {
        const double e = s - gam;
        fn += x * fma(0.5 * D[k], x, LOMPC_G(k)) + 0.5 * c * e * e;
        if (NSEG > 1) {
#pragma unroll
          for (int j = 1; j < NSEG; ++j) fn += (slope[j] - slope[j - 1]) * dpos(x - brk[j]);
        }
      }
      LOMPC_STAGE_FENCE();
    }
#pragma unroll
    for (int i = 0; i < kCW; ++i) codes[i] = ncodes[i];
    s_end = s;
    return fn;
  };

  // Phase 1, optimistic: every rollout is accepted, no objective is evaluated and the proximal terms are
  // compiled out (two thirds of the forward sweep's FP64 work and two FMAs per backward stage).  Only for QPs
  // whose stage costs are all strictly convex (min_k d_k > 0: every small EV; a large EV unless lmbd_r = 0 and some
  // lmbd3_k = 0): there a rollout that raises the objective is rare (none in 1e7 random QPs), and a QP that is not
  // done after kOptimistic iterations continues, from where it is, in the safeguarded loop below.  With a
  // d_k = 0 the stage minimiser jumps between breakpoints and the safeguard earns its keep (closed-loop-scale
  // sparse prices: 7.0 iterations with it, 7.9 without).  The choice is made per WARP (one vote), so that the
  // lanes of a warp never sit in different loops; the EVs of a price-loop group share their prices anyway.
  constexpr int kOptimistic = 10;
  const bool optimistic = OPT && __all_sync(__activemask(), dmin_hi > 0);
  const int n_opt = optimistic ? kOptimistic : 0;
  for (; !converged && it < max_iter && it < n_opt; ++it) {
    if (backward(std::false_type{})) {
      converged = true;
      break;
    }
    double s_end;
    forward(std::false_type{}, s_end);
    sN = s_end;
  }
  // Phase 2, safeguarded: a rollout is accepted iff it does not raise the objective (to fp64 resolution); a
  // rejected one is retried with a proximal weight mu (x4 per rejection, reset by an acceptance).
  if (!converged && it < max_iter) {
    double s0 = 0.0, f0 = 0.0;
    if (OPT) {  // objective of the iterate the optimistic phase ended with (or of the start, if it was skipped)
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const double x = W[k];
      s0 += x;
      const double e = s0 - gam;
      f0 += x * fma(0.5 * D[k], x, LOMPC_G(k)) + 0.5 * c * e * e;
      if (NSEG > 1) {
#pragma unroll
        for (int j = 1; j < NSEG; ++j) f0 += (slope[j] - slope[j - 1]) * dpos(x - brk[j]);
      }
      LOMPC_STAGE_FENCE();
    }
    f = f0;
    }
    for (; it < max_iter; ++it) {
      if (backward(std::true_type{})) {
        converged = true;
        break;
      }
      double s_end;
      const double fn = forward(std::true_type{}, s_end);
      if (fn <= f + (fbase + 1e-15 * (fabs(f) + fabs(fn)))) {
        f = dmin2(f, fn);
        sN = s_end;
        mu = 0.0;
      } else {
#pragma unroll
        for (int k = 0; k < N; ++k) W[k] = WN[k * T];
#pragma unroll
        for (int i = 0; i < kCW; ++i) codes[i] = codes_old[i];
        mu = fmax(4.0 * c, 4.0 * mu);
        if (mu > 1e30) break;
      }
    }
  }
  if (!converged && st == LOMPC_ST_OK) st = LOMPC_ST_MAXITER;
  l2sum_out = l2sum;
  gscale_out = gscale;
  viol_out = __hiloint2double(vh, vh ? -1 : 0);  // upper bound of the last sweep's violation (2^-20 relative)
  st_out = st;
  it_out = it;
#undef LOMPC_G
}

template <int N, int NSEG, int T, int MINB, bool GREG>
__global__ void __launch_bounds__(T, MINB) lompc_solve_reg_kernel(const Consts cs, const SolveArgs a) {
  extern __shared__ double smem[];
  const int t = threadIdx.x;
  double* tab = smem + (size_t)RegSmem<N, NSEG, T, GREG>::kArrays * N * T;
  piece_table<NSEG>(cs, tab, t);
  __syncthreads();
  const int64_t b = (int64_t)blockIdx.x * T + t;
  if (b >= a.B) return;
  const int64_t row = a.group_of ? (int64_t)a.group_of[b] : b;
  if (a.skip && a.skip[row]) return;
  const double* lm = a.lmbd + row * a.lmbd_stride;
  const double lr = a.lmbd_r[row * a.lmbd_r_stride];
  const double gam = a.gamma[b];
  double W[N], D[N], GR[GREG ? N : 1];
  double l2sum, gscale, viol;
  int st, it;
  const bool warm = a.w_init != nullptr;
  if (warm) {
    const double* wi = a.w_init + b * (int64_t)N;
#pragma unroll
    for (int k = 0; k < N; ++k) W[k] = wi[k];
  }
  solve_reg<N, NSEG, T, GREG>(cs, lm, lr, gam, a.tol, a.max_iter, warm, a.vec16 != 0, smem + t, tab, W, D, GR, l2sum,
                              gscale, viol, st, it);
  const double* GS = smem + t + 3 * N * T;
#define LOMPC_G(k) (GREG ? GR[GREG ? (k) : 0] : GS[(k) * T])
  const double c = cs.c, wmax = cs.w_max;
  double slope[NSEG], brk[NSEG + 1];
#pragma unroll
  for (int i = 0; i <= NSEG; ++i) brk[i] = cs.brk[i];
#pragma unroll
  for (int j = 0; j < NSEG; ++j) slope[j] = cs.slope[j];

  // ---- outputs ----
  double cost = cs.theta * wmax * l2sum;
  double s = 0.0;
  double* wo = a.w_out ? a.w_out + b * (int64_t)N : nullptr;
  if (wo) {
    if (a.vec16 && N % 2 == 0) {
      double2* wo2 = reinterpret_cast<double2*>(wo);
#pragma unroll
      for (int k = 0; k < N; k += 2) wo2[k / 2] = make_double2(W[k], W[k + 1]);
    } else {
#pragma unroll
      for (int k = 0; k < N; ++k) wo[k] = W[k];
    }
  }
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const double x = W[k];
    s += x;
    cost += x * fma(0.5 * D[k], x, LOMPC_G(k)) + 0.5 * c * s * (s - 2.0 * gam);
    if (NSEG > 1) {
#pragma unroll
      for (int j = 1; j < NSEG; ++j) cost += (slope[j] - slope[j - 1]) * dpos(x - brk[j]);
    }
    LOMPC_STAGE_FENCE();
  }
  if (a.cost_out) a.cost_out[b] = cost;
  if (a.err_out) {
    const double* wr = a.w_ref + row * (int64_t)N;
    const double kap = lr / cs.delta;
    double cum = 0.0, e2 = 0.0;
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const double v = W[k] - wr[k];
      cum += v;
      e2 += cum * cum + kap * v * v;
      LOMPC_STAGE_FENCE();
    }
    a.err_out[b] = sqrt(e2);
  }
  const double w0 = W[0];
  if (a.w0_out) a.w0_out[b] = w0;
  if (a.price0_out)
    a.price0_out[b] = cs.theta * (w0 * lm[0] + (wmax - w0) * lm[N]) + cs.q_scale * w0 * w0 * lm[2 * N] +
                      cs.theta2 * w0 * w0 * lr;
  if (a.status) a.status[b] = st;
  if (a.iters) a.iters[b] = it;
  if (a.kkt_res) a.kkt_res[b] = viol / gscale;
#undef LOMPC_G
}

// ---------------------------------------------------------------------------------------------------------
// The same kernel with the rows moved by the bulk-copy engine (cp.async.bulk, SASS UBLKCP) instead of per-thread
// loads and stores.  A thread's price row is 3N contiguous doubles, the rows of a warp are 576 B apart: every
// LDG.128 of the kernel above touches 32 lines (32 passes through the L1 tag stage for 512 useful bytes; ncu: 10 M
// of the kernel's 21 M L1 wavefronts are these loads, 3.3 M the result stores, and the L1 data pipe - which the
// shared-memory traffic of the sweeps needs too - is the kernel's second-busiest unit).  Here every thread issues
// ONE bulk copy of its row into shared memory (no L1 pass, no registers, no LSU work for the global side) and
// reads it back with conflict-free LDS.128 (row stride 37 x 16 B: odd); the result row leaves through a
// private 13 x 16 B slot and one bulk store.  The staging area of the price rows ALIASES the Riccati gain arrays:
// a row is dead once d and g are in registers, the gains are first written after that (one CTA barrier in
// between, solve_reg<SYNC_SETUP>), so the shared-memory footprint grows only by the result slots (T x 208 B).
// Plain batches only: one row per QP (no group_of / skip / warm start / fused error outputs), 16-byte aligned.
// OUT_ALIAS: the result slots alias the gain arrays too (one more CTA barrier, after the solve) - for the shape that
// fills the SM's shared memory with ONE CTA, where the threads that finish early have nothing else to do anyway.
template <int N, int NSEG, int T, bool OUT_ALIAS>
struct RegTmaSmem {
  static constexpr int kRowIn = 3 * N * 8 + 16;    // bytes between staged price rows (odd multiple of 16)
  static constexpr int kRowOut = N * 8 + 16;       // bytes between result slots (odd multiple of 16)
  static_assert((kRowIn / 16) % 2 == 1 && (kRowOut / 16) % 2 == 1, "conflict-free 16-byte accesses need odd strides");
  static constexpr size_t kArrayBytes = RegSmem<N, NSEG, T, true>::kArrays * (size_t)N * T * 8;
  static constexpr size_t kStageBytes = (size_t)T * kRowIn;
  static constexpr size_t kRegion = ((kArrayBytes > kStageBytes ? kArrayBytes : kStageBytes) + 15) / 16 * 16;
  static constexpr size_t oTab = kRegion;                                    // piece table
  static constexpr size_t oTabEnd = oTab + RegSmem<N, NSEG, T, true>::kTab * 8;
  static constexpr size_t oOut = OUT_ALIAS ? 0 : oTabEnd;                     // result slots
  static constexpr size_t oBar = OUT_ALIAS ? oTabEnd : oOut + (size_t)T * kRowOut;  // mbarrier
  static_assert(!OUT_ALIAS || (size_t)T * kRowOut <= kRegion, "result slots must fit the aliased region");
  static constexpr size_t bytes = oBar + 16;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int N, int NSEG, int T, int MINB, bool OUT_ALIAS>
__global__ void __launch_bounds__(T, MINB) lompc_solve_reg_tma_kernel(const Consts cs, const SolveArgs a) {
  using L = RegTmaSmem<N, NSEG, T, OUT_ALIAS>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* smem = reinterpret_cast<double*>(smem_raw);
  double* tab = reinterpret_cast<double*>(smem_raw + L::oTab);
  const int t = threadIdx.x;
  const unsigned bar = smem_u32(smem_raw + L::oBar);
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(T));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  piece_table<NSEG>(cs, tab, t);
  __syncthreads();
  const int64_t b = (int64_t)blockIdx.x * T + t;
  const bool active = b < a.B;
  // ---- every thread: announce the bytes of its row, start the copy, arrive
  if (active) {
    const unsigned dst = smem_u32(smem_raw + (size_t)t * L::kRowIn);
    const double* src = a.lmbd + b * (int64_t)(3 * N);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(3 * N * 8) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(3 * N * 8), "r"(bar)
                 : "memory");
  } else {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
  }
  // (Tried: cp.async.bulk.prefetch.L2 of the row the NEXT CTA of this SM slot will load - 188 -> 198 us at 524,288
  // small-EV QPs: the copies of the resident CTAs already keep the memory system busy, the extra requests only queue.)
  // A thread past the end of the batch has to reach the barriers: it repeats a QP and writes nothing - the QP of its
  // warp's first lane, so that the warp's votes (optimistic phase or not) are those of its real QPs.
  const int64_t bw = (int64_t)blockIdx.x * T + (t & ~31);
  const int tq_ = active ? t : (bw < a.B ? (t & ~31) : 0);
  const int64_t bq = (int64_t)blockIdx.x * T + tq_;
  const double lr = a.lmbd_r[bq];
  const double gam = a.gamma[bq];
  {
    unsigned done = 0;
    while (!done)
      asm volatile(
          "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}"
          : "=r"(done)
          : "r"(bar)
          : "memory");
  }
  const double* lm = reinterpret_cast<const double*>(smem_raw + (size_t)tq_ * L::kRowIn);
  double W[N], D[N], GR[N];
  double l2sum, gscale, viol;
  int st, it;
  solve_reg<N, NSEG, T, true, true, true>(cs, lm, lr, gam, a.tol, a.max_iter, false, true, smem + t, tab, W, D, GR, l2sum,
                                          gscale, viol, st, it);
  if (OUT_ALIAS) __syncthreads();  // nobody reads or writes the gain arrays any more
  if (!active) return;
  const double c = cs.c, wmax = cs.w_max;
  double slope[NSEG], brk[NSEG + 1];
#pragma unroll
  for (int i = 0; i <= NSEG; ++i) brk[i] = cs.brk[i];
#pragma unroll
  for (int j = 0; j < NSEG; ++j) slope[j] = cs.slope[j];
  // ---- outputs: the result row through this thread's slot and one bulk store
  double2* slot = reinterpret_cast<double2*>(smem_raw + L::oOut + (size_t)t * L::kRowOut);
#pragma unroll
  for (int k = 0; k < N; k += 2) slot[k / 2] = make_double2(W[k], W[k + 1]);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(a.w_out + b * (int64_t)N),
               "r"(smem_u32(slot)), "r"(N * 8)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  double cost = cs.theta * wmax * l2sum;
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const double x = W[k];
    s += x;
    cost += x * fma(0.5 * D[k], x, GR[k]) + 0.5 * c * s * (s - 2.0 * gam);
    if (NSEG > 1) {
#pragma unroll
      for (int j = 1; j < NSEG; ++j) cost += (slope[j] - slope[j - 1]) * dpos(x - brk[j]);
    }
    LOMPC_STAGE_FENCE();
  }
  if (a.cost_out) a.cost_out[b] = cost;
  if (a.status) a.status[b] = st;
  if (a.iters) a.iters[b] = it;
  if (a.kkt_res) a.kkt_res[b] = viol / gscale;
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the slot must outlive the store's read
}

}  // namespace lompc
