// K1, latency variant: one QP per LANE GROUP of a warp (time-parallel sweeps), for small and
// medium batches where one-QP-per-thread (lompc_solve_reg.cuh) cannot fill the GPU.
//
// Same algorithm as lompc_solve.cuh / lompc_solve_reg.cuh (active-set Newton: backward KKT test +
// Riccati recursion, forward stage-optimal rollout, optimistic phase, then the safeguarded phase
// with the proximal retry) and the same decisions (epsilon-band, tolerances, piece codes), so the
// iteration counts agree with the thread kernels; but the horizon is cut into LPQ = N / SPL blocks
// of SPL consecutive stages, one block per lane, and every sweep over the horizon becomes
//   a serial pass over the SPL stages a lane owns  +  a log2(LPQ)-step scan across the group:
//   * states s_k (prefix sum of w) and costates p_k = c sum_{j>=k}(s_j - gamma), written as
//     c [(N-k)(s_k - gamma) + sum_{i>k}(N-i) w_i] so that BOTH scans start at once;
//   * the Riccati recursion in its homogeneous form (pa, pb, pr), which is LINEAR: stage k is a
//     3x3 matrix [[A, 0], [v', m]] (7 entries, power-of-two normalised), a lane composes the SPL
//     matrices of its block, the group runs a suffix scan of block operators, and each lane then
//     rolls the triple through its own stages to form the gains (reciprocals off every chain);
//   * the forward rollout s_k = s_{k-1} + x_k(s_{k-1}) is piecewise affine in the state: with a
//     GUESS of the piece every stage lands on (the piece codes of the iterate, moved one piece by
//     the KKT test) a block is an affine map, the group scans the maps, every lane replays its SPL
//     stages EXACTLY from the scanned entry state and reports the pieces it really landed on; the
//     guess is replaced by them until nothing changes (lane l is final after l + 1 passes at the
//     latest; one or two passes in practice, one at convergence).
// N = 24, SPL = 3: 8 lanes per QP, 4 QPs per warp, no shared memory, ~60 registers.  The groups of
// a warp iterate in lock step until all of them are done (finished groups are predicated off).
// One launch serves up to kMaxWarpSegs segments (EV types) -- a warp belongs to one segment.
#pragma once
#include "lompc_common.cuh"

namespace lompc {

constexpr unsigned kFullMask = 0xffffffffu;

// Per-phase cycle counters of tools/prof_warp.cu (compiled out of the library).
#ifdef LOMPC_WARP_PROF
__device__ unsigned long long g_warp_prof[8];  // setup, A, B, C, epilogue cycles; warps; loop trips; C passes
#define LOMPC_PROF_T(var) const long long var = clock64()
#define LOMPC_PROF_ADD(i, v) prof[i] += (unsigned long long)(v)
#else
#define LOMPC_PROF_T(var)
#define LOMPC_PROF_ADD(i, v)
#endif

template <int LPQ>
__device__ __forceinline__ double grp_sum(double v) {
#pragma unroll
  for (int d = LPQ / 2; d >= 1; d >>= 1) v += __shfl_xor_sync(kFullMask, v, d, LPQ);
  return v;
}
template <int LPQ>
__device__ __forceinline__ int grp_max_int(int v) {
#pragma unroll
  for (int d = LPQ / 2; d >= 1; d >>= 1) v = max(v, __shfl_xor_sync(kFullMask, v, d, LPQ));
  return v;
}
template <int LPQ>
__device__ __forceinline__ int grp_min_int(int v) {
#pragma unroll
  for (int d = LPQ / 2; d >= 1; d >>= 1) v = min(v, __shfl_xor_sync(kFullMask, v, d, LPQ));
  return v;
}
template <int LPQ>
__device__ __forceinline__ double grp_max_nonneg(double v) {  // v >= 0
#pragma unroll
  for (int d = LPQ / 2; d >= 1; d >>= 1) v = dmax2(v, __shfl_xor_sync(kFullMask, v, d, LPQ));
  return v;
}
template <int LPQ>
__device__ __forceinline__ bool grp_any(bool p, int lane) {
  const unsigned bal = __ballot_sync(kFullMask, p);
  const unsigned gmask = (LPQ == 32 ? kFullMask : ((1u << LPQ) - 1u)) << (lane & ~(LPQ - 1));
  return (bal & gmask) != 0u;
}

// Stage operator of the homogeneous Riccati recursion, x_k = M_k x_{k+1}, x = (pa, pb, pr)':
//   [ a11 a12 0 ]
//   [ a21 a22 0 ]
//   [ v1  v2  m ]
struct RicMat {
  double a11, a12, a21, a22, v1, v2, m;
};
// L * R (L = the earlier stage, applied after R)
__device__ __forceinline__ RicMat ric_mul(const RicMat& L, const RicMat& R) {
  RicMat o;
  o.a11 = fma(L.a11, R.a11, L.a12 * R.a21);
  o.a12 = fma(L.a11, R.a12, L.a12 * R.a22);
  o.a21 = fma(L.a21, R.a11, L.a22 * R.a21);
  o.a22 = fma(L.a21, R.a12, L.a22 * R.a22);
  o.v1 = fma(L.m, R.v1, fma(L.v1, R.a11, L.v2 * R.a21));
  o.v2 = fma(L.m, R.v2, fma(L.v1, R.a12, L.v2 * R.a22));
  o.m = L.m * R.m;
  return o;
}
__device__ __forceinline__ RicMat ric_shfl_down(const RicMat& x, int d, int width) {
  RicMat o;
  o.a11 = __shfl_down_sync(kFullMask, x.a11, d, width);
  o.a12 = __shfl_down_sync(kFullMask, x.a12, d, width);
  o.a21 = __shfl_down_sync(kFullMask, x.a21, d, width);
  o.a22 = __shfl_down_sync(kFullMask, x.a22, d, width);
  o.v1 = __shfl_down_sync(kFullMask, x.v1, d, width);
  o.v2 = __shfl_down_sync(kFullMask, x.v2, d, width);
  o.m = __shfl_down_sync(kFullMask, x.m, d, width);
  return o;
}

// 2^-e for x = f 2^e (x positive, finite, normal): an exact scale that brings x into [1, 2).
__device__ __forceinline__ double pow2_inv_scale(double x) {
  const int be = (__double2hiint(x) >> 20) & 0x7ff;
  return __hiloint2double((2046 - be) << 20, 0);
}

// p ? a : b as ONE select: a C++ ternary whose arms are cheap still compiles to a divergent branch around the
// unselected arm in several places of the sweeps (BSSY / BRA / BSYNC per stage); a selp cannot.
__device__ __forceinline__ double dsel(bool p, double a, double b) {
  double r;
  asm("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %3, 0;\n\tselp.f64 %0, %1, %2, q;\n\t}" : "=d"(r) : "d"(a), "d"(b), "r"((int)p));
  return r;
}

// v[i] for a run-time i in 0..3 by a select tree (a run-time index into a register array would go through local
// memory, one into the kernel parameters through the constant cache: ~40 cycles on the critical path either way).
__device__ __forceinline__ double sel4(int i, double v0, double v1, double v2, double v3) {
  const bool b0 = (i & 1) != 0;
  return dsel((i & 2) != 0, dsel(b0, v3, v2), dsel(b0, v1, v0));
}

// Slopes / breakpoints of the separable term in registers (NSEG = 1: small EV, NSEG = 4: the large-EV pwl).
template <int NSEG>
struct Pwl {
  double slope[NSEG];
  double brk[NSEG + 1];
  __device__ __forceinline__ explicit Pwl(const Consts& cs) {
#pragma unroll
    for (int i = 0; i < NSEG; ++i) slope[i] = cs.slope[i];
#pragma unroll
    for (int i = 0; i <= NSEG; ++i) brk[i] = cs.brk[i];
  }
  __device__ __forceinline__ double slope_at(int j) const {  // 0 <= j < NSEG
    if (NSEG == 4) return sel4(j, slope[0], slope[1], slope[2], slope[3]);
    return slope[0];
  }
  __device__ __forceinline__ double brk_at(int i) const {  // 0 <= i <= NSEG
    if (NSEG == 4) return dsel(i == 4, brk[NSEG], sel4(i, brk[0], brk[1], brk[2], brk[3]));
    return dsel(i != 0, brk[NSEG], brk[0]);
  }
};

// What one lane group solves: pointers may be global or shared memory (the fused price loop keeps prices, warm
// starts and solutions in shared memory).
struct WarpProblem {
  const double* lm;      // [3N] prices
  double lr, gam;        // lmbd_r, gamma
  const double* w_init;  // [N] feasible starting point (the solution at the previous prices) or NULL: start at 0
  double* w_out;         // [N]
  double* cost_out;      // scalars of this QP, each may be NULL
  int32_t* status;
  int32_t* iters;
  double* kkt_res;
  unsigned char* codes_out;  // [N] piece codes of the solution (see lompc_solve_reg.cuh: piece_table) or NULL
  double tol;
  int max_iter;
};

// The solve of one lane group.  `li` = lane index inside the group (owns stages li*SPL .. li*SPL+SPL-1),
// `live` = the group has a QP (idle groups stay in the shuffles of their warp).
template <int N, int NSEG, int SPL>
__device__ __forceinline__ void solve_warp_core(const Consts& cs, const WarpProblem& P, const bool live, const int lane,
                                                int& st_out) {
  constexpr int LPQ = N / SPL;
  static_assert(N % SPL == 0 && (LPQ & (LPQ - 1)) == 0 && LPQ <= 32 && LPQ >= 1, "N = SPL * 2^m, at most 32 lanes");
  static_assert(NSEG == 1 || NSEG == 4, "small EV (one piece) or the large-EV pwl (four)");
  const int li = lane & (LPQ - 1);
  const int k0 = li * SPL;
  const double c = cs.c, wmax = cs.w_max;
  const Pwl<NSEG> pw(cs);
#ifdef LOMPC_WARP_PROF
  unsigned long long prof[8] = {0, 0, 0, 0, 0, 1, 0, 0};
#endif
  LOMPC_PROF_T(t_begin);

  // ---- problem data: this lane's SPL stages of the three price segments (lompc.py:101-135) ----
  const double* lm = P.lm;
  const double lr = live ? P.lr : 0.0;
  const double gam = live ? P.gam : 0.0;
  const bool warm = P.w_init != nullptr;
  double W[SPL], D[SPL], G[SPL], WN[SPL];
  int CD[SPL], CDO[SPL];  // piece codes of W / of the parked iterate (see lompc_solve_reg.cuh: piece_table)
  bool neg = !(gam >= 0.0) || !(lr >= 0.0);  // nonneg parameters, lompc.py:78-82 (NaN is not nonneg)
  double l2loc = 0.0, gmaxloc = 0.0;
  int dmin_hi = 0x7ff00000;
#pragma unroll
  for (int j = 0; j < SPL; ++j) {
    double l1 = 0.0, l2 = 0.0, l3 = 0.0;
    if (live) {
      l1 = lm[k0 + j];
      l2 = lm[N + k0 + j];
      l3 = lm[2 * N + k0 + j];
    }
    neg |= !(l1 >= 0.0) || !(l2 >= 0.0) || !(l3 >= 0.0);
    G[j] = cs.theta * (l1 - l2);
    D[j] = 2.0 * (lr * cs.theta2 + cs.q_scale * l3) + cs.d_base;
    dmin_hi = min(dmin_hi, __double2hiint(D[j]));
    W[j] = 0.0;
    WN[j] = 0.0;
    CD[j] = 0;
    CDO[j] = 0;
    gmaxloc = dmax2(gmaxloc, fabs(G[j]));
    l2loc += l2;
  }
  if (warm && live) {
    // feasible starting point + its piece codes by comparison (a coordinate within `band` of a breakpoint counts
    // as sitting on it), as in lompc_solve_reg.cuh
    const double bandw = 1e-9 * wmax;
    const double* wi = P.w_init + k0;
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      const double x = dmin2(dpos(wi[j]), wmax);
      W[j] = x;
      int cd = 1;
      if (NSEG > 1) {
#pragma unroll
        for (int i = 1; i < NSEG; ++i) cd += (x >= pw.brk[i] - bandw ? 1 : 0) + (x > pw.brk[i] + bandw ? 1 : 0);
      }
      if (x >= pw.brk[NSEG] - bandw) cd = 2 * NSEG;
      if (x <= bandw) cd = 0;
      CD[j] = cd;
    }
  }
  // one butterfly for the three group reductions (independent chains share the shuffle latencies)
#pragma unroll
  for (int d = LPQ / 2; d >= 1; d >>= 1) {
    const double t1 = __shfl_xor_sync(kFullMask, l2loc, d, LPQ);
    const double t2 = __shfl_xor_sync(kFullMask, gmaxloc, d, LPQ);
    const int t3 = __shfl_xor_sync(kFullMask, dmin_hi, d, LPQ);
    l2loc += t1;
    gmaxloc = dmax2(gmaxloc, t2);
    dmin_hi = min(dmin_hi, t3);
  }
  const double l2sum = l2loc, gmax = gmaxloc;
  int st = LOMPC_ST_OK;
  if (grp_any<LPQ>(neg, lane)) st = LOMPC_ST_NEGATIVE;
  if (gam > cs.y_max) st = LOMPC_ST_BAD_GAMMA;
  const double gscale = fmax(1.0, gmax + c * N * cs.y_max);
  const double tq = P.tol * gscale;
  const int tqh = __double2hiint(tq);
  const double cg = c * gam;
  const double band = 1e-9 * wmax;
  const double fbase = 1e-15 * (c * N * cs.y_max * cs.y_max + N * wmax * (gmax + pw.slope[NSEG - 1]));
  const double top_lo = pw.brk[NSEG] - band;

  constexpr int kOptimistic = 10;  // as in lompc_solve_reg.cuh
  const int n_opt = dmin_hi > 0 ? kOptimistic : 0;
  int it = 0;
  bool done = !live || st != LOMPC_ST_OK;  // group-uniform
  bool converged = false;
  bool have_f = false;      // f holds the objective of an accepted iterate (safeguarded phase)
  bool pending = false;     // W is a rollout that has not been accepted yet
  double f = 0.0, mu = 0.0;
  double sl[SPL];           // local inclusive prefix sums of W
  double off = 0.0;         // state entering this lane's block
  int vh_lane = 0;          // this lane's largest KKT violation (high word) in the last test of a live group

  // states of the iterate: s_k = off + sl[j]
  auto scan_states = [&](double& texc) {
    double run = 0.0, utot = 0.0;
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      run += W[j];
      sl[j] = run;
      utot = fma((double)(N - (k0 + j)), W[j], utot);
    }
    double incl = run, tincl = utot;
#pragma unroll
    for (int d = 1; d < LPQ; d <<= 1) {
      const double t1 = __shfl_up_sync(kFullMask, incl, d, LPQ);
      const double t2 = __shfl_down_sync(kFullMask, tincl, d, LPQ);
      if (li >= d) incl += t1;
      if (li + d < LPQ) tincl += t2;
    }
    off = incl - run;
    texc = tincl - utot;
  };
  // objective of W (without kappa0), from the states scan_states left behind
  auto objective = [&]() -> double {
    double floc = 0.0;
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      const double x = W[j];
      const double e = (off + sl[j]) - gam;
      floc += x * fma(0.5 * D[j], x, G[j]) + 0.5 * c * e * e;
      if (NSEG > 1) {
#pragma unroll
        for (int i = 1; i < NSEG; ++i) floc += (pw.slope[i] - pw.slope[i - 1]) * dpos(x - pw.brk[i]);
      }
    }
    return grp_sum<LPQ>(floc);
  };

  LOMPC_PROF_T(t_setup);
  LOMPC_PROF_ADD(0, t_setup - t_begin);
  while (true) {
    // ================= A: states, acceptance of the pending rollout, KKT test =================
    LOMPC_PROF_T(t_a);
    LOMPC_PROF_ADD(6, 1);
    double texc;
    scan_states(texc);
    const bool safeg = it >= n_opt;  // group-uniform
    if (__any_sync(kFullMask, safeg && !done)) {
      const double fn = objective();
      if (safeg && !done) {
        if (pending && have_f && !(fn <= f + (fbase + 1e-15 * (fabs(f) + fabs(fn))))) {
          // rejected: back to the parked iterate, retry with a larger proximal weight
#pragma unroll
          for (int j = 0; j < SPL; ++j) {
            W[j] = WN[j];
            CD[j] = CDO[j];
          }
          mu = fmax(4.0 * c, 4.0 * mu);
          if (mu > 1e30) done = true;  // gives up (LOMPC_ST_MAXITER)
        } else {
          f = have_f ? dmin2(f, fn) : fn;
          have_f = true;
          mu = 0.0;
        }
      }
      // a rejection changed W in some group: its states are recomputed (the other groups recompute the same values)
      if (__any_sync(kFullMask, safeg && !done && mu > 0.0)) scan_states(texc);
    }
    pending = false;
    if (it >= P.max_iter) done = true;

    // costate, gradient, KKT test; decides binding / the working piece of every stage
    bool binding[SPL];
    double slw[SPL];
    int vh = 0;
    {
      double tloc = texc;
#pragma unroll
      for (int j = SPL - 1; j >= 0; --j) {
        const int k = k0 + j;
        const double wk = W[j];
        const double e = (off + sl[j]) - gam;
        const double p = c * fma((double)(N - k), e, tloc);
        tloc = fma((double)(N - k), wk, tloc);
        const double q = fma(D[j], wk, G[j]) + p;
        const double mq = -q;
        double s_lo, s_hi;
        bool atbp;
        bool top = false, bot = false;
        if (NSEG > 1) {
          const int cd = CD[j];
          const int i = cd >> 1;  // breakpoint i (even code) or piece i (odd code)
          atbp = (cd & 1) == 0;
          const double s_here = pw.slope_at(i & 3);                      // slope[i]     (i < NSEG)
          const double s_left = pw.slope_at((i - 1) & 3);                // slope[i - 1] (i > 0)
          s_hi = dsel(atbp && i == NSEG, 1e300, s_here);
          s_lo = dsel(atbp, dsel(i == 0, -1e300, s_left), s_here);
        } else {
          top = wk >= top_lo;
          bot = wk <= band;
          s_hi = top ? 1e300 : pw.slope[0];
          s_lo = bot ? -1e300 : pw.slope[0];
          atbp = top | bot;
        }
        const double va = mq - s_hi, vb = s_lo - mq;
        const bool right = va > tq, left = vb > tq;
        binding[j] = atbp && !right && !left;
        slw[j] = dsel(left, s_lo, s_hi);
        vh = max(vh, max(__double2hiint(va), __double2hiint(vb)));
        // guess of the piece the rollout will land on: a coordinate pushed off a breakpoint moves into the
        // neighbouring piece, everything else stays where it is
        if (NSEG > 1) {
          CDO[j] = CD[j] + ((atbp && right) ? 1 : 0) - ((atbp && left) ? 1 : 0);
        } else {
          CDO[j] = (atbp && (right || left)) ? 1 : (top ? 2 : (bot ? 0 : 1));
        }
      }
    }
    // converged <=> every lane's violation is below the tolerance: one vote instead of a max-reduction
    // (the maximum itself is only reported, see the epilogue)
    const bool viol_any = grp_any<LPQ>(vh >= tqh, lane);
    if (!done) {
      vh_lane = vh;
      if (!viol_any) {
        converged = true;
        done = true;
      }
    }
    LOMPC_PROF_T(t_b);
    LOMPC_PROF_ADD(1, t_b - t_a);
    if (__all_sync(kFullMask, done)) break;

    // ================= B: Riccati recursion (block operators + suffix scan) -> gains =================
    double KK[SPL];          // x = KK s + KP[piece] inside a piece
    double KP[SPL][NSEG];    // kappa - slope_i / (dm + Q): the stationary point of piece i at s = 0
    {
      RicMat blk;
#pragma unroll
      for (int j = SPL - 1; j >= 0; --j) {
        const double wk = W[j];
        const double dm = D[j] + mu;
        const double gm = fma(-mu, wk, G[j]);
        RicMat M;
        {  // both forms are a handful of flops: computed side by side and selected (no divergent branch)
          const bool bd = binding[j];
          const double h = gm + dsel(bd, 0.0, slw[j]);
          const double sg = pow2_inv_scale(dm + c);
          const double dms = dm * sg;
          M.a11 = dsel(bd, 1.0, dms);
          M.a12 = dsel(bd, c, dms * c);
          M.a21 = dsel(bd, 0.0, sg);
          M.a22 = dsel(bd, 1.0, (dm + c) * sg);
          M.v1 = dsel(bd, wk, -(h * sg));
          M.v2 = dsel(bd, fma(c, wk, -cg), -(fma(dm, cg, c * h) * sg));
          M.m = dsel(bd, 1.0, dms);
        }
        if (j == SPL - 1)
          blk = M;
        else
          blk = ric_mul(M, blk);
      }
      // inclusive suffix scan over the lanes of the group: S_l = B_l B_{l+1} ... B_{LPQ-1}.  Only the second
      // column of S is used (the triple leaving the horizon's end is (0, 1, 0)'), so the last step forms only that.
#pragma unroll
      for (int d = 1; d < LPQ / 2; d <<= 1) {
        const RicMat o = ric_shfl_down(blk, d, LPQ);
        const RicMat pr_ = ric_mul(blk, o);
        const bool act = li + d < LPQ;
        blk.a11 = dsel(act, pr_.a11, blk.a11);
        blk.a12 = dsel(act, pr_.a12, blk.a12);
        blk.a21 = dsel(act, pr_.a21, blk.a21);
        blk.a22 = dsel(act, pr_.a22, blk.a22);
        blk.v1 = dsel(act, pr_.v1, blk.v1);
        blk.v2 = dsel(act, pr_.v2, blk.v2);
        blk.m = dsel(act, pr_.m, blk.m);
      }
      double ca = blk.a12, cb = blk.a22, cr = blk.v2;
      if (LPQ > 1) {
        constexpr int d = LPQ / 2;
        const double oa = __shfl_down_sync(kFullMask, ca, d, LPQ);
        const double ob = __shfl_down_sync(kFullMask, cb, d, LPQ);
        const double orr = __shfl_down_sync(kFullMask, cr, d, LPQ);
        const bool act = li + d < LPQ;
        ca = dsel(act, fma(blk.a11, oa, blk.a12 * ob), ca);
        cb = dsel(act, fma(blk.a21, oa, blk.a22 * ob), cb);
        cr = dsel(act, fma(blk.m, orr, fma(blk.v1, oa, blk.v2 * ob)), cr);
      }
      // triple entering this block from above: second column of S_{l+1} ((0, 1, 0)' for the last block)
      double pa = __shfl_down_sync(kFullMask, ca, 1, LPQ);
      double pb = __shfl_down_sync(kFullMask, cb, 1, LPQ);
      double pr = __shfl_down_sync(kFullMask, cr, 1, LPQ);
      pa = dsel(li == LPQ - 1, 0.0, pa);
      pb = dsel(li == LPQ - 1, 1.0, pb);
      pr = dsel(li == LPQ - 1, 0.0, pr);
#pragma unroll
      for (int j = SPL - 1; j >= 0; --j) {
        const double wk = W[j];
        const double dm = D[j] + mu;
        const double gm = fma(-mu, wk, G[j]);
        const double tq_ = fma(c, pb, pa);
        const double tu = fma(-cg, pb, pr);
        const double bn = fma(dm, pb, tq_);
        const double ib = fast_rcp(bn);
        const double kap = -(fma(gm, pb, tu) * ib);
        const double inv = pb * ib;
        KK[j] = -(tq_ * ib);
#pragma unroll
        for (int i = 0; i < NSEG; ++i) KP[j][i] = NSEG > 1 ? fma(-pw.slope[i], inv, kap) : kap;
        {
          const bool bd = binding[j];
          const double hs = gm + dsel(bd, 0.0, slw[j]);
          pa = dsel(bd, tq_, dm * tq_);
          pr = dsel(bd, fma(tq_, wk, tu), fma(dm, tu, -tq_ * hs));
          pb = dsel(bd, pb, bn);
          if (SPL > 4) {  // long blocks: keep the triple in range (exact power-of-two rescale)
            const double sg = pow2_inv_scale(pb);
            pa *= sg;
            pb *= sg;
            pr *= sg;
          }
        }
      }
    }

    // ================= C: forward rollout (guess the pieces, scan, replay, repeat) =================
    LOMPC_PROF_T(t_c);
    LOMPC_PROF_ADD(2, t_c - t_b);
    {
      double X[SPL];
      int CN[SPL];
#pragma unroll
      for (int j = 0; j < SPL; ++j) CN[j] = CDO[j];  // the guess (CDO is rewritten with the parked codes below)
      for (int pass = 0; pass <= N; ++pass) {
        LOMPC_PROF_ADD(7, 1);
        // Under the guess every stage is an affine map of the state; ps / pt = the state BEFORE stage j as a
        // function of the state entering the block (s_j = ps[j] s_in + pt[j]), am / bm = the whole block.
        double ps[SPL], pt[SPL];
        double am = 1.0, bm = 0.0;
#pragma unroll
        for (int j = 0; j < SPL; ++j) {
          const int cd = CN[j];
          double al, be;  // x = al * s + be
          if (NSEG > 1) {
            const int i = cd >> 1;
            const bool inside = (cd & 1) != 0;
            al = dsel(inside, KK[j], 0.0);
            be = dsel(inside, sel4(i & 3, KP[j][0], KP[j][NSEG > 1 ? 1 : 0], KP[j][NSEG > 1 ? 2 : 0], KP[j][NSEG > 1 ? 3 : 0]),
                      pw.brk_at(i));
          } else {
            al = dsel(cd == 1, KK[j], 0.0);
            be = dsel(cd == 1, KP[j][0], dsel(cd == 2, wmax, 0.0));
          }
          ps[j] = am;
          pt[j] = bm;
          const double a1 = 1.0 + al;
          am *= a1;
          bm = fma(a1, bm, be);
        }
        // inclusive prefix scan of the block maps; the state entering block l is the offset of the composite of blocks < l
#pragma unroll
        for (int d = 1; d < LPQ; d <<= 1) {
          const double oa = __shfl_up_sync(kFullMask, am, d, LPQ);
          const double ob = __shfl_up_sync(kFullMask, bm, d, LPQ);
          const bool act = li >= d;
          bm = dsel(act, fma(am, ob, bm), bm);
          am = dsel(act, am * oa, am);
        }
        double s_in = __shfl_up_sync(kFullMask, bm, 1, LPQ);
        s_in = dsel(li == 0, 0.0, s_in);
        // Replay: the stages of the block are evaluated side by side, each from its state under the guess.  A stage's
        // answer is exact when every stage before it (in this block and the blocks before) landed where it was
        // guessed; the first stage of the horizon that did not is exact too and corrects its guess, so every
        // pass fixes at least one stage and a pass without a miss is the exact rollout.
        bool mis = false;
#pragma unroll
        for (int j = 0; j < SPL; ++j) {
          const double s = fma(ps[j], s_in, pt[j]);
          double x;
          int cdn;
          if (NSEG > 1) {
            double m[NSEG];
            int mc[NSEG];
#pragma unroll
            for (int i = 0; i < NSEG; ++i) {
              const double cj = fma(KK[j], s, KP[j][i]);
              const bool below = cj < pw.brk[i + 1];
              m[i] = below ? cj : pw.brk[i + 1];
              mc[i] = below ? 2 * i + 1 : 2 * i + 2;
            }
#pragma unroll
            for (int h = 1; h < NSEG; h *= 2) {
#pragma unroll
              for (int i = 0; i + h < NSEG; i += 2 * h) {
                const bool first = m[i] > m[i + h];
                m[i] = first ? m[i] : m[i + h];
                mc[i] = first ? mc[i] : mc[i + h];
              }
            }
            const bool ng = __double2hiint(m[0]) < 0;
            x = dpos(m[0]);
            cdn = ng ? 0 : mc[0];
          } else {
            const double x0 = fma(KK[j], s, KP[j][0]);
            const bool over = x0 > wmax, under = x0 < 0.0;
            x = over ? wmax : x0;
            x = under ? 0.0 : x;
            cdn = under ? 0 : (over ? 2 : 1);
          }
          mis |= cdn != CN[j];
          CN[j] = cdn;
          X[j] = x;
        }
        if (!__any_sync(kFullMask, mis && !done)) break;
      }
      if (!done) {
#pragma unroll
        for (int j = 0; j < SPL; ++j) {
          WN[j] = W[j];
          CDO[j] = CD[j];
          W[j] = X[j];
          CD[j] = CN[j];
        }
        pending = true;
        ++it;
      }
    }
    LOMPC_PROF_T(t_e);
    LOMPC_PROF_ADD(3, t_e - t_c);
  }
  LOMPC_PROF_T(t_out);
  if (!converged && st == LOMPC_ST_OK) st = LOMPC_ST_MAXITER;
  st_out = live ? st : LOMPC_ST_OK;

  // ================= outputs (the states of the final iterate are in off / sl) =================
  double closs = 0.0;
  if (live) {
    double* wo = P.w_out + k0;
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      const double x = W[j];
      const double s = off + sl[j];
      wo[j] = x;
      if (P.codes_out) P.codes_out[k0 + j] = (unsigned char)CD[j];
      closs += x * fma(0.5 * D[j], x, G[j]) + 0.5 * c * s * (s - 2.0 * gam);
      if (NSEG > 1) {
#pragma unroll
        for (int i = 1; i < NSEG; ++i) closs += (pw.slope[i] - pw.slope[i - 1]) * dpos(x - pw.brk[i]);
      }
    }
  }
#pragma unroll
  for (int d = LPQ / 2; d >= 1; d >>= 1) {
    const double t1 = __shfl_xor_sync(kFullMask, closs, d, LPQ);
    const int t2 = __shfl_xor_sync(kFullMask, vh_lane, d, LPQ);
    closs += t1;
    vh_lane = max(vh_lane, t2);
  }
  if (live && li == 0) {
    if (P.cost_out) *P.cost_out = cs.theta * wmax * l2sum + closs;
    if (P.status) *P.status = st;
    if (P.iters) *P.iters = it;
    if (P.kkt_res) *P.kkt_res = __hiloint2double(vh_lane, vh_lane ? -1 : 0) / gscale;
  }
#ifdef LOMPC_WARP_PROF
  prof[4] = (unsigned long long)(clock64() - t_out);
  if (lane == 0)
    for (int i = 0; i < 8; ++i) atomicAdd(&g_warp_prof[i], prof[i]);
#endif
}

// The batched entry: QP b of a SolveArgs batch.  Group mode of the price loop (price_solver.py:196-214): QP b uses
// the prices of row group_of[b]; rows whose group has converged are skipped; w_init = the QP's solution at the
// previous prices (warm start).
template <int N, int NSEG, int SPL>
__device__ __forceinline__ void solve_warp(const Consts& cs, const SolveArgs& a, const int64_t b, const bool live_in,
                                           const int lane, int& st_out) {
  const int64_t qp = live_in ? b : 0;
  const int64_t row = (live_in && a.group_of) ? (int64_t)a.group_of[qp] : qp;
  const bool live = live_in && !(a.skip && a.skip[row]);
  WarpProblem P;
  P.lm = a.lmbd + row * a.lmbd_stride;
  P.lr = live ? a.lmbd_r[row * a.lmbd_r_stride] : 0.0;
  P.gam = live ? a.gamma[qp] : 0.0;
  P.w_init = a.w_init ? a.w_init + qp * (int64_t)N : nullptr;
  P.w_out = a.w_out + qp * (int64_t)N;
  P.cost_out = a.cost_out ? a.cost_out + qp : nullptr;
  P.status = a.status ? a.status + qp : nullptr;
  P.iters = a.iters ? a.iters + qp : nullptr;
  P.kkt_res = a.kkt_res ? a.kkt_res + qp : nullptr;
  P.codes_out = nullptr;
  P.tol = a.tol;
  P.max_iter = a.max_iter;
  solve_warp_core<N, NSEG, SPL>(cs, P, live, lane, st_out);
}

// One warp = 32 / LPQ QPs of one segment, ONE warp per CTA: the warp index is blockIdx.x, so the segment, its EV
// type and every loop exit are provably warp-uniform and ptxas emits the shuffles without convergence barriers
// (with several warps per CTA every SHFL of the sweeps was wrapped in WARPSYNC / ENDCOLLECTIVE); 32-thread CTAs
// also let the block scheduler spread a small batch over all SMs.
template <int N, int SPL>
__global__ void __launch_bounds__(32, 1) lompc_solve_warp_kernel(const __grid_constant__ WarpArgs wa) {
  constexpr int LPQ = N / SPL, QPW = 32 / LPQ;
  const int lane = threadIdx.x;
  const int warp = blockIdx.x;
  int sg = 0;
#pragma unroll
  for (int i = 1; i < kMaxWarpSegs; ++i)
    if (i < wa.nsegs && warp >= wa.seg[i].warp_begin) sg = i;
  const WarpSeg& S = wa.seg[sg];
  const int64_t b = (int64_t)(warp - S.warp_begin) * QPW + lane / LPQ;
  const bool live = b < S.a.B;
  int st;
  if (S.cs.large)
    solve_warp<N, 4, SPL>(S.cs, S.a, b, live, lane, st);
  else
    solve_warp<N, 1, SPL>(S.cs, S.a, b, live, lane, st);
  if (wa.summary) {
    // worst status of the launch, tagged with the call's epoch (no memset between calls): every warp with a
    // failure reports it, warp 0 always reports
    const int worst = __reduce_max_sync(kFullMask, st);
    if (lane == 0 && (worst != LOMPC_ST_OK || warp == 0)) {
      const unsigned long long ep = wa.epoch_src ? *wa.epoch_src : 0ull;
      atomicMax(wa.summary, ep * 4ull + (unsigned long long)worst);
    }
  }
}

}  // namespace lompc
