// K5: the whole of PriceSolver.compute_optimal_prices (reference price_solver.py:79-174) for
// one group (station, EV type, partition) in ONE CTA, device-resident from the warm start to
// the regularised prices: no host round trip, no kernel launch and no global-memory traffic
// per iteration.  Groups are independent (they couple only through the BiMPC plan they
// track), so every CTA runs its own group to convergence; the grid is the set of groups.
//
// Per iteration (price_solver.py:111-140):
//   * every thread solves the LoMPC QP of one EV of the group at the current prices (K1 as a
//     device function, iterate in registers); one extra "virtual EV" solves the QP at
//     gamma_sc (price_solver.py:106,132) in the same pass;
//   * thread k < N adds up w_i[k] over the EVs IN EV ORDER (the reference's own summation
//     order, price_solver.py:205 - bit-identical to the phase-split kernels);
//   * warp 0 runs the convergence test, the exact non-negative QP of the price step
//     (price_step_warp) on shared-memory scratch, and the dual-cost bookkeeping.
// The current price row lives in shared memory; EV data (one SoC per EV) is read once.
#pragma once
#include "lompc_common.cuh"
#include "lompc_price.cuh"
#include "lompc_solve_reg.cuh"

namespace lompc {

struct FusedArgs {
  int G;
  int r;
  int tol_type_max;
  double eps_reg, eps_tol;
  int max_iter;            // MAX_PRICE_SOLVER_ITERATIONS
  double qp_tol;           // K1 knobs
  int qp_max_iter;
  const int32_t* group_off;  // [G+1]
  const double* y0;          // [B] SoCs, EVs sorted by group
  const double* w_ref;       // [G,N]
  const double* lmbd_r;      // [G]
  double* prices;            // [G,3N] in: warm start, out: regularised prices
  int32_t* iters;            // [G] value of `iter` at exit, -1 for an empty group
  double* price_pre;         // [G]
  double* price_post;        // [G]
  double* w_k_out;           // [G,N] or NULL
  double* hist_ac;           // [G,hist_cap] or NULL
  double* hist_pred;
  int hist_cap;
  int32_t* flags;            // [0] += groups whose NNQP hit its iteration cap, [1] = some y0 outside [0, y_max],
                             // [2] = max over groups of the LoMPC passes run (loop length of the slowest group),
                             // [3] += LoMPC solves that ended with a status other than OK,
                             // [16] += groups whose pivot pool overflowed (parametric loop, lompc_price_warp.cuh)
  double* w_scratch;         // [B + G, N] LoMPC iterates of groups with more than T - 1 EVs (rows b0 + i; the
                             // virtual EV of group g in row B + g); smaller groups keep them in registers
  int64_t B;
  int chain_S, chain_P;      // price_station_chain_kernel: stations, partitions (G = chain_P * chain_S)
  double* chain_prev;        // [chain_S, 3N] warm start carried along the chain (in/out)
  const int32_t* chain_order;  // [chain_S] station handled by CTA i (a permutation) or NULL
  int compact_step;            // price step with rolled loops (set by the launcher for grids of several waves)
  unsigned long long* qp_count;  // [0] total LoMPC QP solves, [1] / [2] SM cycles summed over groups spent in
                                 // the LoMPC passes / in thread 0's price step, [3] K1 iterations summed over
                                 // the solves, [4] warp passes in which no lane iterated, [5] warp passes (or NULL)
};

template <int N, int NSEG, int T, bool GREG>
struct FusedSmem {
  static constexpr int kK1 = RegSmem<N, NSEG, T, GREG>::kArrays * N * T + RegSmem<N, NSEG, T, GREG>::kTab;
  // LM[3N] WREF[N] WK[N] WSUM[N] + price-step scratch (3*3N + 9N) + ERR[T] + 8 scalars
  static constexpr int kDoubles = kK1 + 3 * N + 3 * N + 18 * N + T + 8;
  static constexpr size_t bytes = (size_t)kDoubles * sizeof(double) + 3 * N + 16;
};

// The loop of ONE group, run by the whole CTA.  `p_in` = warm start (3N doubles), the regularised
// prices go to `p_out` and, if not NULL, to `p_out2`.  Returns false for an empty group (never
// solved, charging_station.py:277,293; nothing is written but iters = -1).
template <int N, int NSEG, int T, bool GREG>
__device__ __forceinline__ bool group_loop_body(const Consts& cs, const FusedArgs& a, const int g, double* smem,
                                                const double* p_in, double* p_out, double* p_out2) {
  const int tid = threadIdx.x;
  const int b0 = a.group_off[g], b1 = a.group_off[g + 1];
  const int n = b1 - b0;
  if (n <= 0) {
    if (tid == 0) a.iters[g] = -1;
    return false;
  }
  constexpr int kK1 = FusedSmem<N, NSEG, T, GREG>::kK1;
  double* TAB = smem + kK1 - RegSmem<N, NSEG, T, GREG>::kTab;  // piece table of K1 (filled by the kernels below)
  double* LM = smem + kK1;          // [3N] current prices
  double* WREF = LM + 3 * N;        // [N]
  double* WK = WREF + N;            // [N] LoMPC solution at gamma_sc for LM
  double* WSUM = WK + N;            // [N]
  double* WS = WSUM + N;            // price-step scratch, 3r + 9N <= 18N doubles
  double* ERR = WS + 18 * N;        // [T]
  double* SC = ERR + T;             // scalars: 0 cost_sc, 1 flag
  unsigned char* WSB = reinterpret_cast<unsigned char*>(SC + 8);  // [r]
  double* WN = smem + 2 * N * T;    // K1's candidate array doubles as the [k][tid] transpose buffer

  // ---- set_charge_levels (price_solver.py:66-77): exact min / max over the group
  double mn = 1e300, mx = -1e300;
  bool bad = false;
  for (int i = tid; i < n; i += T) {
    const double y = a.y0[b0 + i];
    if (!(y >= 0.0 && y <= cs.y_max)) bad = true;
    mn = fmin(mn, y);
    mx = fmax(mx, y);
  }
  ERR[tid] = mn;
  __syncthreads();
  if (tid == 0) {
    double v = ERR[0];
    for (int i = 1; i < T; ++i) v = fmin(v, ERR[i]);
    SC[2] = v;
  }
  __syncthreads();
  ERR[tid] = mx;
  __syncthreads();
  if (tid == 0) {
    double v = ERR[0];
    for (int i = 1; i < T; ++i) v = fmax(v, ERR[i]);
    SC[3] = v;
  }
  if (bad) atomicExch(a.flags + 1, 1);
  for (int k = tid; k < 3 * N; k += T) LM[k] = p_in[k];
  for (int k = tid; k < N; k += T) WREF[k] = a.w_ref[(size_t)g * N + k];
  __syncthreads();
  mn = SC[2];
  mx = SC[3];
  const double y0_rng = (mx - mn) / 2;
  const double gamma_sc = cs.y_max - (mx + mn) / 2;
  const double lr = a.lmbd_r[g];
  const double kappa = lr / cs.delta;
  const double tolg = sqrt((double)N) * y0_rng + a.eps_tol;  // price_solver.py:184
  // gains of A_bar = A'A + kappa I, once per group (they sit where price_step_warp expects them)
  if (tid < 32) abar_factor<N>(N, kappa, WS + 3 * a.r + 6 * N, tid);

  double dual_cost = 0.0, lamdiff = 0.0, dec_pred = 0.0;  // thread 0 only
  int it = 0, nnqp_bad = 0;
  unsigned long long solves = 0, k1_iters = 0, warp_pass0 = 0, warp_pass = 0;
  const bool single = n + 1 <= T;  // one LoMPC pass covers the group: iterates stay in registers
  double W[N];
  long long cyc_qp = 0, cyc_step = 0;  // thread 0: cycles in the LoMPC passes / in the price step
  for (;; ++it) {
    const long long t_a = clock64();
    // ---- LoMPC pass at the current prices: EVs 0..n-1 and the virtual EV n (gamma_sc)
    double wsum = 0.0, emax = 0.0;
    for (int c0 = 0; c0 <= n; c0 += T) {
      const int i = c0 + tid;
      if (i <= n) {
        const double gam = (i < n) ? cs.y_max - a.y0[b0 + i] : gamma_sc;
        double D[N], GR[GREG ? N : 1];
        double l2sum, gscale, viol;
        int st, qit;
        // every QP starts from its own solution at the previous prices (registers, or the
        // scratch rows when the group needs more than one pass of T threads)
        const bool warm = it > 0;
        double* wrow = a.w_scratch + (size_t)(i < n ? (int64_t)b0 + i : a.B + g) * N;
        if (!single && warm) {
#pragma unroll
          for (int k = 0; k < N; ++k) W[k] = wrow[k];
        }
        // (large EV: safeguarded loop only - closed-loop prices have d_k = 0 stages anyway, the fused kernel is
        // instruction-fetch bound and a second copy of the sweeps costs more than it saves: 39.6k vs 33.0k cycles
        // per pass.  Small EV: every stage is strictly convex, the optimistic phase always applies: -11 % per pass)
        solve_reg<N, NSEG, T, GREG, (NSEG == 1)>(cs, LM, lr, gam, a.qp_tol, a.qp_max_iter, warm, false, smem + tid, TAB, W, D, GR, l2sum,
                                    gscale, viol, st, qit);
        if (st != LOMPC_ST_OK) atomicAdd(a.flags + 3, 1);  // a LoMPC solve that did not converge (never observed)
        if (a.qp_count) {
          k1_iters += qit;
          const unsigned act = __activemask();
          const bool none = __all_sync(act, qit == 0);
          if ((tid & 31) == 0) {
            warp_pass0 += none;
            ++warp_pass;
          }
        }
        if (!single) {
#pragma unroll
          for (int k = 0; k < N; ++k) wrow[k] = W[k];
        }
        if (i < n) {
#pragma unroll
          for (int k = 0; k < N; ++k) WN[k * T + tid] = W[k];
          if (a.tol_type_max) {  // price_solver.py:207
            double cum = 0.0, e2 = 0.0;
#pragma unroll
            for (int k = 0; k < N; ++k) {
              const double v = W[k] - WREF[k];
              cum += v;
              e2 += cum * cum + kappa * v * v;
            }
            ERR[tid] = sqrt(e2);
          }
        } else {
          // cost at gamma_sc = the dual cost (lompc.py:155: full objective incl. theta w_max sum(lmbd2))
          const double* GS = smem + tid + 3 * N * T;
          double cost = cs.theta * cs.w_max * l2sum, s = 0.0;
#pragma unroll
          for (int k = 0; k < N; ++k) {
            const double x = W[k];
            WK[k] = x;
            s += x;
            const double gk = GREG ? GR[GREG ? k : 0] : GS[k * T];
            cost += x * fma(0.5 * D[k], x, gk) + 0.5 * cs.c * s * (s - 2.0 * gam);
            if (NSEG > 1) {
#pragma unroll
              for (int j = 1; j < NSEG; ++j) cost += (cs.slope[j] - cs.slope[j - 1]) * dpos(x - cs.brk[j]);
            }
            LOMPC_STAGE_FENCE();
          }
          SC[0] = cost;
        }
      }
      __syncthreads();
      const int cnt = min(T, n - c0);  // real EVs in this chunk (may be <= 0 for the chunk holding only the virtual EV)
      if (tid < N) {
        const double* col = WN + tid * T;
        for (int j = 0; j < cnt; ++j) wsum += col[j];
      } else if (tid == N && a.tol_type_max) {
        for (int j = 0; j < cnt; ++j) emax = fmax(emax, ERR[j]);
        SC[4] = (c0 == 0) ? emax : fmax(SC[4], emax);
      }
      __syncthreads();
    }
    if (tid < N) WSUM[tid] = wsum;
    solves += (unsigned long long)(tid == 0 ? n + 1 : 0);
    __syncthreads();
    // ---- warp 0: bookkeeping of the previous step, convergence test (lane 0), price step (all lanes)
    if (tid < 32) {
      const long long t_b = clock64();
      int flag = 0;
      if (tid == 0) {
        cyc_qp += t_b - t_a;
        const double cost_sc = SC[0];
        if (it > 0 && a.hist_ac && it - 1 < a.hist_cap) {  // price_solver.py:135-139
          a.hist_ac[(size_t)g * a.hist_cap + it - 1] = cost_sc - dual_cost + lamdiff;
          a.hist_pred[(size_t)g * a.hist_cap + it - 1] = dec_pred;
        }
        dual_cost = cost_sc;
        if (it >= a.max_iter) {
          flag = 2;  // the loop ran out: `iter` ends at max_iter - 1 (price_solver.py:111)
        } else {
          double w_avg_err, w0_err;
          price_errors(N, kappa, WSUM, (double)n, WREF, w_avg_err, w0_err);
          const double w_err = a.tol_type_max ? SC[4] : w_avg_err;
          if (w_err <= tolg) flag = 1;  // price_solver.py:125
        }
        SC[1] = (double)flag;
      }
      flag = __shfl_sync(0xffffffffu, flag, 0);
      if (flag == 0) {
        int st;
        // compact (rolled) price step when the grid is several waves deep: there the kernel is instruction-fetch
        // bound and the step's unrolled code is a third of what an MM iteration streams through the cache
        if (a.compact_step)
          price_step_warp<0>(cs, a.r, kappa, a.eps_reg, LM, WK, WREF, WS, WSB, tid, it == 0, it > 0,
                             a.hist_ac != nullptr, true, lamdiff, dec_pred, st);
        else
          price_step_warp<N>(cs, a.r, kappa, a.eps_reg, LM, WK, WREF, WS, WSB, tid, it == 0, it > 0,
                             a.hist_ac != nullptr, true, lamdiff, dec_pred, st);
        nnqp_bad |= st;
      }
      if (tid == 0) cyc_step += clock64() - t_b;
    }
    __syncthreads();
    if (SC[1] != 0.0) break;
  }
  if (a.qp_count) {
    if (k1_iters) atomicAdd(a.qp_count + 3, k1_iters);
    if (warp_pass) {
      atomicAdd(a.qp_count + 4, warp_pass0);
      atomicAdd(a.qp_count + 5, warp_pass);
    }
  }
  // ---- regularisation (price_solver.py:145-147) and outputs
  if (tid == 0) {
    a.iters[g] = (SC[1] == 2.0) ? a.max_iter - 1 : it;
    double pre, post;
    regularize_core(cs, a.r, WK, LM, pre, post);
    a.price_pre[g] = pre;
    a.price_post[g] = post;
    if (nnqp_bad & 1) atomicAdd(a.flags, 1);
    if (nnqp_bad & 2) atomicAdd(a.flags + 17, 1);  // groups that took the NNQP fallback (informational)
    atomicMax(a.flags + 2, it);
    if (a.qp_count) {
      atomicAdd(a.qp_count, solves);
      atomicAdd(a.qp_count + 1, (unsigned long long)cyc_qp);
      atomicAdd(a.qp_count + 2, (unsigned long long)cyc_step);
    }
  }
  __syncthreads();
  for (int k = tid; k < 3 * N; k += T) {
    p_out[k] = LM[k];
    if (p_out2) p_out2[k] = LM[k];
  }
  if (a.w_k_out)
    for (int k = tid; k < N; k += T) a.w_k_out[(size_t)g * N + k] = WK[k];
  return true;
}

// Grid = the groups: every CTA runs its own group.
template <int N, int NSEG, int T, int MINB, bool GREG>
__global__ void __launch_bounds__(T, MINB) price_group_loop_kernel(const Consts cs, const FusedArgs a) {
  extern __shared__ double smem[];
  piece_table<NSEG>(cs, smem + FusedSmem<N, NSEG, T, GREG>::kK1 - RegSmem<N, NSEG, T, GREG>::kTab, threadIdx.x);
  __syncthreads();
  const int g = blockIdx.x;
  double* row = a.prices + (size_t)g * 3 * N;
  group_loop_body<N, NSEG, T, GREG>(cs, a, g, smem, row, row, nullptr);
}

// Grid = the stations: every CTA runs the P partitions of ITS station one after the other, each
// warm-started from the prices of the last non-empty partition before it - the warm-start chain of
// the reference, whose PriceSolver object (and its prev_prices) is shared by the partitions of an EV
// type (price_solver.py:56,104,166; charging_station.py:273-304).  Groups are partition-major
// (g = p S + s); chain_prev[S,3N] is the carried warm start (in/out), a.prices[G,3N] receives every
// group's prices (zeros for an empty group, charging_station.py:270).  Stations never wait for each
// other: the step ends when the slowest STATION is done, not after the sum of the slowest groups.
template <int N, int NSEG, int T, int MINB, bool GREG>
__global__ void __launch_bounds__(T, MINB) price_station_chain_kernel(const Consts cs, const FusedArgs a) {
  extern __shared__ double smem[];
  piece_table<NSEG>(cs, smem + FusedSmem<N, NSEG, T, GREG>::kK1 - RegSmem<N, NSEG, T, GREG>::kTab, threadIdx.x);
  __syncthreads();
  // launch order: longest expected chains first (the caller sorts stations by last step's iteration
  // totals), so that a station stuck at the iteration cap does not start in the last wave
  const int S = a.chain_S, s = a.chain_order ? a.chain_order[blockIdx.x] : (int)blockIdx.x;
  double* prev = a.chain_prev + (size_t)s * 3 * N;
  for (int p = 0; p < a.chain_P; ++p) {
    const int g = p * S + s;
    double* row = a.prices + (size_t)g * 3 * N;
    const bool solved = group_loop_body<N, NSEG, T, GREG>(cs, a, g, smem, prev, prev, row);
    if (!solved)
      for (int k = threadIdx.x; k < 3 * N; k += T) row[k] = 0.0;
    __syncthreads();  // prev (global) is re-read by the next group; shared memory is reused
  }
}

}  // namespace lompc
