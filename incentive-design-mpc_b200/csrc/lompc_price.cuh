// K2-K5: the price loop around the LoMPC solve (reference price_solver.py,
// price_regularizer.py), batched over G independent (station, EV-type, partition)
// groups.  EVs are sorted by group: group g owns EVs [group_off[g], group_off[g+1]).
//
//   group_of_kernel      EV -> group index (binary search in group_off)
//   group_stats_kernel   PriceSolver.set_charge_levels        price_solver.py:66-77
//   colsum_kernel        w_avg / w_err_max of _get_w_err      price_solver.py:199-213
//   group_step_kernel    convergence test + _price_gradient_descent_step
//                                                             price_solver.py:121-131,216-246
//   bookkeep_kernel      dual-cost statistics                 price_solver.py:135-140
//   regularize_kernel    _regularize_prices (closed form of the LP, price_regularizer.py:68-85)
//   mean_kernel          mean first-step price of get_w0_price0   price_solver.py:281-284
//
// All group-level reductions run in a fixed order (the reference's own order:
// EV 0, 1, ... of the group), so results do not depend on the launch geometry.
#pragma once
#include "lompc_common.cuh"

namespace lompc {

// Unroll factor of the price step's recursions at a compile-time horizon NT: UF stages per loop trip (kFullUnroll: all
// of them).  Fully unrolled code is fastest for a group on its own; at fleet scale (8 warps per SM at unrelated points
// of a 290 KB kernel, instruction fetch the top stall) 4 stages per trip are: price loops of a 4,096-station step
// 46.8 ms fully unrolled, 41.1 ms by 8, 40.0 ms by 4 (round 2) - the arithmetic is the same in every variant.
constexpr int kFullUnroll = 1024;
template <int NT, int UF>
struct K3Unroll {
  static constexpr int value = NT ? (UF < NT ? UF : NT) : 1;
};


struct PriceArgs {
  int G;
  int r;               // 2N ("linear") or 3N ("linear-convex"), price_solver.py:44-47
  int tol_type_max;    // settings.PRICE_SOLVER_TOL_TYPE == "max"
  double eps_reg;      // settings.PRICE_SOLVER_EPS_REG
  double eps_tol;      // settings.PRICE_SOLVER_EPS_TOL
  const int32_t* group_off;  // [G+1]
  const double* w_ref;       // [G,N]
  const double* lmbd_r;      // [G]
  const double* y0_rng;      // [G]
  double* lmbd;              // [G,3N] current prices (updated in place)
  double* w_k;               // [G,N]  LoMPC solution at gamma_sc for the current prices
  double* w_avg;             // [G,N] mean of w over the group's EVs, or the SUM when cnt != NULL
  const double* cnt;         // [G] number of EVs per group over all ranks (NULL: w_avg already is the mean)
  double* w_err_max;         // [G]
  double* w_avg_err;         // [G]
  double* w0_err;            // [G]
  double* dual_cost;         // [G]
  double* cost_new;          // [G]
  double* lamdiff_phi;       // [G] (lmbd_k - lmbd_k_new) @ phi(w_ref), first iteration only (aliasing quirk :140)
  double* dec_pred;          // [G]
  int32_t* skip;             // [G] 1 = converged / empty
  int32_t* iters;            // [G]
  int32_t* nnqp_status;      // [G] 0 ok, bit 0 = the price step's NNQP stopped at its iteration cap, bit 1 = it took the fallback
  int32_t* n_active;         // [1]
  double* hist_ac;           // [G,hist_cap] or NULL
  double* hist_pred;         // [G,hist_cap] or NULL
  int hist_cap;
  unsigned char* wsb;        // scratch, r * G bytes (keeps the free set between iterations)
  int cold;                  // 1: never warm-start the free set (stand-alone price_step_dev)
  int want_dec;              // 1: compute dec_pred even without a history buffer
  // ---- sharded loop without host round trips / with the aggregate exchanged through peer memory (all optional)
  volatile int32_t* publish_ring;  // pinned host ring of (tag, n_active) pairs, slot it % publish_slots; NULL: none
  int publish_slots;
  int peer_world;            // > 1: w_avg is gathered from the ranks' peer regions inside group_step_kernel
  int peer_rank;
  char* peer_region[8];
  unsigned long long peer_tag0;  // flag value of iteration 0, minus one
  int32_t* peer_timeout;     // set to 1 when a rank did not deliver in time
  unsigned int* blocks_done; // colsum_kernel's "last block signals" counter
  // ---- column sums formed inside group_step_kernel (one rank holds every EV: nothing to reduce between the EV
  // phase and the group phase, price_shard_local_sums); NULL: w_avg / w_err_max were filled by colsum_kernel
  const double* cs_w_ev;     // [B,N] the EVs' solutions
  const double* cs_err_ev;   // [B] their error norms (tol type "max") or NULL
  const int32_t* cs_off;     // [G+1]
};

constexpr size_t kPeerFlagBytes = 1024;  // head of a peer region: one 64-bit flag per rank
constexpr int kMaxPeers = 8;             // == the size of PriceArgs::peer_region
struct PeerView {                        // the peer regions attached to a handle (price_shard_attach_peers)
  int rank, world;
  char* region[kMaxPeers];
};
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

__global__ void group_of_kernel(int64_t B, int G, const int32_t* __restrict__ off, int32_t* __restrict__ group_of) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int lo = 0, hi = G;  // find g with off[g] <= b < off[g+1]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if ((int64_t)off[mid] <= b) lo = mid; else hi = mid;
  }
  group_of[b] = lo;
}

// price_solver.py:66-77 for every group; also gamma_i = y_max - y0_i (price_solver.py:201,279).
__global__ void group_stats_kernel(const Consts cs, int G, const int32_t* __restrict__ off,
                                   const double* __restrict__ y0, double* __restrict__ gamma,
                                   double* __restrict__ y0_rng, double* __restrict__ gamma_sc,
                                   double* __restrict__ gamma_sm, int32_t* __restrict__ bad) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  const int b0 = off[g], b1 = off[g + 1];
  if (b1 <= b0) {
    y0_rng[g] = 0.0; gamma_sc[g] = 0.0; gamma_sm[g] = 0.0;
    return;
  }
  double mn = y0[b0], mx = y0[b0], sum = 0.0;
  for (int b = b0; b < b1; ++b) {
    const double y = y0[b];
    if (!(y >= 0.0 && y <= cs.y_max)) atomicExch(bad, 1);  // assert of price_solver.py:71
    mn = fmin(mn, y);
    mx = fmax(mx, y);
    sum += y;
    gamma[b] = cs.y_max - y;
  }
  y0_rng[g] = (mx - mn) / 2;
  gamma_sc[g] = cs.y_max - (mx + mn) / 2;
  gamma_sm[g] = cs.y_max - sum / (b1 - b0);
}

// Sharded variant of set_charge_levels: per-group LOCAL min / max / sum / count of y0 (the
// caller all-reduces them with MIN / MAX / SUM / SUM), then stats_finalize_kernel.
__global__ void group_stats_local_kernel(const Consts cs, int G, const int32_t* __restrict__ off,
                                         const double* __restrict__ y0, double* __restrict__ gamma,
                                         double* __restrict__ smin, double* __restrict__ smax,
                                         double* __restrict__ ssum, double* __restrict__ scnt,
                                         int32_t* __restrict__ bad) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  const int b0 = off[g], b1 = off[g + 1];
  double mn = 1e300, mx = -1e300, sum = 0.0;
  for (int b = b0; b < b1; ++b) {
    const double y = y0[b];
    // y0 outside [0, y_max] (price_solver.py:71; NaN included) is reported THROUGH the statistics: the minimum
    // is forced below zero, so that after the caller's MIN all-reduce every rank sees it (stats_finalize_kernel)
    if (!(y >= 0.0 && y <= cs.y_max)) {
      atomicExch(bad, 1);
      mn = -1e300;
    }
    mn = fmin(mn, y);
    mx = fmax(mx, y);
    sum += y;
    gamma[b] = cs.y_max - y;
  }
  smin[g] = mn; smax[g] = mx; ssum[g] = sum; scnt[g] = (double)(b1 - b0);
}

__global__ void stats_finalize_kernel(const Consts cs, int G, int max_iter, const double* __restrict__ smin,
                                      const double* __restrict__ smax, const double* __restrict__ ssum,
                                      const double* __restrict__ scnt, double* __restrict__ y0_rng,
                                      double* __restrict__ gamma_sc, double* __restrict__ gamma_sm,
                                      int32_t* __restrict__ skip, int32_t* __restrict__ iters,
                                      int32_t* __restrict__ nnqp_status, int32_t* __restrict__ empty_out,
                                      int32_t* __restrict__ bad) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  const bool empty = !(scnt[g] > 0.0);
  // the assert of price_solver.py:71 on the REDUCED statistics: every rank raises, not only the owner of the EV
  if (!empty && !(smin[g] >= 0.0 && smax[g] <= cs.y_max)) atomicExch(bad, 1);
  empty_out[g] = empty;
  y0_rng[g] = empty ? 0.0 : (smax[g] - smin[g]) / 2;
  gamma_sc[g] = empty ? 0.0 : cs.y_max - (smax[g] + smin[g]) / 2;
  gamma_sm[g] = empty ? 0.0 : cs.y_max - ssum[g] / scnt[g];
  skip[g] = empty;
  iters[g] = empty ? -1 : max_iter - 1;  // -1: the reference logs niter = -1 for an empty partition (charging_station.py:404-411)
  nnqp_status[g] = 0;
}

// One thread per (group, time step): w_avg[g,k] = mean_i w_i[k] in EV order; thread k = 0
// also takes max_i err_i (price_solver.py:199-210).
// ---- the aggregate exchange of the sharded price loop over NVLink PEER MEMORY (instead of an NCCL all-reduce) ----
// Every rank owns a region [flags (kPeerFlagBytes) | partial sums, buffer 0 | buffer 1] that all ranks have mapped
// (CUDA IPC).  Iteration `it`: colsum_signal_kernel writes this rank's [G, N] partial sums into buffer it & 1 of its
// OWN region and its last block raises flag[rank] = tag in EVERY rank's region; group_step_kernel (first kernel of the
// group phase) waits, per group, until all `world` flags of its own region carry the tag and adds the partial sums of
// ranks 0, 1, ... in that fixed order straight out of the peers' memory - every rank gets the same bits, no
// collective library call and no extra launch.  Double buffering + stream order make the flags enough: a rank
// rewrites buffer b two iterations later, after the peers' flags of the iteration in between, which they raise only
// after they have finished reading b.
//
// Raises this rank's flag in every rank's peer region: called by ONE thread block after the partial sums are complete.
__device__ __forceinline__ void peer_raise_flags(int world, int rank, char* const* region, unsigned long long tag) {
  const int r = threadIdx.x;
  if (r < world) {
    unsigned long long* flag = reinterpret_cast<unsigned long long*>(region[r]) + rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(tag) : "memory");
  }
}

template <bool SIGNAL>
__device__ __forceinline__ void colsum_body(int N, int G, const int32_t* __restrict__ off, const int32_t* __restrict__ skip,
                                            const double* __restrict__ w_ev, const double* __restrict__ err_ev,
                                            double* __restrict__ w_avg, double* __restrict__ w_err_max, int sum_only);

__global__ void colsum_kernel(int N, int G, const int32_t* __restrict__ off, const int32_t* __restrict__ skip,
                              const double* __restrict__ w_ev, const double* __restrict__ err_ev,
                              double* __restrict__ w_avg, double* __restrict__ w_err_max, int sum_only) {
  colsum_body<false>(N, G, off, skip, w_ev, err_ev, w_avg, w_err_max, sum_only);
}

// colsum_kernel + the flag of the peer exchange: the block that finishes LAST (a counter in global memory) raises
// this rank's flag on every rank - one launch less per iteration than a separate signalling kernel.
__global__ void colsum_signal_kernel(int N, int G, const int32_t* __restrict__ off, const int32_t* __restrict__ skip,
                                     const double* __restrict__ w_ev, double* __restrict__ sums, const PriceArgs p,
                                     unsigned long long tag) {
  colsum_body<false>(N, G, off, skip, w_ev, nullptr, sums, nullptr, 1);
  __shared__ unsigned int last;
  __threadfence();  // this block's sums before its tick of the counter
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(p.blocks_done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last) {
    if (threadIdx.x == 0) *p.blocks_done = 0u;
    __threadfence_system();
    peer_raise_flags(p.peer_world, p.peer_rank, p.peer_region, tag);
  }
}

template <bool SIGNAL>
__device__ __forceinline__ void colsum_body(int N, int G, const int32_t* __restrict__ off, const int32_t* __restrict__ skip,
                                            const double* __restrict__ w_ev, const double* __restrict__ err_ev,
                                            double* __restrict__ w_avg, double* __restrict__ w_err_max, int sum_only) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)G * N) return;
  const int g = (int)(idx / N), k = (int)(idx % N);
  if (skip && skip[g]) {
    // a converged group's row is not read again, but a multi-GPU caller keeps all-reducing the whole buffer:
    // a stale partial sum would be re-summed over the ranks every iteration and grow without bound
    if (sum_only) w_avg[idx] = 0.0;
    return;
  }
  const int b0 = off[g], b1 = off[g + 1];
  // The additions stay in EV order (the reference's own, price_solver.py:205); the LOADS do not depend on each other:
  // eight are issued before the first addition, so a group of 32 EVs costs 4 trips to L2 instead of 32.
  double sum = 0.0;
  const double* col = w_ev + k;
  int b = b0;
  for (; b + 8 <= b1; b += 8) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = col[(int64_t)(b + u) * N];
#pragma unroll
    for (int u = 0; u < 8; ++u) sum += v[u];
  }
  for (; b < b1; ++b) sum += col[(int64_t)b * N];
  w_avg[idx] = sum_only ? sum : sum / (b1 - b0);  // sum_only: the caller all-reduces, then divides by the global count
  if (k == 0 && err_ev) {
    double m = 0.0;
    for (int b = b0; b < b1; ++b) m = fmax(m, err_ev[b]);
    w_err_max[g] = m;
  }
}

// x = (diag(d) + c A'A)^{-1} b by the scalar Riccati recursion (A = tril(ones)), by ONE WARP on
// contiguous vectors (every lane must call).  d_k = dvec[k] (+ dadd).  K/KAP are scratch.
// Homogeneous form of the recursion (P = pa/pb, r = pr/pb): lane 0 runs the dependent chain (two
// FMAs per stage) and only STORES the numerators / denominators; the N reciprocals and gains are
// then formed by all lanes at once (off the chain), and lane 0 runs the forward substitution.
// NT = compile-time horizon (fully unrolled: the loads of a stage are issued ahead of the chain) or
// 0 for a run-time N (rolled, compact code: the fused price loop is instruction-fetch bound at fleet scale).
template <int NT, int UF = kFullUnroll>
__device__ __forceinline__ void ric_solve(int Nrt, const double* dvec, double dadd, double c, const double* bvec,
                                          double* x, double* K, double* KAP, int lane) {
  const int N = NT ? NT : Nrt;
  if (lane == 0) {
    double pa = 0.0, pb = 1.0, pr = 0.0;
#pragma unroll(K3Unroll<NT, UF>::value)
    for (int k = N - 1; k >= 0; --k) {
      const double d = (dvec ? dvec[k] : 0.0) + dadd;
      const double gk = -bvec[k];
      const double tq = fma(c, pb, pa);   // Q pb,  Q = c + P
      const double bn = fma(d, pb, tq);   // (d + Q) pb
      K[k] = tq;                          // numerator of Q / (d + Q)
      KAP[k] = fma(gk, pb, pr);           // numerator of (r + g_k) / (d + Q)
      x[k] = bn;                          // common denominator
      pa = d * tq;                        // P <- Q d / (d + Q)
      pr = fma(d, pr, -tq * gk);          // r <- (d r - Q g_k) / (d + Q)
      pb = bn;
      if ((k & 7) == 0 && pb > 0x1p600) {  // keep the homogeneous triple in range (exact rescale)
        pa *= 0x1p-600;
        pb *= 0x1p-600;
        pr *= 0x1p-600;
      }
    }
  }
  __syncwarp();
  for (int k = lane; k < N; k += 32) {
    const double ib = fast_rcp(x[k]);
    K[k] *= ib;
    KAP[k] *= ib;
  }
  __syncwarp();
  if (lane == 0) {
    double s = 0.0;
#pragma unroll(K3Unroll<NT, UF>::value)
    for (int k = 0; k < N; ++k) {  // s_{k+1} = s_k + x_k = (1 - K_k) s_k - kappa_k: ONE dependent FMA per stage
      const double kk = K[k], kap = KAP[k];
      x[k] = -fma(kk, s, kap);
      s = fma(1.0 - kk, s, -kap);
    }
  }
}

// A_bar = A'A + kappa I is the same for every solve of a group: its Riccati gains are computed ONCE
// (fac[0:N] = K_k = Q_k/(kappa + Q_k), fac[N:2N] = 1 - K_k, fac[2N:3N] = 1/(kappa + Q_k), Q_k = 1 + kappa K_{k+1})
// and a solve is then two chains of one dependent FMA per stage, without reciprocals.
template <int NT>
__device__ __forceinline__ void abar_factor(int Nrt, double kappa, double* fac, int lane) {
  const int N = NT ? NT : Nrt;
  if (lane == 0) {
    double P = 0.0;
    for (int k = N - 1; k >= 0; --k) {
      const double Q = 1.0 + P;
      const double ib = 1.0 / (kappa + Q);
      const double K = Q * ib;
      fac[k] = K;
      fac[N + k] = kappa * ib;
      fac[2 * N + k] = ib;
      P = kappa * K;
    }
  }
  __syncwarp();
}

// x = A_bar^{-1} b with the cached gains (every lane must call; lane 0 runs the two chains, KAP is scratch).
template <int NT, int UF = kFullUnroll>
__device__ __forceinline__ void abar_solve(int Nrt, const double* fac, const double* bvec, double* x, double* KAP,
                                           int lane) {
  const int N = NT ? NT : Nrt;
  if (lane == 0) {
    double r = 0.0;
#pragma unroll(K3Unroll<NT, UF>::value)
    for (int k = N - 1; k >= 0; --k) {  // kappa_k = (r_{k+1} - b_k)/(kappa + Q_k);  r_k = (1 - K_k) r_{k+1} + K_k b_k
      const double bk = bvec[k];
      KAP[k] = (r - bk) * fac[2 * N + k];
      r = fma(fac[N + k], r, fac[k] * bk);
    }
    double s = 0.0;
#pragma unroll(K3Unroll<NT, UF>::value)
    for (int k = 0; k < N; ++k) {
      const double kap = KAP[k];
      x[k] = -fma(fac[k], s, kap);
      s = fma(fac[N + k], s, -kap);
    }
  }
}

// Scratch of one price step: (3r + 9N) doubles (the last 3N: the cached gains of A_bar).
__host__ __device__ inline int price_step_scratch_doubles(int N, int r) { return 3 * r + 9 * N; }

// Test hook: non-zero makes every price step take the fallback below (price_debug_force_nnqp_fallback()).
__device__ int g_nnqp_force_fallback = 0;

// Fallback of the price step's NNQP: Lawson-Hanson's primal active-set iteration (one index enters per outer
// iteration - the most negative half gradient -, a ratio test along the way to the free-set solution drops the
// blocking ones), started from l = 0.  It cannot cycle the way the primal-dual iteration of price_step_warp
// does on a few degenerate groups (about 1 price step in 10^6 at fleet scale), at the price of one Riccati
// solve per index instead of a handful per step; the oracle's NNLS (oracle/price_oracle.py::nnqp_exact) is the
// same algorithm on the dense factor.  Same conventions as price_step_warp (every lane calls; LAM, TERM and
// FREE are left as its converged branch leaves them).  Returns 0, or 1 when its own cap of 3r outer iterations
// (Lawson-Hanson's customary bound) is hit.
__device__ __noinline__ int nnqp_lawson_hanson(int N, int nb, double th, double m, double eps, double kappa, double gs,
                                               const double* C3, const double* RHO, const double* FAC, double* LAM,
                                               double* TERM, double* U, double* V, double* TD, double* KS,
                                               double* KAPS, unsigned char* FREE, int lane) {
  const unsigned full = 0xffffffffu;
  const int r = nb * N;
  const double inv2m = 1.0 / (2.0 * m), inv_eps = 1.0 / eps;
  for (int i = lane; i < r; i += 32) {
    LAM[i] = 0.0;
    FREE[i] = 0;
  }
  __syncwarp();
  int st = 1;
  for (int oit = 0; oit <= 3 * r; ++oit) {
    // half gradient P l - rho at the current (feasible) l, through v = A_bar^{-1} B'l
    for (int k = lane; k < N; k += 32) {
      const double coef[3] = {th, -th, C3[k]};
      double u = 0.0;
      for (int j = 0; j < nb; ++j) u += coef[j] * LAM[j * N + k];
      U[k] = u;
    }
    __syncwarp();
    abar_solve<0>(N, FAC, U, V, KAPS, lane);
    __syncwarp();
    double best = -1e-12 * gs;
    int bi = 0x7fffffff;
    for (int k = lane; k < N; k += 32) {
      const double coef[3] = {th, -th, C3[k]};
      const double v = V[k] * inv2m;
      for (int j = 0; j < nb; ++j) {
        const int i = j * N + k;
        const double l = LAM[i];
        const double Pl = eps * l + coef[j] * v;
        TERM[i] = l * (Pl - 2.0 * RHO[i]);
        const double hg = Pl - RHO[i];
        if (FREE[i] == 0 && hg < best) {  // (FREE = 2: an index that entered and left again without a step)
          best = hg;
          bi = i;
        }
      }
    }
    for (int o = 16; o > 0; o >>= 1) {  // most negative half gradient, lowest index on ties
      const double ob = __shfl_xor_sync(full, best, o);
      const int oi = __shfl_xor_sync(full, bi, o);
      if (ob < best || (ob == best && oi < bi)) {
        best = ob;
        bi = oi;
      }
    }
    if (bi == 0x7fffffff) {
      st = 0;
      break;
    }
    if (oit == 3 * r) break;
    if (lane == 0) FREE[bi] = 1;
    __syncwarp();
    for (int iit = 0; iit <= 3 * r; ++iit) {
      // s = the minimiser over the free set (Woodbury, as in price_step_warp); kept in TERM
      for (int k = lane; k < N; k += 32) {
        const double coef[3] = {th, -th, C3[k]};
        double t = 0.0, rhs = 0.0;
        for (int j = 0; j < nb; ++j)
          if (FREE[j * N + k] == 1) {
            t += coef[j] * coef[j];
            rhs += coef[j] * RHO[j * N + k];
          }
        TD[k] = t;
        U[k] = rhs;
      }
      __syncwarp();
      ric_solve<0>(N, TD, 2.0 * m * eps * kappa, 2.0 * m * eps, U, V, KS, KAPS, lane);
      __syncwarp();
      double alpha = 2.0;
      for (int k = lane; k < N; k += 32) {
        const double coef[3] = {th, -th, C3[k]};
        for (int j = 0; j < nb; ++j) {
          const int i = j * N + k;
          if (FREE[i] != 1) continue;
          const double sv = (RHO[i] - coef[j] * V[k]) * inv_eps;
          TERM[i] = sv;
          if (sv <= 0.0) alpha = fmin(alpha, LAM[i] / (LAM[i] - sv));  // l_i >= 0 >= s_i, not both 0 unless i just entered
        }
      }
      for (int o = 16; o > 0; o >>= 1) alpha = fmin(alpha, __shfl_xor_sync(full, alpha, o));
      if (!(alpha >= 0.0)) alpha = 0.0;  // 0/0 of an index that entered with a non-positive target
      __syncwarp();
      for (int k = lane; k < N; k += 32)
        for (int j = 0; j < nb; ++j) {
          const int i = j * N + k;
          if (FREE[i] != 1) continue;
          const double sv = TERM[i], l = LAM[i];
          if (alpha > 1.0) {
            LAM[i] = sv;
          } else if (sv <= 0.0 && !(l / (l - sv) > alpha)) {  // the blocking indices leave at exactly 0
            LAM[i] = 0.0;
            FREE[i] = (i == bi && iit == 0) ? 2 : 0;
          } else {
            LAM[i] = fmax(l + alpha * (sv - l), 0.0);
          }
        }
      __syncwarp();
      if (alpha > 1.0) break;
    }
    if (FREE[bi] != 2)  // a step was taken: indices set aside (round-off made them enter with a non-positive target)
      for (int i = lane; i < r; i += 32)  // may be tried again
        if (FREE[i] == 2) FREE[i] = 0;
    __syncwarp();
  }
  __syncwarp();
  for (int i = lane; i < r; i += 32) FREE[i] = (FREE[i] == 1);
  __syncwarp();
  return st;
}

// The exact solution of the non-negative QP of the price step
//     min_{l >= 0} l'P l + q'l,  P = Dphi A_bar^{-1} Dphi'/(2m) + eps I,  q = -2 P l_k - (phi(w_k) - phi(w_ref))
// (price_solver.py:216-246) by a primal-dual active-set iteration, for ONE group by ONE WARP
// (every lane of the warp must call; lanes stride over the horizon, lane 0 runs the O(N)
// Riccati recursions).  P is never formed: with B = Dphi[:r] (three diagonals) the free-set
// system is solved through Woodbury,
//     l_F = (rho_F - B_F z)/eps,   (2 m eps A_bar + B_F'B_F) z = B_F' rho_F,   rho = -q/2,
// and A_bar = A'A + kappa I makes that an O(N) Riccati solve like the LoMPC's own.
// lk[3N] is updated in place; ws = price_step_scratch_doubles() doubles and FREE = r bytes of
// (shared-memory) scratch private to the warp.  FREE carries the free set from one price
// iteration to the next (`warm`): the result depends only on the FINAL free set, and that
// rarely changes between iterations, so a warm start costs one verification pass.  Sums that
// feed statistics are taken by lane 0 in (k, j) order, so the result does not depend on the
// number of lanes.
template <int NT, int UF = kFullUnroll>
__device__ __forceinline__ void price_step_warp(const Consts& cs, int r, double kappa, double eps, double* lk,
                                                const double* wk, const double* wr, double* ws,
                                                unsigned char* FREE, int lane, bool first, bool warm,
                                                bool need_dec, bool have_fac, double& lamdiff_out,
                                                double& dec_pred_out, int& status_out) {
  const int N = NT ? NT : cs.N;
  double* LAM = ws;        // [r] new prices
  double* RHO = LAM + r;   // [r]
  double* TERM = RHO + r;  // [r] summands of the ordered sums
  double* C3 = TERM + r;   // [N] third diagonal of Dphi': 2 q w_k
  double* KS = C3 + N;     // [N]
  double* KAPS = KS + N;   // [N]
  double* U = KAPS + N;    // [N]
  double* V = U + N;       // [N]
  double* TD = V + N;      // [N]
  double* FAC = TD + N;    // [3N] gains of A_bar (abar_factor): computed here unless the caller keeps them
  const int nb = r / N;    // 2 or 3 price blocks
  const double th = cs.theta, qs = cs.q_scale, m = cs.c;
  // FP64 division is a ~30-instruction sequence: divide once, multiply everywhere
  const double inv2m = 1.0 / (2.0 * m), inv_eps = 1.0 / eps;
  const unsigned full = 0xffffffffu;
  auto ordered_sum = [&](void) {  // lane 0: sum of TERM in (k, j) order
    double acc = 0.0;
    if (lane == 0)
      for (int k = 0; k < N; ++k)
        for (int j = 0; j < nb; ++j) acc += TERM[j * N + k];
    return acc;
  };
  // u = B' l_k ; v = A_bar^{-1} u ; rho = P l_k + (phi(w_k) - phi(w_ref))/2
  for (int k = lane; k < N; k += 32) {
    const double c3 = 2.0 * qs * wk[k];
    C3[k] = c3;
    U[k] = th * (lk[k] - lk[N + k]) + (nb == 3 ? c3 * lk[2 * N + k] : 0.0);
  }
  if (!have_fac) abar_factor<NT>(N, kappa, FAC, lane);
  __syncwarp();
  abar_solve<NT, UF>(N, FAC, U, V, KAPS, lane);
  __syncwarp();
  double gs = 1.0;
  for (int k = lane; k < N; k += 32) {
    const double v = V[k] * inv2m;
    const double dw = wk[k] - wr[k];
    const double dphi[3] = {th * dw, -th * dw, qs * (wk[k] * wk[k] - wr[k] * wr[k])};
    const double coef[3] = {th, -th, C3[k]};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (j >= nb) break;
      const double l = lk[j * N + k];
      const double Pl = eps * l + coef[j] * v;
      const double rho = Pl + 0.5 * dphi[j];
      RHO[j * N + k] = rho;
      gs = fmax(gs, fabs(rho));
      TERM[j * N + k] = l * (Pl - 2.0 * rho);  // l'P l + q'l with q = -2 rho
      // cold start of the free set: the half-gradient at l_k is -dphi/2
      if (!warm) FREE[j * N + k] = (l > 0.0) || (dphi[j] > 0.0);
    }
  }
  for (int o = 16; o > 0; o >>= 1) gs = fmax(gs, __shfl_xor_sync(full, gs, o));
  __syncwarp();
  const double F0 = need_dec ? ordered_sum() : 0.0;
  // ---- primal-dual active set
  int st = 1;
  for (int pit = 0; pit < 96; ++pit) {
    if (pit == 32 && warm) {  // a warm start that does not settle: restart cold
      for (int k = lane; k < N; k += 32) {
        const double dw = wk[k] - wr[k];
        const double dphi[3] = {th * dw, -th * dw, qs * (wk[k] * wk[k] - wr[k] * wr[k])};
#pragma unroll
        for (int j = 0; j < 3; ++j)
          if (j < nb) FREE[j * N + k] = (lk[j * N + k] > 0.0) || (dphi[j] > 0.0);
      }
    }
    for (int k = lane; k < N; k += 32) {
      const double coef[3] = {th, -th, C3[k]};
      double t = 0.0, rhs = 0.0;
#pragma unroll
      for (int j = 0; j < 3; ++j)
        if (j < nb && FREE[j * N + k]) {
          t += coef[j] * coef[j];
          rhs += coef[j] * RHO[j * N + k];
        }
      TD[k] = t;
      U[k] = rhs;
    }
    __syncwarp();
    ric_solve<NT, UF>(N, TD, 2.0 * m * eps * kappa, 2.0 * m * eps, U, V, KS, KAPS, lane);  // z
    __syncwarp();
    for (int k = lane; k < N; k += 32) {
      const double coef[3] = {th, -th, C3[k]};
      const double z = V[k];
      double u = 0.0;
  #pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (j >= nb) break;
        const int i = j * N + k;
        const double l = FREE[i] ? (RHO[i] - coef[j] * z) * inv_eps : 0.0;
        LAM[i] = l;
        u += coef[j] * l;
      }
      U[k] = u;
    }
    __syncwarp();
    abar_solve<NT, UF>(N, FAC, U, V, KAPS, lane);  // v = A_bar^{-1} B' l
    __syncwarp();
    bool same = true;
    for (int k = lane; k < N; k += 32) {
      const double coef[3] = {th, -th, C3[k]};
      const double v = V[k] * inv2m;
  #pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (j >= nb) break;
        const int i = j * N + k;
        const double l = LAM[i];
        const double Pl = eps * l + coef[j] * v;
        const double hg = Pl - RHO[i];  // half gradient
        TERM[i] = l * (Pl - 2.0 * RHO[i]);
        const bool fr = FREE[i];
        const bool nf = fr ? (l > 0.0) : (hg < -1e-12 * gs);
        if (nf != fr) same = false;
        FREE[i] = nf;
      }
    }
    same = __all_sync(full, same);
    __syncwarp();
    if (same) {
      st = 0;
      break;
    }
    if (pit == 0 && g_nnqp_force_fallback) break;
  }
  if (st != 0)  // (rare: see nnqp_lawson_hanson; bit 1 of the status reports its use)
    st = 2 | nnqp_lawson_hanson(N, nb, th, m, eps, kappa, gs, C3, RHO, FAC, LAM, TERM, U, V, TD, KS, KAPS, FREE, lane);
  const double F1 = need_dec ? ordered_sum() : 0.0;
  __syncwarp();
  // ---- write back (price_solver.py:129,135-140)
  for (int k = lane; k < N; k += 32) {
    const double phir[3] = {th * wr[k], th * (cs.w_max - wr[k]), qs * wr[k] * wr[k]};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (j >= nb) break;
      const double ln = fmax(LAM[j * N + k], 0.0);
      if (first) TERM[j * N + k] = (lk[j * N + k] - ln) * phir[j];
      lk[j * N + k] = ln;
    }
  }
  __syncwarp();
  lamdiff_out = first ? ordered_sum() : 0.0;  // valid on lane 0
  dec_pred_out = F0 - F1;                      // valid on lane 0
  status_out = st;
  __syncwarp();
}

// Errors of _get_w_err (price_solver.py:211-214) for a mean trajectory w_sum / n.
__device__ __forceinline__ void price_errors(int N, double kappa, const double* w_sum, double n, const double* w_ref,
                                             double& w_avg_err, double& w0_err) {
  const double inv_n = 1.0 / n;
  double cum = 0.0, e2 = 0.0;
  for (int k = 0; k < N; ++k) {
    const double v = w_sum[k] * inv_n - w_ref[k];
    cum += v;
    e2 += cum * cum + kappa * v * v;
  }
  w_avg_err = sqrt(e2);
  w0_err = fabs(w_sum[0] * inv_n - w_ref[0]);
}

// One WARP per group: errors, the convergence test (price_solver.py:121-127) and -- for
// groups that go on -- the price step.  Dynamic shared memory: per warp
// price_step_scratch_doubles() doubles + r bytes (rounded up to 8).
template <int NT>  // compile-time horizon of the price step's recursions (unrolled), or 0 for any N
__global__ void group_step_kernel(const Consts cs, const PriceArgs p, int it) {
  extern __shared__ double gs_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int g = blockIdx.x * (blockDim.x >> 5) + wib;
  const int G = p.G, N = cs.N, r = p.r;
  if (g >= G || p.skip[g]) return;  // warp-uniform
  const int per_warp = price_step_scratch_doubles(N, r) + (r + 7) / 8;
  double* ws = gs_smem + (size_t)wib * per_warp;
  unsigned char* FREE = reinterpret_cast<unsigned char*>(ws + price_step_scratch_doubles(N, r));
  const double kappa = p.lmbd_r[g] / cs.delta;
  int done = 0;
  if (p.peer_world > 1) {
    // the aggregate exchange, fused: wait for every rank's flag of this iteration, then add the ranks' partial sums
    // of this group in rank order straight out of their memory (same bits on every rank)
    const unsigned long long tag = p.peer_tag0 + (unsigned long long)it + 1;
    int good = 1;
    if (lane == 0) {
      const unsigned long long* flags = reinterpret_cast<const unsigned long long*>(p.peer_region[p.peer_rank]);
      const long long t0 = clock64();
      for (int r = 0; r < p.peer_world && good; ++r)
        while (ld_acquire_sys_u64(flags + r) < tag)
          if (clock64() - t0 > 4000000000LL) {  // ~2 s: a peer is gone - report instead of hanging the GPU
            good = 0;
            break;
          }
      if (!good) atomicExch(p.peer_timeout, 1);
    }
    good = __shfl_sync(0xffffffffu, good, 0);
    for (int k = lane; k < N; k += 32) {
      double sum = 0.0;
      if (good)
        for (int r = 0; r < p.peer_world; ++r) {
          const double* part = reinterpret_cast<const double*>(p.peer_region[r] + kPeerFlagBytes) +
                               (size_t)(it & 1) * G * N + (size_t)g * N;
          sum += ld_relaxed_sys_f64(part + k);
        }
      p.w_avg[(size_t)g * N + k] = sum;
    }
    __syncwarp();
  }
  if (p.cs_w_ev) {
    // colsum_kernel's work for this group, by its own warp (one launch less per iteration): lane k adds column k in
    // EV order (price_solver.py:205) with eight independent loads in flight - the same additions, the same bits
    const int b0 = p.cs_off[g], b1 = p.cs_off[g + 1];
    for (int k = lane; k < N; k += 32) {
      double sum = 0.0;
      const double* col = p.cs_w_ev + k;
      int b = b0;
      for (; b + 8 <= b1; b += 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = col[(int64_t)(b + u) * N];
#pragma unroll
        for (int u = 0; u < 8; ++u) sum += v[u];
      }
      for (; b < b1; ++b) sum += col[(int64_t)b * N];
      p.w_avg[(size_t)g * N + k] = sum;  // the SUM: price_errors divides by cnt
    }
    if (p.cs_err_ev) {
      double m = 0.0;
      for (int b = b0 + lane; b < b1; b += 32) m = fmax(m, p.cs_err_ev[b]);
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, d));
      if (lane == 0) p.w_err_max[g] = m;
    }
    __syncwarp();
  }
  if (lane == 0) {
    double w_avg_err, w0_err;
    price_errors(N, kappa, p.w_avg + (size_t)g * N, p.cnt ? p.cnt[g] : 1.0, p.w_ref + (size_t)g * N, w_avg_err,
                 w0_err);
    p.w_avg_err[g] = w_avg_err;
    p.w0_err[g] = w0_err;
    const double tol = sqrt((double)N) * p.y0_rng[g] + p.eps_tol;  // price_solver.py:184
    const double w_err = p.tol_type_max ? p.w_err_max[g] : w_avg_err;
    if (w_err <= tol) {  // price_solver.py:125
      p.skip[g] = 1;
      p.iters[g] = it;
      done = 1;
    } else {
      atomicAdd(p.n_active, 1);
    }
  }
  done = __shfl_sync(0xffffffffu, done, 0);
  if (done) return;
  const bool warm = it > 0 && !p.cold;
  unsigned char* gfree = p.wsb + (size_t)g * r;  // the free set lives in global memory between launches
  if (warm)
    for (int i = lane; i < r; i += 32) FREE[i] = gfree[i];
  __syncwarp();
  double lamdiff, dec;
  int st;
  price_step_warp<NT>(cs, r, kappa, p.eps_reg, p.lmbd + (size_t)g * 3 * N, p.w_k + (size_t)g * N,
                      p.w_ref + (size_t)g * N, ws, FREE, lane, it == 0, warm, p.hist_ac != nullptr || p.want_dec, false,
                      lamdiff, dec, st);
  for (int i = lane; i < r; i += 32) gfree[i] = FREE[i];
  if (lane == 0) {
    p.nnqp_status[g] = st;
    p.lamdiff_phi[g] = lamdiff;
    p.dec_pred[g] = dec;
  }
}

__global__ void bookkeep_kernel(const PriceArgs p, int it) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g == 0 && p.publish_ring) {
    // last kernel of the iteration: publish the number of still-active groups to the host's pinned ring (a mapped
    // store the host polls without a CUDA call; tag = it + 1 marks the slot as written) and reset the counter
    volatile int32_t* slot = p.publish_ring + 2 * (it % p.publish_slots);
    slot[1] = *p.n_active;
    __threadfence_system();
    slot[0] = it + 1;
    *p.n_active = 0;
  }
  if (g >= p.G || p.skip[g]) return;
  const double ac = p.cost_new[g] - p.dual_cost[g] + p.lamdiff_phi[g];  // price_solver.py:135-137
  p.dual_cost[g] = p.cost_new[g];
  if (p.hist_ac && it < p.hist_cap) {
    p.hist_ac[(size_t)g * p.hist_cap + it] = ac;
    p.hist_pred[(size_t)g * p.hist_cap + it] = p.dec_pred[g];
  }
}

// price_solver.py:145-147,248-255 with the LP of price_regularizer.py:68-85 in closed form:
// row k of Dphi' touches only (x1_k, x2_k, x3_k); b_k = theta(l1-l2) + 2 q w_k l3;
// b_k < 0: x2 = -b_k/theta;  b_k >= 0: x3 = b_k/(2 q w_k) if r = 3N and w_k > 0 (unit cost
// w_k/2 beats x1's w_k) else x1 = b_k/theta.  "w_k > 0" means w_k above the LoMPC solver's eps-band
// (1e-9 w_max): for a w_k of rounding-error size the LP is degenerate and x3 = b_k/(2 q w_k) would be ~1e11.
__device__ __forceinline__ void regularize_core(const Consts& cs, int r, const double* w, double* l, double& pre_out,
                                                double& post_out) {
  const int N = cs.N;
  const double th = cs.theta, qs = cs.q_scale, wm = cs.w_max;
  double pre = 0.0, post = 0.0;
  for (int k = 0; k < N; ++k) {
    const double wk = w[k];
    const double ph1 = th * wk, ph2 = th * (wm - wk), ph3 = qs * wk * wk;
    const double l3 = (r == 3 * N) ? l[2 * N + k] : 0.0;
    pre += ph1 * l[k] + ph2 * l[N + k] + ph3 * l3;
    const double bk = th * (l[k] - l[N + k]) + 2.0 * qs * wk * l3;
    double x1 = 0.0, x2 = 0.0, x3 = 0.0;
    if (bk < 0.0) x2 = -bk / th;
    else if (r == 3 * N && wk > 1e-9 * wm) x3 = bk / (2.0 * qs * wk);  // (w_k inside the LoMPC solver's eps-band of 0 counts as 0)
    else x1 = bk / th;
    l[k] = x1;
    l[N + k] = x2;
    if (r == 3 * N) l[2 * N + k] = x3;
    post += ph1 * x1 + ph2 * x2 + ph3 * x3;
  }
  pre_out = pre;
  post_out = post;
}

__global__ void regularize_kernel(const Consts cs, int G, int r, const double* __restrict__ w_k,
                                  double* __restrict__ lmbd, double* __restrict__ price_pre,
                                  double* __restrict__ price_post, const int32_t* __restrict__ empty) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  if (empty && empty[g]) return;
  double pre, post;
  regularize_core(cs, r, w_k + (size_t)g * cs.N, lmbd + (size_t)g * 3 * cs.N, pre, post);
  price_pre[g] = pre;
  price_post[g] = post;
}

// PriceRegularizer.solve_price_regularization (price_regularizer.py:68-85) for a constraint
// matrix with the block-diagonal pattern A = [diag(a_0) ... diag(a_{nb-1})] (every use in the
// reference: Dphi' and the [I, -I] of test_price_regularizer.py): the LP separates into N
// one-row LPs  min c'x, a'x = b_k, x >= 0, whose optimum puts all weight on the column with
// the smallest cost per unit of b_k.  status: 0 ok, 1 infeasible row.
__global__ void lp_rows_kernel(int N, int nb, const double* __restrict__ a, const double* __restrict__ b,
                               const double* __restrict__ c, double* __restrict__ x, int32_t* __restrict__ status) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= N) return;
  for (int j = 0; j < nb; ++j) x[j * N + k] = 0.0;
  const double bk = b[k];
  if (bk == 0.0) return;
  int best = -1;
  double best_ratio = 0.0;
  for (int j = 0; j < nb; ++j) {
    const double aj = a[j * N + k];
    if (aj * bk > 0.0) {
      const double ratio = c[j * N + k] / fabs(aj);
      if (best < 0 || ratio < best_ratio) {
        best = j;
        best_ratio = ratio;
      }
    }
  }
  if (best < 0) {
    atomicExch(status, 1);
    return;
  }
  x[best * N + k] = bk / a[best * N + k];
}

__global__ void mean_kernel(int G, const int32_t* __restrict__ off, const double* __restrict__ x_ev,
                            double* __restrict__ mean) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  const int b0 = off[g], b1 = off[g + 1];
  double s = 0.0;
  for (int b = b0; b < b1; ++b) s += x_ev[b];
  mean[g] = b1 > b0 ? s / (b1 - b0) : 0.0;
}

// Helpers of the per-partition fallback of price_solve_chain_dev (horizons without a fused kernel).
__global__ void rebase_offsets_kernel(int S, const int32_t* __restrict__ off_abs, int32_t* __restrict__ off_reb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= S) off_reb[i] = off_abs[i] - off_abs[0];
}

// rows[s,:] = the group is non-empty ? prev[s,:] : 0   (charging_station.py:270,286)
__global__ void chain_rows_kernel(int S, int row, const int32_t* __restrict__ off_abs, const double* __restrict__ prev,
                                  double* __restrict__ rows) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)S * row) return;
  const int s = (int)(i / row);
  rows[i] = off_abs[s + 1] > off_abs[s] ? prev[i] : 0.0;
}

__global__ void init_groups_kernel(int G, int max_iter, const int32_t* __restrict__ off,
                                   int32_t* __restrict__ skip, int32_t* __restrict__ iters,
                                   int32_t* __restrict__ nnqp_status) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  skip[g] = off[g + 1] <= off[g];  // empty groups are never solved (charging_station.py:277,293)
  iters[g] = max_iter - 1;         // value of `iter` if the loop never breaks (price_solver.py:111)
  nnqp_status[g] = 0;
}

}  // namespace lompc
