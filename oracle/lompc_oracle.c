/*
 * CPU ORACLE in C (test / baseline infrastructure, NOT a product path).
 *
 * Restates the reference's CPU implementation of LoMPC.solve_lompc
 * (lompc.py:137-156): cvxpy hands the QP of lompc.py:73-135 to CLARABEL
 * (settings.py:11), a primal-dual interior-point method.  cvxpy/Clarabel are
 * not in /root/reference and not installable here (PARITY UNPINNED, see
 * oracle/lompc_oracle.py), so this file restates the published algorithm class:
 * a Mehrotra predictor-corrector IPM on
 *       min 1/2 x'Px + q'x   s.t.  Gx + s = h, s >= 0
 * in the canonical form cvxpy emits (small EV: x = w, rows -w<=0, w<=w_max;
 * large EV: x = [w; t] with the four epigraph rows of cv.maximum,
 * lompc.py:111-115), stopped at Clarabel's default 1e-8 tolerances.  The dense
 * normal equations exploit that every row of G has at most two non-zeros.
 *
 * It also holds the C twin of the EXACT active-set oracle (oracle/lompc_oracle.py::solve_active_set),
 * which the price-loop / closed-loop oracles use at the reference's full sizes.
 *
 * Used by: tests/ (checked against oracle/lompc_oracle.py) and bench.py's
 * cpu_baseline / --impl reference legs (timed on all host threads, OpenMP).
 * Build: make -C oracle   ->  oracle/liblompc_oracle.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXN 128           /* horizon limit of this oracle */
#define MAXV (2 * MAXN)    /* variables (large EV: w and t) */
#define MAXM (6 * MAXN)    /* inequality rows */

typedef struct {
  int N, large;
  double delta, theta, y_max, w_max;
} oracle_consts;

static const double PWL_A[4] = {0.0, 1.0, 1.5, 2.0};       /* lompc.py:113-114 slopes   */
static const double PWL_O[4] = {0.0, 0.125, 0.375, 0.75};  /* lompc.py:113-114 offsets  */

/* objective exactly as the cvxpy expression tree (lompc.py:101-135) */
static double lompc_cost(const oracle_consts* c, const double* w, const double* lmbd,
                         double lmbd_r, double gamma) {
  const int N = c->N;
  const double th = c->theta, wm = c->w_max;
  const double qs = 3.0 * th / (4.0 * wm); /* lompc.py:67 */
  double cost = 0.0, y = 0.0, sy2 = 0.0, sy = 0.0, lp = 0.0, qp = 0.0, rp = 0.0;
  for (int k = 0; k < N; ++k) {
    if (!c->large) {
      cost += th * th * (w[k] / 0.9) * (w[k] / 0.9); /* lompc.py:107 */
    } else {
      const double x = w[k] / wm;
      double m = 0.0;
      for (int j = 1; j < 4; ++j) m = fmax(m, PWL_A[j] * x - PWL_O[j]);
      cost += (th * wm) * (th * wm) * m; /* lompc.py:116 */
    }
    y += w[k];
    sy2 += y * y;
    sy += y;
    lp += lmbd[k] * w[k] + lmbd[N + k] * (wm - w[k]);
    qp += lmbd[2 * N + k] * w[k] * w[k];
    rp += w[k] * w[k];
  }
  cost += c->delta * th * th * (sy2 - 2.0 * gamma * sy); /* lompc.py:120-124 */
  return cost + th * lp + qs * qp + lmbd_r * th * th * rp; /* lompc.py:128-136 */
}

/* in-place dense Cholesky K = L L' (lower), returns 0 on success */
static int chol(double* K, int n) {
  for (int j = 0; j < n; ++j) {
    double d = K[j * n + j];
    for (int p = 0; p < j; ++p) d -= K[j * n + p] * K[j * n + p];
    if (!(d > 0.0)) return 1;
    d = sqrt(d);
    K[j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double v = K[i * n + j];
      for (int p = 0; p < j; ++p) v -= K[i * n + p] * K[j * n + p];
      K[i * n + j] = v / d;
    }
  }
  return 0;
}
static void chol_solve(const double* L, int n, double* b) {
  for (int i = 0; i < n; ++i) {
    double v = b[i];
    for (int p = 0; p < i; ++p) v -= L[i * n + p] * b[p];
    b[i] = v / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double v = b[i];
    for (int p = i + 1; p < n; ++p) v -= L[p * n + i] * b[p];
    b[i] = v / L[i * n + i];
  }
}

/* G has rows r = blk*N + k: blk 0: -w_k ; blk 1: +w_k ; blk 2+j: (a_j/wm) w_k - t_k */
static inline void G_apply(const oracle_consts* c, const double* x, double* out) {
  const int N = c->N;
  for (int k = 0; k < N; ++k) {
    out[k] = -x[k];
    out[N + k] = x[k];
    if (c->large)
      for (int j = 0; j < 4; ++j) out[(2 + j) * N + k] = PWL_A[j] / c->w_max * x[k] - x[N + k];
  }
}
static inline void Gt_apply(const oracle_consts* c, const double* z, double* out) {
  const int N = c->N;
  for (int k = 0; k < N; ++k) {
    double a = -z[k] + z[N + k], t = 0.0;
    if (c->large)
      for (int j = 0; j < 4; ++j) {
        a += PWL_A[j] / c->w_max * z[(2 + j) * N + k];
        t -= z[(2 + j) * N + k];
      }
    out[k] = a;
    if (c->large) out[N + k] = t;
  }
}

/* One QP.  Returns the IPM iteration count (negative on numerical failure). */
static int solve_one(const oracle_consts* c, const double* lmbd, double lmbd_r, double gamma,
                     double tol, int max_iter, double* w_out, double* cost_out) {
  const int N = c->N, n = c->large ? 2 * N : N, m = c->large ? 6 * N : 2 * N;
  const double th = c->theta, wm = c->w_max;
  const double qs = 3.0 * th / (4.0 * wm), cc = 2.0 * c->delta * th * th;
  static _Thread_local double P[MAXV * MAXV], K[MAXV * MAXV];
  double q[MAXV], h[MAXM], x[MAXV], s[MAXM], z[MAXM], rd[MAXV], rp[MAXM], rc[MAXM];
  double dxa[MAXV], dsa[MAXM], dza[MAXM], dx[MAXV], ds[MAXM], dz[MAXM], tmp[MAXM], tv[MAXV];
  /* P = blkdiag(H, 0), H = diag(d) + c A'A, (A'A)_ij = N - max(i,j)  (lompc.py:69,119-124) */
  memset(P, 0, sizeof(double) * n * n);
  for (int i = 0; i < N; ++i) {
    for (int j = 0; j < N; ++j) P[i * n + j] = cc * (N - (i > j ? i : j));
    double d = 2.0 * (lmbd_r * th * th + qs * lmbd[2 * N + i]);
    if (!c->large) d += 2.0 * th * th / 0.81;
    P[i * n + i] += d;
    q[i] = th * (lmbd[i] - lmbd[N + i]) - cc * gamma * (N - i);
    if (c->large) q[N + i] = (th * wm) * (th * wm);
  }
  double qinf = 0.0, hinf = 0.0;
  for (int k = 0; k < N; ++k) {
    h[k] = 0.0;
    h[N + k] = wm;
    if (c->large)
      for (int j = 0; j < 4; ++j) h[(2 + j) * N + k] = PWL_O[j];
  }
  for (int i = 0; i < n; ++i) qinf = fmax(qinf, fabs(q[i]));
  for (int i = 0; i < m; ++i) hinf = fmax(hinf, fabs(h[i]));
  for (int k = 0; k < N; ++k) {
    x[k] = 0.5 * wm;
    if (c->large) x[N + k] = 1.0;
  }
  G_apply(c, x, tmp);
  for (int i = 0; i < m; ++i) {
    s[i] = fmax(h[i] - tmp[i], 1e-2);
    z[i] = 1.0;
  }
  int it;
  for (it = 0; it < max_iter; ++it) {
    /* residuals */
    Gt_apply(c, z, tv);
    double xPx = 0.0, qx = 0.0, hz = 0.0, mu = 0.0, rdinf = 0.0, rpinf = 0.0;
    for (int i = 0; i < n; ++i) {
      double v = 0.0;
      for (int j = 0; j < n; ++j) v += P[i * n + j] * x[j];
      xPx += x[i] * v;
      qx += q[i] * x[i];
      rd[i] = v + q[i] + tv[i];
      rdinf = fmax(rdinf, fabs(rd[i]));
    }
    G_apply(c, x, tmp);
    for (int i = 0; i < m; ++i) {
      rp[i] = tmp[i] + s[i] - h[i];
      rpinf = fmax(rpinf, fabs(rp[i]));
      hz += h[i] * z[i];
      mu += s[i] * z[i];
    }
    mu /= m;
    const double pobj = 0.5 * xPx + qx, dobj = -0.5 * xPx - hz;
    if (rdinf <= tol * fmax(1.0, qinf) && rpinf <= tol * fmax(1.0, hinf) &&
        fabs(pobj - dobj) <= tol * fmax(1.0, fmin(fabs(pobj), fabs(dobj))))
      break;
    /* K = P + G' diag(z/s) G : only the (w,w), (w,t), (t,t) diagonals change */
    memcpy(K, P, sizeof(double) * n * n);
    for (int k = 0; k < N; ++k) {
      double ww = z[k] / s[k] + z[N + k] / s[N + k], wt = 0.0, tt = 0.0;
      if (c->large)
        for (int j = 0; j < 4; ++j) {
          const double wj = z[(2 + j) * N + k] / s[(2 + j) * N + k], a = PWL_A[j] / wm;
          ww += a * a * wj;
          wt -= a * wj;
          tt += wj;
        }
      K[k * n + k] += ww;
      if (c->large) {
        K[k * n + N + k] += wt;
        K[(N + k) * n + k] += wt;
        K[(N + k) * n + N + k] += tt;
      }
    }
    if (chol(K, n)) { /* lost positive definiteness at round-off level: regularise once, else stop */
      memcpy(K, P, sizeof(double) * n * n);
      for (int k = 0; k < N; ++k) {
        double ww = z[k] / s[k] + z[N + k] / s[N + k], wt = 0.0, tt = 0.0;
        if (c->large)
          for (int j = 0; j < 4; ++j) {
            const double wj = z[(2 + j) * N + k] / s[(2 + j) * N + k], a = PWL_A[j] / wm;
            ww += a * a * wj;
            wt -= a * wj;
            tt += wj;
          }
        K[k * n + k] += ww * (1.0 + 1e-9);
        if (c->large) {
          K[k * n + N + k] += wt;
          K[(N + k) * n + k] += wt;
          K[(N + k) * n + N + k] += tt * (1.0 + 1e-9);
        }
      }
      if (chol(K, n)) break;
    }
    for (int pass = 0; pass < 2; ++pass) {
      double* ddx = pass ? dx : dxa;
      double* dds = pass ? ds : dsa;
      double* ddz = pass ? dz : dza;
      if (pass == 0) {
        for (int i = 0; i < m; ++i) rc[i] = s[i] * z[i];
      } else {
        /* Mehrotra centering + second-order correction */
        double aaff = 1.0;
        for (int i = 0; i < m; ++i) {
          if (dsa[i] < 0.0) aaff = fmin(aaff, -s[i] / dsa[i]);
          if (dza[i] < 0.0) aaff = fmin(aaff, -z[i] / dza[i]);
        }
        double mua = 0.0;
        for (int i = 0; i < m; ++i) mua += (s[i] + aaff * dsa[i]) * (z[i] + aaff * dza[i]);
        mua /= m;
        const double sig = (mua / mu) * (mua / mu) * (mua / mu);
        for (int i = 0; i < m; ++i) rc[i] = s[i] * z[i] + dsa[i] * dza[i] - sig * mu;
      }
      /* (P + G'WG) dx = -rd - G'((-rc + z rp)/s) ; ds = -rp - G dx ; dz = (-rc - z ds)/s */
      for (int i = 0; i < m; ++i) tmp[i] = (-rc[i] + z[i] * rp[i]) / s[i];
      Gt_apply(c, tmp, tv);
      for (int i = 0; i < n; ++i) ddx[i] = -rd[i] - tv[i];
      chol_solve(K, n, ddx);
      G_apply(c, ddx, tmp);
      for (int i = 0; i < m; ++i) {
        dds[i] = -rp[i] - tmp[i];
        ddz[i] = (-rc[i] - z[i] * dds[i]) / s[i];
      }
    }
    double a = 1.0;
    for (int i = 0; i < m; ++i) {
      if (ds[i] < 0.0) a = fmin(a, -s[i] / ds[i]);
      if (dz[i] < 0.0) a = fmin(a, -z[i] / dz[i]);
    }
    a = fmin(1.0, 0.99 * a);
    for (int i = 0; i < n; ++i) x[i] += a * dx[i];
    for (int i = 0; i < m; ++i) {
      s[i] += a * ds[i];
      z[i] += a * dz[i];
    }
  }
  memcpy(w_out, x, sizeof(double) * N);
  *cost_out = lompc_cost(c, x, lmbd, lmbd_r, gamma); /* self.cost.value, lompc.py:155 */
  return it;
}

/* Batched entry point (OpenMP over QPs).  Returns the thread count used, or <0. */
int oracle_solve_lompc_batch(int N, double delta, double theta, double y_max, double w_max,
                             int large, int64_t B, const double* lmbd, int64_t lmbd_stride,
                             const double* lmbd_r, int64_t lmbd_r_stride, const double* gamma,
                             double tol, int max_iter, int nthreads, double* w_out,
                             double* cost_out, int32_t* iters) {
  if (N < 1 || N > MAXN) return -1;
  oracle_consts c = {N, large, delta, theta, y_max, w_max};
  int used = 1;
#ifdef _OPENMP
  omp_set_num_threads(nthreads > 0 ? nthreads : omp_get_num_procs());
  used = omp_get_max_threads();
#endif
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t b = 0; b < B; ++b) {
    double cost;
    int it = solve_one(&c, lmbd + b * lmbd_stride, lmbd_r[b * lmbd_r_stride], gamma[b], tol,
                       max_iter, w_out + b * N, &cost);
    cost_out[b] = cost;
    if (iters) iters[b] = it;
  }
  return used;
}

/* ------------------------------------------------------------------------------------------------
 * Exact solve: the C twin of oracle/lompc_oracle.py::solve_active_set (same steps, same tie rules),
 * so that the price-loop and closed-loop oracles can afford the reference's full sizes (500 + 500
 * EVs, 49 steps).  Each coordinate is FREE inside a piece of the pwl (linear term = that piece's
 * slope) or FIXED on a breakpoint; a feasible-descent primal active-set iteration with one change
 * per step and dense Cholesky solves on the free block of H = diag(d) + c A'A, (A'A)_ij = N - max(i,j).
 * ------------------------------------------------------------------------------------------------ */
static int solve_exact_one(const oracle_consts* c, const double* lmbd, double lmbd_r, double gamma,
                           int max_iter, double* w_out, double* cost_out) {
  const int N = c->N;
  const double th = c->theta, wm = c->w_max;
  const double qs = 3.0 * th / (4.0 * wm), cc = 2.0 * c->delta * th * th;
  double d[MAXN], g[MAXN], w[MAXN], wt[MAXN], rhs[MAXN], grad[MAXN];
  static _Thread_local double Kf[MAXN * MAXN];
  int seg[MAXN], at[MAXN], fixed[MAXN], idx[MAXN];
  double brk[5], slope[4];
  int nseg;
  if (!c->large) {
    nseg = 1; brk[0] = 0.0; brk[1] = wm; slope[0] = 0.0;
  } else {
    nseg = 4;
    const double scale = (th * wm) * (th * wm) / wm;
    brk[0] = 0.0; brk[1] = 0.125 * wm; brk[2] = 0.5 * wm; brk[3] = 0.75 * wm; brk[4] = wm;
    for (int j = 0; j < 4; ++j) slope[j] = scale * PWL_A[j];
  }
  double gmax = 0.0;
  for (int k = 0; k < N; ++k) {
    d[k] = 2.0 * (lmbd_r * th * th + qs * lmbd[2 * N + k]) + (c->large ? 0.0 : 2.0 * th * th / 0.81);
    g[k] = th * (lmbd[k] - lmbd[N + k]) - cc * gamma * (double)(N - k);
    gmax = fmax(gmax, fabs(g[k]));
    w[k] = 0.0; seg[k] = 0; at[k] = 0; fixed[k] = 1;
  }
  int it = 0;
  for (it = 0; it < max_iter; ++it) {
    int nf = 0;
    for (int k = 0; k < N; ++k) if (!fixed[k]) idx[nf++] = k;
    memcpy(wt, w, sizeof(double) * N);
    if (nf > 0) {
      for (int a = 0; a < nf; ++a) {
        const int i = idx[a];
        double r = -(g[i] + slope[seg[i]]);
        for (int k = 0; k < N; ++k)
          if (fixed[k]) r -= cc * (double)(N - (i > k ? i : k)) * w[k];
        rhs[a] = r;
        for (int b2 = 0; b2 < nf; ++b2) {
          const int j = idx[b2];
          Kf[a * nf + b2] = cc * (double)(N - (i > j ? i : j)) + (i == j ? d[i] : 0.0);
        }
      }
      if (chol(Kf, nf)) return -(it + 1);
      chol_solve(Kf, nf, rhs);
      for (int a = 0; a < nf; ++a) wt[idx[a]] = rhs[a];
    }
    /* ratio test against the ends of the pieces the free coordinates live in */
    double alpha = 1.0;
    int blk = -1, blk_at = -1;
    for (int a = 0; a < nf; ++a) {
      const int k = idx[a];
      const double lo = brk[seg[k]], hi = brk[seg[k] + 1], dk = wt[k] - w[k];
      if (dk < 0 && wt[k] < lo) {
        const double al = (lo - w[k]) / dk;
        if (al < alpha) { alpha = al; blk = k; blk_at = seg[k]; }
      } else if (dk > 0 && wt[k] > hi) {
        const double al = (hi - w[k]) / dk;
        if (al < alpha) { alpha = al; blk = k; blk_at = seg[k] + 1; }
      }
    }
    if (blk >= 0) {
      for (int k = 0; k < N; ++k) w[k] += alpha * (wt[k] - w[k]);
      fixed[blk] = 1; at[blk] = blk_at; w[blk] = brk[blk_at];
      continue;
    }
    memcpy(w, wt, sizeof(double) * N);
    /* multipliers of the fixed coordinates: need -grad_k in [slope_left, slope_right] */
    for (int i = 0; i < N; ++i) {
      double r = d[i] * w[i] + g[i];
      for (int k = 0; k < N; ++k) r += cc * (double)(N - (i > k ? i : k)) * w[k];
      grad[i] = r;
    }
    double worst = 0.0;
    int wk = -1, wseg = -1;
    for (int k = 0; k < N; ++k) {
      if (!fixed[k]) continue;
      const int i = at[k];
      const double s_lo = i > 0 ? slope[i - 1] : -INFINITY, s_hi = i < nseg ? slope[i] : INFINITY;
      if (-grad[k] < s_lo && s_lo + grad[k] > worst) { worst = s_lo + grad[k]; wk = k; wseg = i - 1; }
      else if (-grad[k] > s_hi && -grad[k] - s_hi > worst) { worst = -grad[k] - s_hi; wk = k; wseg = i; }
    }
    if (wk < 0 || worst <= 1e-13 * fmax(1.0, gmax)) break;
    fixed[wk] = 0; seg[wk] = wseg;
  }
  memcpy(w_out, w, sizeof(double) * N);
  *cost_out = lompc_cost(c, w, lmbd, lmbd_r, gamma);
  return it;
}

/* Batched exact solve (OpenMP over QPs); same conventions as oracle_solve_lompc_batch. */
int oracle_solve_lompc_exact_batch(int N, double delta, double theta, double y_max, double w_max,
                                   int large, int64_t B, const double* lmbd, int64_t lmbd_stride,
                                   const double* lmbd_r, int64_t lmbd_r_stride, const double* gamma,
                                   int max_iter, int nthreads, double* w_out, double* cost_out,
                                   int32_t* iters) {
  if (N < 1 || N > MAXN) return -1;
  oracle_consts c = {N, large, delta, theta, y_max, w_max};
  int used = 1;
#ifdef _OPENMP
  omp_set_num_threads(nthreads > 0 ? nthreads : omp_get_num_procs());
  used = omp_get_max_threads();
#endif
#pragma omp parallel for schedule(dynamic, 8)
  for (int64_t b = 0; b < B; ++b) {
    double cost;
    int it = solve_exact_one(&c, lmbd + b * lmbd_stride, lmbd_r[b * lmbd_r_stride], gamma[b], max_iter,
                             w_out + b * N, &cost);
    cost_out[b] = cost;
    if (iters) iters[b] = it;
  }
  return used;
}

double oracle_lompc_cost(int N, double delta, double theta, double y_max, double w_max, int large,
                         const double* w, const double* lmbd, double lmbd_r, double gamma) {
  oracle_consts c = {N, large, delta, theta, y_max, w_max};
  return lompc_cost(&c, w, lmbd, lmbd_r, gamma);
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_num_procs();
#else
  return 1;
#endif
}
