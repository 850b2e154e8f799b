"""CPU ORACLE (test infrastructure, NOT a product path) for the upper-level BiMPC.

Restates ``chargingstation/bimpc.py`` of the reference (the convex program cvxpy hands
to CLARABEL's power cone, bimpc.py:142-292) with DENSE numpy algebra:

    min  c_g sum_k u_g[k]^1.7                                                (bimpc.py:220-221)
         + delta * sum_p sum_k omega_k [ a_sp (cumsum(w_s[p])_k - gamma_sm[p])^2
                                        + a_lp (cumsum(w_l[p])_k - gamma_lm[p])^2 ]   (:233-265)
    s.t. 0 <= w_s <= w_max_s, 0 <= w_l <= w_max_l, 0 <= u_g <= u_g_max      (:143-186)
         u_b = u_g - demand - theta_s Mp_s' w_s - theta_l Mp_l' w_l
         -u_b_max + d e1 <= u_b <= u_b_max - d e1                            (:188-203)
         d <= x0 + cumsum(u_b) <= x_max - d                                  (:205-218)
         d = theta_s Mp_s.beta_s + theta_l Mp_l.beta_l

with a_p = (theta Mp_p)^2 (WEIGHTED) or 1, omega_k = exp_rate^(k-N+1) (EXP_UNWEIGHTED) or 1.

PARITY UNPINNED against the real reference: cvxpy/CLARABEL are not installable here and
``test/test_bimpc.py`` only plots.  Pinning used instead: the objective is strictly convex
in u_g and in every cumsum(w_p) with omega_k > 0, hence the optimum is unique; the dense
Mehrotra interior-point solve below is cross-checked against scipy's SLSQP /
trust-constr on small instances (tests/test_bimpc_oracle.py) and certified by
``kkt_certificate`` (feasibility, stationarity and complementarity with multipliers
reconstructed by NNLS), a solver-independent test.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

WEIGHTED, UNWEIGHTED, EXP_UNWEIGHTED = 0, 1, 2  # BiMPCChargingCostType, bimpc.py:12-15


@dataclass
class BiConsts:
    N: int
    P: int
    delta: float
    c_g: float
    u_g_max: float
    u_b_max: float
    x_max: float
    cost_type: int
    exp_rate: float
    theta_s: float
    theta_l: float
    w_max_s: float
    w_max_l: float


def example_consts(N: int = 16, P: int = 12) -> BiConsts:
    """example/real_time_price_control.py:26-52."""
    return BiConsts(N, P, 1e3, 1.0, 1.0, 0.3, 0.3, EXP_UNWEIGHTED, 5.0, 10.0, 50.0, 0.25, 0.15)


def stage_weights(c: BiConsts) -> np.ndarray:
    """bimpc.py:255-257 (ones for the other two cost types)."""
    if c.cost_type == EXP_UNWEIGHTED:
        with np.errstate(all="ignore"):
            return np.power(float(c.exp_rate), np.arange(-c.N + 1, 1, 1).astype(float))
    return np.ones(c.N)


def assemble(c: BiConsts, Mp_s, Mp_l, beta_s, beta_l, gamma_sm, gamma_lm, x0, demand):
    """Dense data of  min c_g sum(u^1.7) + x'Hq x/2 + gq'x + const  s.t. G x <= h,
    x = [w_s (P*N), w_l (P*N), u_g (N)]."""
    N, P = c.N, c.P
    nw = P * N
    n = 2 * nw + N
    A = np.tril(np.ones((N, N)))
    I_N = np.eye(N)
    M = np.zeros((N, 2 * nw))  # aggregate EV load, bimpc.py:189-194
    for p in range(P):
        M[:, p * N:(p + 1) * N] = c.theta_s * Mp_s[p] * I_N
        M[:, nw + p * N:nw + (p + 1) * N] = c.theta_l * Mp_l[p] * I_N
    d_err = c.theta_s * float(Mp_s @ beta_s) + c.theta_l * float(Mp_l @ beta_l)
    e1 = np.zeros(N)
    e1[0] = 1.0
    Ub = np.hstack([-M, I_N])  # u_b = Ub x - demand
    AUb = A @ Ub
    Ad = A @ demand
    ub = np.concatenate([np.full(nw, c.w_max_s), np.full(nw, c.w_max_l), np.full(N, c.u_g_max)])
    G = np.vstack([-np.eye(n), np.eye(n), -Ub, Ub, -AUb, AUb])
    h = np.concatenate([np.zeros(n), ub,
                        c.u_b_max - d_err * e1 - demand, c.u_b_max - d_err * e1 + demand,
                        x0 - d_err - Ad, c.x_max - d_err - x0 + Ad])
    om = stage_weights(c)
    if c.cost_type == WEIGHTED:
        a_s, a_l = (c.theta_s * Mp_s) ** 2, (c.theta_l * Mp_l) ** 2
    else:
        a_s, a_l = np.ones(P), np.ones(P)
    AtWA = A.T @ (om[:, None] * A)
    At_om = A.T @ om
    Hq = np.zeros((n, n))
    gq = np.zeros(n)
    const = 0.0
    for p in range(P):
        for base, a, gam in ((p * N, a_s[p], gamma_sm[p]), (nw + p * N, a_l[p], gamma_lm[p])):
            sl = slice(base, base + N)
            Hq[sl, sl] = 2 * c.delta * a * AtWA
            gq[sl] = -2 * c.delta * a * gam * At_om
            const += c.delta * a * gam ** 2 * np.sum(om)
    return n, nw, G, h, Hq, gq, const


def objective(c: BiConsts, x, nw, Hq, gq, const) -> float:
    u = x[2 * nw:]
    return float(c.c_g * np.sum(np.maximum(u, 0.0) ** 1.7) + 0.5 * x @ Hq @ x + gq @ x + const)


def solve_ipm(c: BiConsts, Mp_s, Mp_l, beta_s, beta_l, gamma_sm, gamma_lm, x0, demand,
              tol: float = 1e-9, max_iter: int = 100, trace: list | None = None):
    """Mehrotra predictor-corrector on the dense KKT system.  Returns
    (w_hat_s [P,N], w_hat_l [P,N], u_g [N], info)."""
    N, P = c.N, c.P
    n, nw, G, h, Hq, gq, const = assemble(c, np.asarray(Mp_s, float), np.asarray(Mp_l, float),
                                          np.asarray(beta_s, float), np.asarray(beta_l, float),
                                          np.asarray(gamma_sm, float), np.asarray(gamma_lm, float),
                                          float(x0), np.asarray(demand, float))
    m = G.shape[0]
    iu = np.arange(2 * nw, n)
    x = np.concatenate([np.full(nw, 0.5 * c.w_max_s), np.full(nw, 0.5 * c.w_max_l), np.full(N, 0.5 * c.u_g_max)])
    s = np.maximum(h - G @ x, 1e-2)
    z = np.ones(m)
    scale_g = max(1.0, float(np.max(np.abs(gq))))
    it, status = 0, 1
    for it in range(max_iter + 1):
        u = np.maximum(x[iu], 1e-300)
        grad = Hq @ x + gq
        grad[iu] += 1.7 * c.c_g * u ** 0.7
        hdiag = np.zeros(n)
        hdiag[iu] = 1.7 * 0.7 * c.c_g * u ** (-0.3)
        rd = grad + G.T @ z
        rp = G @ x + s - h
        mu = float(s @ z) / m
        if trace is not None:
            trace.append((float(np.max(np.abs(rd))), float(np.max(np.abs(rp))), mu))
        if np.max(np.abs(rd)) <= tol * scale_g and np.max(np.abs(rp)) <= tol and mu <= tol:
            status = 0
            break
        if it == max_iter:
            break
        W = z / s
        K = Hq + np.diag(hdiag) + G.T @ (W[:, None] * G)
        try:
            L = np.linalg.cholesky(K)
            ksolve = lambda b: np.linalg.solve(L.T, np.linalg.solve(L, b))  # noqa: E731
        except np.linalg.LinAlgError:  # K numerically semidefinite close to the optimum
            ksolve = lambda b: np.linalg.lstsq(K, b, rcond=None)[0]  # noqa: E731

        def newton(rc):
            dx = ksolve(-rd - G.T @ ((-rc + z * rp) / s))
            ds = -rp - G @ dx
            dz = (-rc - z * ds) / s
            return dx, ds, dz

        def step_len(v, dv):
            neg = dv < 0
            return min(1.0, float(np.min(-v[neg] / dv[neg]))) if np.any(neg) else 1.0

        dx_a, ds_a, dz_a = newton(s * z)
        a_aff = min(step_len(s, ds_a), step_len(z, dz_a))
        mu_aff = float((s + a_aff * ds_a) @ (z + a_aff * dz_a)) / m
        sigma = (mu_aff / mu) ** 3
        dx, ds, dz = newton(s * z + ds_a * dz_a - sigma * mu)
        a = min(1.0, 0.99 * min(step_len(s, ds), step_len(z, dz)))
        x, s, z = x + a * dx, s + a * ds, z + a * dz
    info = {"iters": it, "status": status, "mu": mu, "objective": objective(c, x, nw, Hq, gq, const),
            "x": x.copy(), "z": z.copy(), "s": s.copy()}
    ub = h[n:2 * n]
    x = np.minimum(np.maximum(x, 0.0), ub)
    return x[:nw].reshape(P, N).copy(), x[nw:2 * nw].reshape(P, N).copy(), x[2 * nw:].copy(), info


def kkt_certificate(c: BiConsts, params, w_s, w_l, u_g, tau: float = 1e-2, multipliers: bool = True) -> dict:
    """Solver-independent optimality certificate of a PRIMAL point x = (w_s, w_l, u_g).  The multipliers are
    RECONSTRUCTED here (never taken from the solver under test) as

        z = argmin_{z >= 0} |grad f(x) + G' z|^2 + |diag(h - G x) z|^2 / tau^2      (Lawson-Hanson NNLS),

    i.e. the best dual vector for stationarity that pays for every multiplier on a row with slack, and
    the certificate reports

    * ``max_violation``  : largest constraint violation max(G x - h);
    * ``objective``      : f(x);
    * ``stationarity``   : |grad f(x) + G' z|_inf / max(1, |grad f(x)|_inf);
    * ``complementarity``: max_i z_i (h - G x)_i;
    * ``duality_gap``    : z' (h - G x) - with the stationarity residual r the convex program gives
                           f(x) - f* <= duality_gap + |r|_1 * diam(box).

    The optimum is unique in u_g and in every cumulative charge with omega_k > 0, so a point with small
    violation, stationarity and complementarity IS the optimum in those coordinates.  (Oracle solutions at
    tol 1e-9: stationarity <= 2e-7, complementarity <= 5e-9; a 1e-4 perturbation of one coordinate raises one
    of the two above 5e-5.)"""
    from scipy.optimize import nnls

    n, nw, G, h, Hq, gq, const = assemble(c, *[np.asarray(p, float) for p in params[:6]], float(params[6]),
                                          np.asarray(params[7], float))
    x = np.concatenate([np.asarray(w_s).ravel(), np.asarray(w_l).ravel(), np.asarray(u_g).ravel()])
    slack = h - G @ x
    pos = np.maximum(slack, 0.0)
    grad = Hq @ x + gq
    grad[2 * nw:] += 1.7 * c.c_g * np.maximum(x[2 * nw:], 0.0) ** 0.7
    # rows with a slack above tau cannot carry a multiplier worth its complementarity price: only the others are
    # candidates (keeps the NNLS at a few hundred columns for N = 24, P = 12)
    cand = np.nonzero(pos <= tau)[0]
    z = np.zeros(G.shape[0])
    if not multipliers:  # primal part only (objective, feasibility): the NNLS takes ~20 s at N = 24, P = 12
        return {"objective": objective(c, x, nw, Hq, gq, const), "max_violation": float(np.max(-slack)),
                "stationarity": 0.0, "complementarity": 0.0, "duality_gap": 0.0}
    if cand.size:
        A = np.vstack([G[cand].T, np.diag(pos[cand] / tau)])
        z[cand], _ = nnls(A, np.concatenate([-grad, np.zeros(cand.size)]), maxiter=50 * A.shape[1])
    resid = grad + G.T @ z
    return {"objective": objective(c, x, nw, Hq, gq, const), "max_violation": float(np.max(-slack)),
            "stationarity": float(np.max(np.abs(resid)) / max(1.0, float(np.max(np.abs(grad))))),
            "complementarity": float(np.max(z * pos)), "duality_gap": float(z @ pos)}
