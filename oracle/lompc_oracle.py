"""CPU ORACLE (test infrastructure, NOT a product path) for the lower-level MPC QP.

This module restates, in plain numpy, the arithmetic of the reference's
``chargingstation/lompc.py``.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and
only as the checker.  Nothing under ``incentive-design-mpc_b200/`` imports it.

PARITY UNPINNED.  The reference solves this QP with ``cvxpy`` (unpinned,
``environment.yml:9``) -> ``CLARABEL`` (``settings.py:11``, call site
``lompc.py:152``), a third-party Rust interior-point solver that is neither in
``/root/reference`` nor installable in this image, and the reference's own test
scripts hold no golden vectors or assertions (``test/test_lompc.py:30-40`` only
prints a timing).  What pins this oracle instead:

* the problem is strongly convex, so its optimum is unique;
* three independent solvers agree here: (i) ``solve_ipm`` - a restatement of
  the published algorithm class of Clarabel (primal-dual interior point with a
  Mehrotra predictor-corrector on the QP ``min 1/2 x'Px+q'x, Ax+s=b, s>=0`` in
  the canonical form cvxpy emits for ``lompc.py:73-135``, stopped at Clarabel's
  default 1e-8 tolerances); (ii) ``solve_active_set`` - an exact primal
  active-set method on the piecewise-quadratic form; (iii) scipy's BVLS
  (``tests/test_oracle.py``);
* ``kkt_certificate`` accepts or rejects ANY candidate ``w`` without reference
  to a solver, and turns the residual into a distance-to-optimum bound through
  the strong-convexity modulus.

Notation: ``A = tril(ones)`` (lompc.py:69),
``c = 2*delta*theta^2`` (lompc.py:71), ``q = 3*theta/(4*w_max)`` (lompc.py:67).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

# settings.py:7-9
MIN_MAX_BAT_SOC = 0.75
MAX_MAX_BAT_SOC = 0.9
MAX_BAT_CHARGE_RATE = 0.25

# lompc.py:109-115 -- pwl(x) = max(0, x-.125, 1.5x-.375, 2x-.75), x = w / w_max.
PWL_SLOPES = np.array([0.0, 1.0, 1.5, 2.0])
PWL_OFFSETS = np.array([0.0, 0.125, 0.375, 0.75])
PWL_BREAKS = np.array([0.125, 0.5, 0.75])  # where consecutive pieces cross


@dataclass
class OracleConsts:
    """Mirror of ``LoMPCConstants`` (lompc.py:12-26)."""

    delta: float
    theta: float
    y_max: float
    w_max: float
    ev_type: str


def small_ev_consts() -> OracleConsts:
    # example/real_time_price_control.py:26-31, test/test_lompc.py:16-20
    return OracleConsts(0.05, 10.0, 0.9, 0.25, "small")


def large_ev_consts() -> OracleConsts:
    # example/real_time_price_control.py:33-37, test/test_lompc.py:21-25
    return OracleConsts(0.025, 50.0, 0.9, 0.15, "large")


def check_consts(consts: OracleConsts) -> None:
    # lompc.py:36-38
    assert (consts.y_max >= MIN_MAX_BAT_SOC) and (consts.y_max <= MAX_MAX_BAT_SOC)
    assert (consts.w_max >= 0) and (consts.w_max <= MAX_BAT_CHARGE_RATE)
    assert consts.ev_type in ("small", "large")


# --------------------------------------------------------------------------
# Direct restatement of the cvxpy objective (lompc.py:101-135)
# --------------------------------------------------------------------------
def lompc_cost(N: int, consts: OracleConsts, w: np.ndarray, lmbd: np.ndarray,
               lmbd_r: float, gamma: float) -> float:
    """Objective value exactly as the cvxpy expression tree spells it."""
    th, wm, dl = consts.theta, consts.w_max, consts.delta
    q_scale = 3 * th / (4 * wm)  # lompc.py:67
    cost = 0.0
    if consts.ev_type == "small":
        cost += th ** 2 * np.sum((w / 0.9) ** 2)  # lompc.py:107
    else:
        w_rel = w / wm  # lompc.py:110-116
        pwl = np.sum(np.maximum.reduce(
            [0.0 * w_rel, w_rel - 0.125, 1.5 * w_rel - 0.375, 2 * w_rel - 0.75]))
        cost += (th * wm) ** 2 * pwl
    y = np.cumsum(w)  # A @ w, lompc.py:119
    cost += dl * th ** 2 * (np.sum(y ** 2) - 2 * gamma * np.sum(y))  # lompc.py:120-124
    l_price = th * (lmbd[:N] @ w + lmbd[N:2 * N] @ (wm - w))  # lompc.py:128-131
    q_price = q_scale * lmbd[2 * N:] @ (w * w)  # lompc.py:133
    r_price = lmbd_r * th ** 2 * np.sum(w * w)  # lompc.py:135
    return float(cost + l_price + q_price + r_price)


# --------------------------------------------------------------------------
# Problem data in "H, g, kappa0 + separable pwl" form
# --------------------------------------------------------------------------
def lompc_problem_data(N: int, consts: OracleConsts, lmbd: np.ndarray,
                       lmbd_r: float, gamma: float):
    """Returns d[N], c, g[N], kappa0, pwl_scale such that

        cost(w) = 1/2 w'(diag(d) + c A'A) w + g'w + kappa0
                  + pwl_scale * sum_k pwl(w_k / w_max)        (large EVs only)
    """
    th, wm, dl = consts.theta, consts.w_max, consts.delta
    q_scale = 3 * th / (4 * wm)
    c = 2 * dl * th ** 2
    d = 2 * (lmbd_r * th ** 2 + q_scale * lmbd[2 * N:3 * N])
    if consts.ev_type == "small":
        d = d + 2 * th ** 2 / 0.81
        pwl_scale = 0.0
    else:
        pwl_scale = (th * wm) ** 2
    g = th * (lmbd[:N] - lmbd[N:2 * N]) - c * gamma * np.arange(N, 0, -1)
    kappa0 = th * wm * np.sum(lmbd[N:2 * N])
    return d, c, g, kappa0, pwl_scale


def dense_hessian(N: int, d: np.ndarray, c: float) -> np.ndarray:
    A = np.tril(np.ones((N, N)))
    return np.diag(d) + c * A.T @ A


# --------------------------------------------------------------------------
# (i) Interior-point restatement of the cvxpy -> CLARABEL solve (lompc.py:152)
# --------------------------------------------------------------------------
def _canonical_qp(N, consts, lmbd, lmbd_r, gamma):
    """QP in the conic form cvxpy hands to Clarabel: min 1/2 x'Px + q'x,
    G x + s = h, s >= 0.  Small EV: x = w.  Large EV: x = [w; t] with the
    epigraph rows of ``cv.maximum`` (lompc.py:111-115)."""
    d, c, g, kappa0, pwl_scale = lompc_problem_data(N, consts, lmbd, lmbd_r, gamma)
    H = dense_hessian(N, d, c)
    I = np.eye(N)
    wm = consts.w_max
    if consts.ev_type == "small":
        P, q = H, g
        G = np.vstack([-I, I])  # w >= 0 (lompc.py:74 nonneg), w <= w_max (lompc.py:93)
        h = np.concatenate([np.zeros(N), wm * np.ones(N)])
    else:
        P = np.zeros((2 * N, 2 * N))
        P[:N, :N] = H
        q = np.concatenate([g, pwl_scale * np.ones(N)])
        rows, rhs = [np.hstack([-I, 0 * I]), np.hstack([I, 0 * I])], [np.zeros(N), wm * np.ones(N)]
        for a, o in zip(PWL_SLOPES, PWL_OFFSETS):
            rows.append(np.hstack([(a / wm) * I, -I]))  # a*w/w_max - o <= t
            rhs.append(o * np.ones(N))
        G, h = np.vstack(rows), np.concatenate(rhs)
    return P, q, G, h, kappa0


def solve_ipm(N: int, consts: OracleConsts, lmbd: np.ndarray, lmbd_r: float,
              gamma: float, tol: float = 1e-8, max_iter: int = 200):
    """Primal-dual Mehrotra predictor-corrector IPM; default ``tol`` is
    Clarabel's default gap/feasibility tolerance.  Returns (w, cost, iters)
    with ``cost`` evaluated like ``self.cost.value`` (lompc.py:155), i.e.
    INCLUDING the constant theta*w_max*sum(lmbd2)."""
    assert gamma <= consts.y_max  # lompc.py:87
    assert np.all(lmbd >= 0) and lmbd_r >= 0 and gamma >= 0  # nonneg Parameters, lompc.py:78-82
    P, q, G, h, _ = _canonical_qp(N, consts, lmbd, lmbd_r, gamma)
    n, m = P.shape[0], G.shape[0]
    x = np.zeros(n)
    if consts.ev_type == "small":
        x[:] = 0.5 * consts.w_max
    else:
        x[:N] = 0.5 * consts.w_max
        x[N:] = 1.0
    s = np.maximum(h - G @ x, 1e-2)
    z = np.ones(m)
    it = 0
    for it in range(max_iter):
        rd = P @ x + q + G.T @ z
        rp = G @ x + s - h
        mu = s @ z / m
        pobj = 0.5 * x @ P @ x + q @ x
        dobj = -0.5 * x @ P @ x - h @ z
        gap = abs(pobj - dobj)
        if (np.linalg.norm(rd, np.inf) <= tol * max(1.0, np.linalg.norm(q, np.inf))
                and np.linalg.norm(rp, np.inf) <= tol * max(1.0, np.linalg.norm(h, np.inf))
                and gap <= tol * max(1.0, min(abs(pobj), abs(dobj)))):
            break
        W = z / s
        K = P + G.T @ (W[:, None] * G)

        def newton(rc):
            # P dx + G'dz = -rd ; G dx + ds = -rp ; z ds + s dz = -rc
            # => ds = -rp - G dx, dz = (-rc + z rp + z G dx)/s, reduced system in dx
            rhs = -rd - G.T @ ((-rc + z * rp) / s)
            dx = np.linalg.solve(K, rhs)
            ds = -rp - G @ dx
            dz = (-rc - z * ds) / s
            return dx, ds, dz

        def step_len(v, dv):
            neg = dv < 0
            return min(1.0, float(np.min(-v[neg] / dv[neg]))) if np.any(neg) else 1.0

        dx_a, ds_a, dz_a = newton(s * z)
        a_aff = min(step_len(s, ds_a), step_len(z, dz_a))
        mu_aff = (s + a_aff * ds_a) @ (z + a_aff * dz_a) / m
        sigma = (mu_aff / mu) ** 3
        dx, ds, dz = newton(s * z + ds_a * dz_a - sigma * mu)
        a = 0.99 * min(step_len(s, ds), step_len(z, dz))
        a = min(a, 1.0)
        x, s, z = x + a * dx, s + a * ds, z + a * dz
    w = x[:N].copy()
    return w, lompc_cost(N, consts, w, lmbd, lmbd_r, gamma), it


# --------------------------------------------------------------------------
# (ii) Exact primal active-set method on the piecewise-quadratic form
# --------------------------------------------------------------------------
def _segments(consts: OracleConsts):
    """Breakpoints b[0..S] and the pwl slope (per unit w) on each segment."""
    wm = consts.w_max
    if consts.ev_type == "small":
        return np.array([0.0, wm]), np.array([0.0])
    scale = (consts.theta * wm) ** 2 / wm
    return np.concatenate([[0.0], PWL_BREAKS * wm, [wm]]), scale * PWL_SLOPES


def solve_active_set(N: int, consts: OracleConsts, lmbd: np.ndarray, lmbd_r: float,
                     gamma: float, max_iter: int = 10000):
    """Exact (to fp64 round-off) solution.  Each coordinate is either FREE in a
    segment of the pwl (linear term = that segment's slope) or FIXED at a
    breakpoint; a feasible descent active-set iteration with one change per
    step, dense solves.  Returns (w, cost, iters)."""
    assert gamma <= consts.y_max
    d, c, g, kappa0, _ = lompc_problem_data(N, consts, lmbd, lmbd_r, gamma)
    H = dense_hessian(N, d, c)
    brk, slope = _segments(consts)
    nseg = len(slope)
    w = np.zeros(N)
    seg = np.zeros(N, dtype=int)  # segment index when free
    fixed = np.ones(N, dtype=bool)  # start with everything fixed at w = 0
    at = np.zeros(N, dtype=int)  # breakpoint index when fixed
    it = 0
    for it in range(max_iter):
        free = ~fixed
        wt = w.copy()
        if np.any(free):
            rhs = -(g[free] + slope[seg[free]]) - H[np.ix_(free, fixed)] @ w[fixed]
            wt[free] = np.linalg.solve(H[np.ix_(free, free)], rhs)
        # ratio test against the segment ends of the free coordinates
        alpha, blk, blk_at = 1.0, -1, -1
        for k in np.flatnonzero(free):
            lo, hi = brk[seg[k]], brk[seg[k] + 1]
            dk = wt[k] - w[k]
            if dk < 0 and wt[k] < lo:
                a = (lo - w[k]) / dk
                if a < alpha:
                    alpha, blk, blk_at = a, k, seg[k]
            elif dk > 0 and wt[k] > hi:
                a = (hi - w[k]) / dk
                if a < alpha:
                    alpha, blk, blk_at = a, k, seg[k] + 1
        if blk >= 0:
            w = w + alpha * (wt - w)
            fixed[blk], at[blk] = True, blk_at
            w[blk] = brk[blk_at]
            continue
        w = wt
        # multipliers of the fixed coordinates: need -grad_k in [slope_left, slope_right]
        grad = H @ w + g
        worst, wk, wseg = 0.0, -1, -1
        for k in np.flatnonzero(fixed):
            i = at[k]
            s_lo = slope[i - 1] if i > 0 else -np.inf
            s_hi = slope[i] if i < nseg else np.inf
            if -grad[k] < s_lo and s_lo + grad[k] > worst:
                worst, wk, wseg = s_lo + grad[k], k, i - 1
            elif -grad[k] > s_hi and -grad[k] - s_hi > worst:
                worst, wk, wseg = -grad[k] - s_hi, k, i
        if wk < 0 or worst <= 1e-13 * max(1.0, np.max(np.abs(g))):
            break
        fixed[wk], seg[wk] = False, wseg
    return w, lompc_cost(N, consts, w, lmbd, lmbd_r, gamma), it


# --------------------------------------------------------------------------
# (iii) Solver-independent KKT certificate
# --------------------------------------------------------------------------
def kkt_certificate(N: int, consts: OracleConsts, w: np.ndarray, lmbd: np.ndarray,
                    lmbd_r: float, gamma: float, band: float = 1e-12):
    """Returns (violation_inf, dist_bound).

    ``violation_inf`` is the infinity-norm distance of ``-(Hw+g)`` from the
    subdifferential of ``pwl + box indicator`` at ``w`` (coordinates within
    ``band*w_max`` of a breakpoint are treated as sitting on it), joined with
    the box infeasibility.  ``dist_bound = ||viol||_2 / lambda_min(H) +
    sqrt(N)*band*w_max`` bounds ``||w - w*||_2`` by strong convexity.
    """
    d, c, g, _, _ = lompc_problem_data(N, consts, lmbd, lmbd_r, gamma)
    H = dense_hessian(N, d, c)
    brk, slope = _segments(consts)
    nseg = len(slope)
    tol = band * consts.w_max
    grad = H @ w + g
    viol = np.zeros(N)
    for k in range(N):
        near = np.flatnonzero(np.abs(w[k] - brk) <= tol)
        if len(near):
            i = near[0]
            s_lo = slope[i - 1] if i > 0 else -np.inf
            s_hi = slope[i] if i < nseg else np.inf
        else:
            j = int(np.searchsorted(brk, w[k]) - 1)
            j = min(max(j, 0), nseg - 1)
            s_lo = s_hi = slope[j]
        viol[k] = max(s_lo + grad[k], -grad[k] - s_hi, 0.0)
        viol[k] = max(viol[k], -w[k], w[k] - consts.w_max)
    lam_min = float(np.linalg.eigvalsh(H)[0])
    return float(np.max(viol)), float(np.linalg.norm(viol) / lam_min + np.sqrt(N) * tol)


# --------------------------------------------------------------------------
# Feature map (lompc.py:164-187)
# --------------------------------------------------------------------------
def phi(N: int, consts: OracleConsts, w: np.ndarray) -> np.ndarray:
    assert w.shape == (N,)  # lompc.py:173
    q_scale = 3 * consts.theta / (4 * consts.w_max)
    return np.hstack((consts.theta * w, consts.theta * (consts.w_max - w), q_scale * (w * w)))


def Dphi(N: int, consts: OracleConsts, w: np.ndarray) -> np.ndarray:
    assert w.shape == (N,)  # lompc.py:180
    q_scale = 3 * consts.theta / (4 * consts.w_max)
    return np.vstack((consts.theta * np.eye(N), -consts.theta * np.eye(N),
                      2 * q_scale * np.diag(w)))


def get_price0(N: int, consts: OracleConsts, w: np.ndarray, lmbd: np.ndarray,
               lmbd_r: float) -> float:
    # lompc.py:164-170
    q_scale = 3 * consts.theta / (4 * consts.w_max)
    return float(consts.theta * (w[0] * lmbd[0] + (consts.w_max - w[0]) * lmbd[N])
                 + q_scale * w[0] ** 2 * lmbd[2 * N]
                 + consts.theta ** 2 * w[0] ** 2 * lmbd_r)


def solve_lompc(N: int, consts: OracleConsts, lmbd: np.ndarray, lmbd_r: float,
                gamma: float):
    """Oracle answer for ``LoMPC.solve_lompc`` (lompc.py:137-156): the exact
    optimum.  (The reference returns Clarabel's ~1e-8-accurate iterate.)"""
    w, cost, _ = solve_active_set(N, consts, lmbd, lmbd_r, gamma)
    return w, cost
