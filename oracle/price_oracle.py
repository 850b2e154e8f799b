"""CPU ORACLE (test infrastructure, NOT a product path) for the price loop.

Restates ``chargingstation/price_solver.py`` and ``price_regularizer.py`` of the
reference in numpy/scipy, on top of the exact LoMPC oracle
(``oracle/lompc_oracle.py``).  PARITY UNPINNED for the same reason as there:
the reference's arithmetic lives in cvxpy -> CLARABEL (``price_solver.py:40,241``)
and cvxpy's default LP solver (``price_regularizer.py:45,83``), none of which is
installable here, and ``test/test_price_solver.py`` / ``test_price_regularizer.py``
print but never assert.  Pinning used instead:

* the non-negative QP of ``_price_gradient_descent_step`` is strongly convex
  (``P >= eps_reg * I``, price_solver.py:229-231) -> unique optimum; solved here
  exactly with Lawson-Hanson NNLS on the same Cholesky form the reference poses
  (``price_solver.py:265,270``) and certified by ``nnqp_kkt``;
* the regulariser LP is solved with HiGHS and compared with the closed form
  (separable per time step) on objective / feasibility /
  complementarity - the invariants ``test_price_regularizer.py:13-18`` prints.
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import linprog, nnls

from oracle import lompc_oracle as orc

# settings.py:13-19
MAX_PRICE_SOLVER_ITERATIONS = 1000
PRICE_SOLVER_TOL_TYPE = "avg"
PRICE_SOLVER_EPS_REG = 0.01
PRICE_SOLVER_EPS_TOL = 0.01


# --------------------------------------------------------------------------
# price_regularizer.py:68-85
# --------------------------------------------------------------------------
def solve_price_regularization_lp(A: np.ndarray, b: np.ndarray, c: np.ndarray) -> np.ndarray:
    """min c'x s.t. Ax = b, x >= 0 with HiGHS (the reference uses cvxpy's default solver)."""
    res = linprog(c, A_eq=A, b_eq=b, bounds=(0, None), method="highs")
    assert res.status == 0, res.message
    return res.x


def regularize_closed_form(N: int, consts: orc.OracleConsts, r: int, w: np.ndarray,
                           lmbd: np.ndarray) -> np.ndarray:
    """Closed form of the LP when A = Dphi(w)', b = A lmbd, c = phi(w)
    (price_solver.py:248-255): row k of A touches only x_k, x_{N+k}, x_{2N+k}.
    With b_k = theta(l1_k - l2_k) + 2 q w_k l3_k:  b_k < 0 -> x2_k = -b_k/theta;
    b_k >= 0 and r = 3N and w_k > 0 -> x3_k = b_k/(2 q w_k) (unit cost w_k/2 beats
    x1's w_k); otherwise x1_k = b_k/theta."""
    th = consts.theta
    q = 3 * th / (4 * consts.w_max)
    l3 = lmbd[2 * N:3 * N] if r == 3 * N else np.zeros(N)
    bk = th * (lmbd[:N] - lmbd[N:2 * N]) + 2 * q * w * l3
    x = np.zeros(r)
    for k in range(N):
        if bk[k] < 0:
            x[N + k] = -bk[k] / th
        elif r == 3 * N and w[k] > 1e-9 * consts.w_max:  # w_k of rounding-error size counts as 0 (degenerate LP)
            x[2 * N + k] = bk[k] / (2 * q * w[k])
        else:
            x[k] = bk[k] / th
    return x


# --------------------------------------------------------------------------
# price_solver.py:216-246 (+ the cvxpy NNQP :257-270)
# --------------------------------------------------------------------------
def nnqp_exact(P: np.ndarray, q: np.ndarray) -> np.ndarray:
    """argmin_{x>=0} x'Px + q'x via NNLS on R x ~ b, P = R'R (price_solver.py:265,270)."""
    R = np.linalg.cholesky(P).T
    b = -0.5 * np.linalg.solve(R.T, q)
    x, _ = nnls(R, b, maxiter=50 * P.shape[0])
    return x


def nnqp_kkt(P: np.ndarray, q: np.ndarray, x: np.ndarray) -> float:
    """KKT residual of the NNQP: grad = 2Px+q, need grad >= 0, x >= 0, x*grad = 0."""
    g = 2 * P @ x + q
    return float(max(np.max(np.maximum(-g, 0)), np.max(np.maximum(-x, 0)),
                     np.max(np.abs(np.where(x > 1e-12 * max(1.0, np.max(x)), g, 0.0)))))


class PriceOracle:
    """Mirror of ``PriceSolver`` (price_solver.py:16-285) on the oracle solvers."""

    def __init__(self, N: int, consts: orc.OracleConsts, price_type: str, fast: bool = False) -> None:
        """``fast``: solve the LoMPC QPs with the C twin of the exact active-set oracle
        (``c_oracle.solve_lompc_exact_batch``, bit-compatible with ``lompc_oracle.solve_active_set`` to 1e-15,
        ~40x faster) so that the reference's full sizes (500 + 500 EVs, N = 24) are affordable."""
        assert price_type in ("linear", "linear-convex")  # price_solver.py:24
        self.fast = fast
        orc.check_consts(consts)
        self.N = N
        self.r = 2 * N if price_type == "linear" else 3 * N  # :44-47
        self.consts = consts
        self.price_type = price_type
        self.prev_prices = np.zeros(self.r)  # :56
        self.A = np.tril(np.ones((N, N)))  # :58
        self.eps_reg = PRICE_SOLVER_EPS_REG
        self.eps_tol = PRICE_SOLVER_EPS_TOL
        self.m = 2 * consts.delta * consts.theta ** 2  # :64, lompc.py:71
        self.lompc_solves = 0

    # -- lompc plumbing
    def _solve(self, lmbd, lmbd_r, gamma):
        self.lompc_solves += 1
        if self.fast:
            from oracle import c_oracle
            w, cost, _ = c_oracle.solve_lompc_exact_batch(self.N, self.consts, lmbd, lmbd_r, np.array([gamma]), nthreads=1)
            return w[0], float(cost[0])
        return orc.solve_lompc(self.N, self.consts, lmbd, lmbd_r, gamma)

    def _solve_evs(self, lmbd, lmbd_r, gamma):
        """w_i for every EV of the group (rows), in EV order."""
        self.lompc_solves += len(gamma)
        if self.fast:
            from oracle import c_oracle
            return c_oracle.solve_lompc_exact_batch(self.N, self.consts, lmbd, lmbd_r, gamma)[0]
        return np.array([orc.solve_lompc(self.N, self.consts, lmbd, lmbd_r, g)[0] for g in gamma]).reshape(-1, self.N)

    def phi(self, w):
        return orc.phi(self.N, self.consts, w)

    def Dphi(self, w):
        return orc.Dphi(self.N, self.consts, w)

    # -- price_solver.py:66-77
    def set_charge_levels(self, y0: np.ndarray) -> None:
        assert all(y0 >= 0) and all(y0 <= self.consts.y_max)
        assert len(y0.shape) == 1
        self.nEVs = len(y0)
        self.y0 = y0
        self.y0_rng = (np.max(y0) - np.min(y0)) / 2
        self.gamma_sc = self.consts.y_max - (np.max(y0) + np.min(y0)) / 2
        self.gamma_sm = self.consts.y_max - np.mean(y0)

    # -- price_solver.py:182-194
    def get_robustness_bounds(self, lmbd_r: float):
        kappa = lmbd_r / self.consts.delta + 1e-5
        w_err_bound = np.sqrt(self.N) * self.y0_rng + self.eps_tol
        w0_err_bound = w_err_bound * np.min((1, 1 / np.sqrt(kappa)))
        return w_err_bound, w0_err_bound

    def _metric(self, lmbd_r: float):
        kappa = lmbd_r / self.consts.delta
        A_bar = self.A.T @ self.A + kappa * np.eye(self.N)
        return A_bar, np.linalg.inv(A_bar)

    # -- price_solver.py:196-214
    def get_w_err(self, lmbd, lmbd_r, w_ref, A_bar):
        w_avg = np.zeros(self.N)
        w_err_max = 0.0
        gamma = self.consts.y_max - self.y0
        w_all = self._solve_evs(lmbd, lmbd_r, gamma)
        for i in range(self.nEVs):
            w_i = w_all[i]
            w_avg += w_i
            w_err_i = np.sqrt((w_i - w_ref) @ A_bar @ (w_i - w_ref))
            w_err_max = max(w_err_max, w_err_i)
        w_avg = w_avg / self.nEVs
        w_avg_err = np.sqrt((w_avg - w_ref) @ A_bar @ (w_avg - w_ref))
        w0_err = np.abs(w_avg[0] - w_ref[0])
        return w_err_max, w0_err, w_avg_err, w_avg

    # -- price_solver.py:216-246
    def price_step_matrices(self, A_bar_inv, w_ref, w, lmbd):
        r = self.r
        phi_ref = self.phi(w_ref)[:r]
        phi = self.phi(w)[:r]
        Dphi = self.Dphi(w)[:r, :]
        P = 1 / (2 * self.m) * Dphi @ A_bar_inv @ Dphi.T + self.eps_reg * np.eye(r)
        q = -2 * P @ lmbd - (phi - phi_ref)
        return P, q

    def price_gradient_descent_step(self, A_bar_inv, w_ref, w, lmbd):
        P, q = self.price_step_matrices(A_bar_inv, w_ref, w, lmbd)
        dual_cost = lmbd @ P @ lmbd + q @ lmbd
        lmbd_next = nnqp_exact(P, q)
        dual_cost_new = lmbd_next @ P @ lmbd_next + q @ lmbd_next
        return lmbd_next, dual_cost - dual_cost_new

    # -- price_solver.py:248-255
    def regularize_prices(self, w, lmbd, closed_form=False):
        if closed_form:
            return regularize_closed_form(self.N, self.consts, self.r, w, lmbd)
        phi = self.phi(w)[:self.r]
        Dphi = self.Dphi(w)[:self.r, :]
        return solve_price_regularization_lp(Dphi.T, Dphi.T @ lmbd, phi)

    # -- price_solver.py:79-174
    def compute_optimal_prices(self, w_ref, lmbd_r, closed_form_reg=True, max_iter=None, trace=None):
        N, r = self.N, self.r
        tol, _ = self.get_robustness_bounds(lmbd_r)
        A_bar, A_bar_inv = self._metric(lmbd_r)
        lmbd_k, lmbd_k_new = np.zeros(3 * N), np.zeros(3 * N)
        lmbd_k[:r] = self.prev_prices
        phi_w_ref = self.phi(w_ref)
        w_k, dual_cost = self._solve(lmbd_k, lmbd_r, self.gamma_sc)
        dec_ac, dec_pred = [], []
        it = 0
        for it in range(MAX_PRICE_SOLVER_ITERATIONS if max_iter is None else max_iter):
            w_err_max, _, w_avg_err, w_avg = self.get_w_err(lmbd_k, lmbd_r, w_ref, A_bar)
            w_err = w_err_max if PRICE_SOLVER_TOL_TYPE == "max" else w_avg_err
            if trace is not None:
                trace.append({"lmbd": lmbd_k.copy(), "w_k": w_k.copy(), "w_avg": w_avg.copy(),
                              "w_avg_err": w_avg_err, "w_err_max": w_err_max})
            if w_err <= tol:
                break
            lmbd_k_new[:r], dec = self.price_gradient_descent_step(A_bar_inv, w_ref, w_k, lmbd_k[:r])
            w_k, dual_cost_new = self._solve(lmbd_k_new, lmbd_r, self.gamma_sc)
            dec_ac.append(dual_cost_new - dual_cost + (lmbd_k - lmbd_k_new) @ phi_w_ref)  # :135-137
            dec_pred.append(dec)
            dual_cost = dual_cost_new
            lmbd_k = lmbd_k_new  # :140 (aliases the two arrays from now on)
        price_pre = self.phi(w_k) @ lmbd_k
        lmbd_k[:r] = self.regularize_prices(w_k, lmbd_k[:r], closed_form=closed_form_reg)
        price_new = self.phi(w_k) @ lmbd_k
        self.prev_prices = lmbd_k[:r]
        stats = {"iter": it, "price_before_reg": price_pre, "price_after_reg": price_new,
                 "dual_cost_decrease_actual": np.array(dec_ac),
                 "dual_cost_decrease_predicted": np.array(dec_pred)}
        return lmbd_k, stats

    # -- price_solver.py:272-285
    def get_w0_price0(self, lmbd, lmbd_r):
        lmbd_ = np.zeros(3 * self.N)
        lmbd_[:self.r] = lmbd
        w0 = np.zeros(self.nEVs)
        price0 = 0.0
        gamma = self.consts.y_max - self.y0
        w_all = self._solve_evs(lmbd_, lmbd_r, gamma)
        for i in range(self.nEVs):
            w_i = w_all[i]
            w0[i] = w_i[0]
            price0 += orc.get_price0(self.N, self.consts, w_i, lmbd_, lmbd_r)
        return w0, price0 / self.nEVs
