"""CPU tests of the BiMPC path: the dense oracle against scipy, and the kernel's algorithm
(csrc/bimpc_solve.cuh compiled for the host, tests/hostsim) against the oracle."""
import numpy as np
import pytest
from scipy.optimize import Bounds, LinearConstraint, minimize

import hostsim
from bimpc_cases import draw_station, stack
from oracle import bimpc_oracle as bo


def test_oracle_matches_scipy_on_a_small_instance():
    """Independent solver (SLSQP on the same dense data): same objective and point."""
    c = bo.example_consts(4, 2)
    c.cost_type = bo.UNWEIGHTED
    rng = np.random.default_rng(0)
    par = draw_station(rng, c, empty=False)
    ws, wl, ug, info = bo.solve_ipm(c, *par, tol=1e-10)
    assert info["status"] == 0
    n, nw, G, h, Hq, gq, const = bo.assemble(c, *par)
    f = lambda x: bo.objective(c, x, nw, Hq, gq, const)  # noqa: E731

    def grad(x):
        g = Hq @ x + gq
        g[2 * nw:] += 1.7 * c.c_g * np.maximum(x[2 * nw:], 0.0) ** 0.7
        return g

    def hess(x):
        H = Hq.copy()
        iu = np.arange(2 * nw, n)
        H[iu, iu] += 1.19 * c.c_g * np.maximum(x[iu], 1e-12) ** (-0.3)
        return H

    ub = h[n:2 * n]
    x0 = 0.5 * ub
    res = minimize(f, x0, jac=grad, hess=hess, method="trust-constr",
                   constraints=[LinearConstraint(G[2 * n:], -np.inf, h[2 * n:])], bounds=Bounds(np.zeros(n), ub),
                   options={"gtol": 1e-10, "xtol": 1e-12, "maxiter": 3000, "barrier_tol": 1e-12})
    # trust-constr stops at a barrier parameter of ~1e-5: it bounds the optimum from above
    assert info["objective"] <= res.fun + 1e-9
    assert abs(res.fun - info["objective"]) <= 1e-5 * max(1.0, abs(res.fun))
    x_ipm = np.concatenate([ws.ravel(), wl.ravel(), ug])
    assert np.max(G @ res.x - h) <= 1e-9
    assert np.max(np.abs(res.x - x_ipm)) <= 2e-3


@pytest.mark.parametrize("cost_type", [bo.WEIGHTED, bo.UNWEIGHTED, bo.EXP_UNWEIGHTED])
@pytest.mark.parametrize("N,P", [(16, 12), (24, 12), (8, 3)])
def test_kernel_algorithm_matches_oracle(cost_type, N, P):
    c = bo.example_consts(N, P)
    c.cost_type = cost_type
    rng = np.random.default_rng(100 * cost_type + N)
    stations = [draw_station(rng, c) for _ in range(3)]
    ws, wl, ug, info = hostsim.bimpc_solve(c, *stack(stations), bo.stage_weights(c))
    assert (info["status"] == 0).all()
    for s, par in enumerate(stations):
        wso, wlo, ugo, io = bo.solve_ipm(c, *par)
        assert io["status"] == 0
        # same algorithm -> same iteration count; objective and the generation schedule are
        # pinned (strictly convex); the split of early charging between partitions is only
        # weakly determined under the exponential weights 5^(k-N+1) (curvature 1e-8)
        assert abs(int(info["iters"][s]) - io["iters"]) <= 3  # (the kernel leaves decoupled empty partitions out)
        k = bo.kkt_certificate(c, par, ws[s], wl[s], ug[s], multipliers=(s == 0 or N * P <= 100))
        assert k["max_violation"] <= 1e-8
        assert abs(k["objective"] - io["objective"]) <= 1e-7 * max(1.0, abs(io["objective"]))
        # solver-independent certificate (multipliers reconstructed by NNLS): the kernel's point is stationary
        # and complementary to 1e-6 / 1e-7 whatever the oracle says
        assert k["stationarity"] <= 1e-6 and k["complementarity"] <= 1e-7, k  # (0.0 where not computed)
        assert np.max(np.abs(ug[s] - ugo)) <= 2e-5
        # Trajectories.  UNWEIGHTED: every cumulative charge has unit curvature -> 1e-5 (north-star bar).
        # WEIGHTED: partition p enters with weight (theta Mp_p)^2 ~ 1e-5..1e-4 of the generation cost's curvature,
        # so two solves that agree to 1e-9 in the objective differ by ~2e-5 in w (the dense oracle, not the
        # kernel, is the less accurate one there: at tol 1e-11 its normal equations stall while the kernel's
        # certificate reaches 1e-10).  EXP_UNWEIGHTED: stage weights 5^(k-N+1) leave the split of early
        # charging between partitions with curvature ~1e-8 (any 1e-9-accurate solver, CLARABEL included).
        tol_w = {bo.UNWEIGHTED: 1e-5, bo.WEIGHTED: 1e-4, bo.EXP_UNWEIGHTED: 3e-2}[cost_type]
        # (with the WEIGHTED cost an empty partition has zero weight: its w is arbitrary - the oracle
        #  returns the analytic centre, the kernel leaves the block out and returns 0)
        ks = par[0] > 0 if cost_type == bo.WEIGHTED else np.ones(P, dtype=bool)
        kl = par[1] > 0 if cost_type == bo.WEIGHTED else np.ones(P, dtype=bool)
        assert np.max(np.abs(ws[s] - wso)[ks]) <= tol_w and np.max(np.abs(wl[s] - wlo)[kl]) <= tol_w


def test_kernel_algorithm_tight_tolerance_is_stable():
    """The block-tridiagonal solve keeps converging where the dense normal equations stall."""
    c = bo.example_consts(16, 12)
    rng = np.random.default_rng(7)
    par = draw_station(rng, c)
    ws, wl, ug, info = hostsim.bimpc_solve(c, *stack([par]), bo.stage_weights(c), tol=1e-12)
    assert info["status"][0] == 0 and info["iters"][0] <= 40


def test_certificate_separates_optimal_from_perturbed_points():
    """bimpc_oracle.kkt_certificate reconstructs multipliers by NNLS: the oracle's own solution passes
    (stationarity / complementarity at solver tolerance), a 1e-4 perturbation of one coordinate does not."""
    for cost_type in (bo.WEIGHTED, bo.UNWEIGHTED, bo.EXP_UNWEIGHTED):
        c = bo.example_consts(8, 3)
        c.cost_type = cost_type
        par = draw_station(np.random.default_rng(5 + cost_type), c)
        ws, wl, ug, info = bo.solve_ipm(c, *par)
        assert info["status"] == 0
        k = bo.kkt_certificate(c, par, ws, wl, ug)
        assert k["max_violation"] <= 1e-9 and k["stationarity"] <= 1e-6 and k["complementarity"] <= 1e-7, k
        ug2 = ug.copy()
        ug2[2] = min(ug2[2] + 1e-4, c.u_g_max)
        k2 = bo.kkt_certificate(c, par, ws, wl, ug2)
        assert max(k2["stationarity"], k2["complementarity"], k2["max_violation"]) >= 1e-5, k2
