"""CPU tests of the oracle itself (oracle/lompc_oracle.py): three independent
solvers and the KKT certificate must agree, and the committed golden vectors
must reproduce.  The reference holds no golden vectors for this path
(SURVEY.md section 4), so the certificate is what pins parity."""
import os

import numpy as np
import pytest
from scipy.optimize import lsq_linear

from oracle import lompc_oracle as orc

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "lompc_golden.npz")


def _draw(rng, N, consts, mode):
    th = consts.theta
    if mode == 0:
        return th * rng.random(3 * N), 3 * N * consts.delta * rng.random(), consts.y_max * rng.random()
    if mode == 1:
        return 0.05 * th * rng.random(3 * N), 0.0, consts.y_max - (0.3 + 0.2 * rng.random())
    lm = np.zeros(3 * N)
    lm[:2 * N] = 0.05 * th * rng.random(2 * N)
    return lm, 0.0, consts.y_max * rng.random()


@pytest.mark.parametrize("ev", ["small", "large"])
@pytest.mark.parametrize("N", [12, 24])
def test_cost_identity_and_certificate(ev, N):
    """1/2 w'Hw + g'w + kappa0 + pwl == the cvxpy expression (lompc.py:101-135)."""
    consts = orc.small_ev_consts() if ev == "small" else orc.large_ev_consts()
    rng = np.random.default_rng(7)
    for trial in range(12):
        lm, lr, gam = _draw(rng, N, consts, trial % 3)
        w, cost, _ = orc.solve_active_set(N, consts, lm, lr, gam)
        d, c, g, k0, ps = orc.lompc_problem_data(N, consts, lm, lr, gam)
        H = orc.dense_hessian(N, d, c)
        x = w / consts.w_max
        pwl = ps * np.sum(np.maximum.reduce([0 * x, x - .125, 1.5 * x - .375, 2 * x - .75]))
        assert abs(0.5 * w @ H @ w + g @ w + k0 + pwl - cost) <= 1e-12 * max(1, abs(cost))
        viol, dist = orc.kkt_certificate(N, consts, w, lm, lr, gam)
        assert viol <= 1e-10 * max(1.0, np.max(np.abs(g)))
        assert dist <= 1e-9
        # a perturbed point must be rejected
        w_bad = np.clip(w + 1e-4 * consts.w_max * rng.standard_normal(N), 0, consts.w_max)
        viol_bad, _ = orc.kkt_certificate(N, consts, w_bad, lm, lr, gam)
        assert viol_bad > 1e-6


@pytest.mark.parametrize("ev", ["small", "large"])
def test_ipm_restatement_agrees_with_exact(ev):
    """The Clarabel-style IPM at its default 1e-8 tolerances lands within the
    north-star tolerance of the exact optimum (w <= 1e-4*w_max here, cost 1e-6)."""
    consts = orc.small_ev_consts() if ev == "small" else orc.large_ev_consts()
    rng = np.random.default_rng(11)
    N = 24
    for trial in range(9):
        lm, lr, gam = _draw(rng, N, consts, trial % 3)
        w, cost, _ = orc.solve_active_set(N, consts, lm, lr, gam)
        w2, cost2, _ = orc.solve_ipm(N, consts, lm, lr, gam)
        assert np.max(np.abs(w - w2)) <= 1e-4 * consts.w_max
        assert abs(cost - cost2) <= 1e-6 * max(1.0, abs(cost))
        w3, cost3, _ = orc.solve_ipm(N, consts, lm, lr, gam, tol=1e-12)
        assert np.max(np.abs(w - w3)) <= 1e-7 * consts.w_max


def test_small_ev_matches_bvls():
    consts = orc.small_ev_consts()
    rng = np.random.default_rng(13)
    N = 24
    for trial in range(9):
        lm, lr, gam = _draw(rng, N, consts, trial % 3)
        w, _, _ = orc.solve_active_set(N, consts, lm, lr, gam)
        d, c, g, _, _ = orc.lompc_problem_data(N, consts, lm, lr, gam)
        L = np.linalg.cholesky(orc.dense_hessian(N, d, c))
        res = lsq_linear(L.T, -np.linalg.solve(L, g), bounds=(0, consts.w_max), method="bvls", tol=1e-14)
        assert np.max(np.abs(res.x - w)) <= 1e-9 * consts.w_max


def test_golden_vectors_reproduce():
    z = np.load(GOLDEN)
    for ev, consts in (("small", orc.small_ev_consts()), ("large", orc.large_ev_consts())):
        for N in (12, 24):
            for mode in range(4):
                key = f"{ev}_N{N}_m{mode}"
                lm, lr, gam = z[key + "_lmbd"], z[key + "_lmbd_r"], z[key + "_gamma"]
                for b in range(0, lm.shape[0], 6):
                    w, cost, _ = orc.solve_active_set(N, consts, lm[b], lr[b], gam[b])
                    assert np.max(np.abs(w - z[key + "_w"][b])) <= 1e-12
                    assert abs(cost - z[key + "_cost"][b]) <= 1e-10 * max(1, abs(cost))
                assert z[key + "_kkt"].max() <= 1e-10 * 1e4


def test_unpriced_solution_shape():
    """test_lompc.py:43-58 plots the unpriced solution: charge early, never exceed gamma."""
    consts = orc.small_ev_consts()
    N = 12
    w, _, _ = orc.solve_active_set(N, consts, np.zeros(3 * N), 0.0, consts.y_max)
    assert np.all(w >= -1e-15) and np.all(w <= consts.w_max + 1e-15)
    assert np.all(np.diff(w) <= 1e-12)
    assert np.cumsum(w)[-1] <= consts.y_max + 1e-12


def test_feature_map():
    consts = orc.large_ev_consts()
    N = 6
    rng = np.random.default_rng(3)
    w = consts.w_max * rng.random(N)
    lm = rng.random(3 * N)
    ph = orc.phi(N, consts, w)
    J = orc.Dphi(N, consts, w)
    eps = 1e-7
    for k in range(N):
        e = np.zeros(N)
        e[k] = eps
        assert np.allclose((orc.phi(N, consts, w + e) - orc.phi(N, consts, w - e)) / (2 * eps), J[:, k], atol=1e-6)
    # price = lmbd @ phi(w) equals the three price terms of lompc.py:126-135 at lmbd_r = 0
    q = 3 * consts.theta / (4 * consts.w_max)
    price = consts.theta * (lm[:N] @ w + lm[N:2 * N] @ (consts.w_max - w)) + q * lm[2 * N:] @ (w * w)
    assert abs(lm @ ph - price) <= 1e-12 * abs(price)


@pytest.mark.parametrize("ev", ["small", "large"])
def test_c_oracle_matches_numpy_oracle(ev):
    """oracle/lompc_oracle.c (the timed CPU baseline) restates the same IPM: it must land
    within the IPM's own accuracy of the exact optimum and reproduce the cost expression."""
    from oracle import c_oracle
    consts = orc.small_ev_consts() if ev == "small" else orc.large_ev_consts()
    rng = np.random.default_rng(17)
    for N in (12, 24):
        B = 48
        lm = consts.theta * rng.random((B, 3 * N))
        lm[B // 2:] *= 0.05
        lr = 3 * N * consts.delta * rng.random(B)
        lr[B // 2:] = 0.0
        gam = consts.y_max * rng.random(B)
        w, cost, iters, used = c_oracle.solve_lompc_batch(N, consts, lm, lr, gam, nthreads=2)
        assert used == 2 and iters.min() > 0
        for b in range(0, B, 4):
            wo, co, _ = orc.solve_active_set(N, consts, lm[b], lr[b], gam[b])
            assert np.max(np.abs(w[b] - wo)) <= 2e-4 * consts.w_max
            assert abs(cost[b] - co) <= 1e-6 * max(1, abs(co))
            assert abs(cost[b] - orc.lompc_cost(N, consts, w[b], lm[b], lr[b], gam[b])) <= 1e-12 * max(1, abs(co))
        wt, ct, _, _ = c_oracle.solve_lompc_batch(N, consts, lm, lr, gam, tol=1e-11, nthreads=2)
        for b in range(0, B, 4):
            wo, co, _ = orc.solve_active_set(N, consts, lm[b], lr[b], gam[b])
            assert np.max(np.abs(wt[b] - wo)) <= 1e-6 * consts.w_max


@pytest.mark.parametrize("ev", ["small", "large"])
@pytest.mark.parametrize("N", [12, 24, 48])
def test_c_exact_oracle_is_the_python_exact_oracle(ev, N):
    """oracle/lompc_oracle.c::solve_exact_one is the C twin of solve_active_set (same steps, same tie rules): the
    full-size price-loop / closed-loop fixtures are generated with it (tests/golden/gen_fullsize_golden.py)."""
    from oracle import c_oracle
    consts = orc.small_ev_consts() if ev == "small" else orc.large_ev_consts()
    rng = np.random.default_rng(N)
    B = 45
    lm = np.zeros((B, 3 * N))
    lr = np.zeros(B)
    gam = consts.y_max * rng.random(B)
    lm[:15] = consts.theta * rng.random((15, 3 * N))
    lr[:15] = 3 * N * consts.delta * rng.random(15)
    lm[15:30] = 0.05 * consts.theta * rng.random((15, 3 * N)) * (rng.random((15, 3 * N)) < 0.5)
    lm[30:, :2 * N] = 0.05 * consts.theta * rng.random((15, 2 * N))
    w, cost, iters = c_oracle.solve_lompc_exact_batch(N, consts, lm, lr, gam)
    for b in range(B):
        wo, co, ito = orc.solve_active_set(N, consts, lm[b], lr[b], gam[b])
        assert np.max(np.abs(w[b] - wo)) <= 1e-14 * consts.w_max
        assert abs(cost[b] - co) <= 1e-12 * max(1.0, abs(co))
        assert iters[b] == ito
    # broadcast prices (the _get_w_err case, price_solver.py:203-204)
    w1, c1, _ = c_oracle.solve_lompc_exact_batch(N, consts, lm[3], lr[3], gam)
    w2, c2, _ = c_oracle.solve_lompc_exact_batch(N, consts, np.tile(lm[3], (B, 1)), np.full(B, lr[3]), gam)
    assert np.array_equal(w1, w2) and np.array_equal(c1, c2)


def test_fast_price_oracle_equals_python_price_oracle():
    """PriceOracle(fast=True) (QPs on the C exact oracle) walks the same loop as the pure-Python one."""
    from oracle.price_oracle import PriceOracle
    consts = orc.large_ev_consts()
    N = 12
    rng = np.random.default_rng(3)
    y0 = 0.3 + 0.05 * rng.random(9)
    w_ref = consts.w_max * rng.random(N) * 0.6
    a, b = PriceOracle(N, consts, "linear-convex"), PriceOracle(N, consts, "linear-convex", fast=True)
    for po in (a, b):
        po.set_charge_levels(y0)
    la, sa = a.compute_optimal_prices(w_ref, 0.0)
    lb, sb = b.compute_optimal_prices(w_ref, 0.0)
    assert sa["iter"] == sb["iter"] and np.max(np.abs(la - lb)) <= 1e-10 * max(1.0, np.max(np.abs(la)))
    wa, pa = a.get_w0_price0(la[: a.r], 0.0)
    wb, pb = b.get_w0_price0(lb[: b.r], 0.0)
    assert np.max(np.abs(wa - wb)) <= 1e-12 and abs(pa - pb) <= 1e-10 * max(1.0, abs(pa))
