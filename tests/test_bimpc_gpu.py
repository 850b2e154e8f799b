"""GPU parity tests of the batched BiMPC kernel (through the C ABI / the BiMPC mirror)
against the dense CPU oracle (oracle/bimpc_oracle.py)."""
import numpy as np
import pytest

from bimpc_cases import draw_station, stack
from oracle import bimpc_oracle as bo

pytestmark = pytest.mark.gpu


def _mirror(c: bo.BiConsts):
    from chargingstation.bimpc import BiMPC, BiMPCChargingCostType, BiMPCConstants
    from chargingstation.lompc import LoMPCConstants
    cb = BiMPCConstants(c.delta, c.c_g, c.u_g_max, c.u_b_max, c.x_max, BiMPCChargingCostType(c.cost_type), c.exp_rate)
    cs = LoMPCConstants(0.05, c.theta_s, 0.9, c.w_max_s, "small")
    cl = LoMPCConstants(0.025, c.theta_l, 0.9, c.w_max_l, "large")
    return BiMPC(c.N, c.P, cb, cs, cl)


@pytest.mark.parametrize("cost_type", [bo.WEIGHTED, bo.UNWEIGHTED, bo.EXP_UNWEIGHTED])
@pytest.mark.parametrize("N,P", [(16, 12), (24, 12), (8, 3)])
def test_batch_against_oracle(cost_type, N, P):
    c = bo.example_consts(N, P)
    c.cost_type = cost_type
    rng = np.random.default_rng(100 * cost_type + N)
    stations = [draw_station(rng, c) for _ in range(6)]
    ws, wl, ug, info = _mirror(c).solve_bimpc_batch(*stack(stations))
    assert (info["status"] == 0).all(), info
    for s, par in enumerate(stations):
        wso, wlo, ugo, io = bo.solve_ipm(c, *par)
        assert abs(int(info["iters"][s]) - io["iters"]) <= 3  # (the kernel leaves decoupled empty partitions out)
        k = bo.kkt_certificate(c, par, ws[s], wl[s], ug[s], multipliers=(s == 0 or N * P <= 100))
        assert k["max_violation"] <= 1e-8
        # north-star bar: objective <= 1e-6 relative
        assert abs(k["objective"] - io["objective"]) <= 1e-7 * max(1.0, abs(io["objective"]))
        assert abs(info["objective"][s] - k["objective"]) <= 1e-9 * max(1.0, abs(k["objective"]))
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


def test_scalar_api_shapes_and_asserts():
    from chargingstation.bimpc import BiMPCParameters
    c = bo.example_consts(16, 12)
    b = _mirror(c)
    par = draw_station(np.random.default_rng(3), c)
    ws, wl, ug = b.solve_bimpc(BiMPCParameters(*par))
    assert ws.shape == (12, 16) and wl.shape == (12, 16) and ug.shape == (16,)
    assert b.get_bat_input_mat().shape == (16, 16)
    assert np.all(ws >= 0) and np.all(ws <= c.w_max_s) and np.all(wl <= c.w_max_l) and np.all(ug <= c.u_g_max)
    bad = list(par)
    bad[7] = par[7][:-1]
    with pytest.raises(AssertionError):  # bimpc.py:283
        b.solve_bimpc(BiMPCParameters(*bad))


def test_large_batch_is_deterministic_and_feasible():
    """Fleet size: 1,024 stations in one launch; repeated rows give bit-identical results."""
    c = bo.example_consts(24, 12)
    rng = np.random.default_rng(11)
    base = [draw_station(rng, c) for _ in range(64)]
    stations = base * 16
    ws, wl, ug, info = _mirror(c).solve_bimpc_batch(*stack(stations))
    assert (info["status"] == 0).all()
    assert np.array_equal(ws[:64], ws[-64:]) and np.array_equal(ug[:64], ug[-64:])
    for s in range(0, 64, 8):
        assert bo.kkt_certificate(c, base[s], ws[s], wl[s], ug[s])["max_violation"] <= 1e-8


def test_infeasible_station_reports_status():
    c = bo.example_consts(16, 12)
    par = list(draw_station(np.random.default_rng(5), c))
    par[7] = par[7] * 10.0  # demand far above u_g_max + battery: no feasible point
    from chargingstation.bimpc import BiMPCParameters
    b = _mirror(c)
    with pytest.warns(RuntimeWarning, match="not solved"):
        ws, wl, ug, info = b.solve_bimpc_batch(*stack([par]))
    assert info["status"][0] != 0
    # the scalar call hands back what the reference does for an unsolved problem: the .value of unsolved cvxpy
    # variables, None (bimpc.py:288-291)
    with pytest.warns(RuntimeWarning):
        out = b.solve_bimpc(BiMPCParameters(*par))
    assert out == (None, None, None) and b.last_info["status"] != 0


@pytest.mark.parametrize("random_Mp", [False, True])
@pytest.mark.parametrize("random_gamma", [False, True])
@pytest.mark.parametrize("early_peak_demand", [False, True])
def test_reference_scenario(random_Mp, random_gamma, early_peak_demand):
    """The scenario of test/test_bimpc.py:45-106 (N = 24, P = 12, 500 + 500 EVs, u_g_max = x_max = 1.5,
    exponential weights, optional random EV distribution / targets / early demand peak).  The reference
    plots the plan against its limits (dashed lines of _plot_figure); here the limits are asserted and the
    objective is compared with the dense oracle."""
    from chargingstation.bimpc import BiMPC, BiMPCChargingCostType, BiMPCConstants, BiMPCParameters
    from chargingstation.demand_data import medium_term_demand_forecast
    from chargingstation.lompc import LoMPCConstants
    N, P, M_s, M_l = 24, 12, 500, 500
    cs, cl = LoMPCConstants(0.05, 10, 0.9, 0.25, "small"), LoMPCConstants(0.025, 50, 0.9, 0.15, "large")
    cb = BiMPCConstants(1e3, 1, 1.5, 0.3, 1.5, BiMPCChargingCostType.EXP_UNWEIGHTED, 5)
    rng = np.random.default_rng(4 * random_Mp + 2 * random_gamma + early_peak_demand)
    simplex = lambda: (lambda v: v / v.sum())(rng.random(P) + 1e-6)  # noqa: E731
    B = cs.theta * M_s + cl.theta * M_l
    Mp_s = M_s * simplex() / B if random_Mp else M_s * np.ones(P) / (P * B)
    Mp_l = M_l * simplex() / B if random_Mp else M_l * np.ones(P) / (P * B)
    beta = np.sqrt(N) * 0.3 / P * np.ones(P)
    gamma_sm = 0.6 * rng.random(P) if random_gamma else 0.6 * np.ones(P)
    gamma_lm = 0.6 * rng.random(P) if random_gamma else 0.6 * np.ones(P)
    if early_peak_demand:
        demand = (medium_term_demand_forecast(24 + N, 1 / 4) / B)[17:17 + N]
    else:
        demand = medium_term_demand_forecast(N, 1 / 4) / B
    par = (Mp_s, Mp_l, beta, beta.copy(), gamma_sm, gamma_lm, 0.0, demand)
    bimpc = BiMPC(N, P, cb, cs, cl)
    w_s, w_l, u_g = bimpc.solve_bimpc(BiMPCParameters(*par))
    assert bimpc.last_info["status"] == 0
    c = bo.BiConsts(N, P, cb.delta, cb.c_g, cb.u_g_max, cb.u_b_max, cb.x_max, bo.EXP_UNWEIGHTED, 5.0, cs.theta,
                    cl.theta, cs.w_max, cl.w_max)
    # the dashed limit lines of the reference's figure
    assert np.all(w_s >= 0) and np.all(w_s <= cs.w_max) and np.all(w_l >= 0) and np.all(w_l <= cl.w_max)
    assert np.all(u_g >= 0) and np.all(u_g <= cb.u_g_max)
    A = bimpc.get_bat_input_mat()
    x_hat = A @ (u_g - demand - cs.theta * Mp_s @ w_s - cl.theta * Mp_l @ w_l)
    d_err = cs.theta * Mp_s @ beta + cl.theta * Mp_l @ beta
    assert np.all(x_hat - d_err >= -1e-8) and np.all(x_hat + d_err <= cb.x_max + 1e-8)
    k = bo.kkt_certificate(c, par, w_s, w_l, u_g)
    assert k["max_violation"] <= 1e-8
    _, _, ugo, io = bo.solve_ipm(c, *par)
    assert io["status"] == 0
    assert abs(k["objective"] - io["objective"]) <= 1e-7 * max(1.0, abs(io["objective"]))
    assert np.max(np.abs(u_g - ugo)) <= 2e-5
