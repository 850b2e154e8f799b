"""Golden fixtures of the BiMPC, the price loop and the one-station closed loop
(tests/golden/gen_loop_golden.py): the oracles reproduce them on the CPU, the CUDA path matches them on
the GPU (no oracle computation at GPU-test time)."""
import os

import numpy as np
import pytest

from oracle import bimpc_oracle as bo
from oracle import lompc_oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _bimpc_case(z, cost_type, N, P):
    key = f"c{cost_type}_N{N}_P{P}"
    ins = [z[f"{key}_{n}"] for n in ("Mp_s", "Mp_l", "beta_s", "beta_l", "gamma_sm", "gamma_lm", "x0", "demand")]
    return key, ins


# ----------------------------------------------------------------------------- CPU: oracle vs golden
def test_bimpc_oracle_reproduces_golden():
    z = np.load(os.path.join(GOLD, "bimpc_golden.npz"))
    c = bo.example_consts(16, 12)
    c.cost_type = bo.UNWEIGHTED
    key, ins = _bimpc_case(z, bo.UNWEIGHTED, 16, 12)
    for s in range(2):
        par = [a[s] for a in ins]
        ws, wl, ug, info = bo.solve_ipm(c, *par)
        assert info["iters"] == z[f"{key}_iters"][s]
        assert np.allclose(ug, z[f"{key}_u_g"][s], atol=1e-12) and np.allclose(ws, z[f"{key}_w_hat_s"][s], atol=1e-12)
        assert z[f"{key}_max_violation"][s] <= 1e-9


def test_price_loop_oracle_reproduces_golden():
    from oracle.price_oracle import PriceOracle
    z = np.load(os.path.join(GOLD, "price_loop_golden.npz"))
    o = orc.small_ev_consts()
    po = PriceOracle(12, o, "linear-convex")
    key = "small_linear-convex"
    for g in range(2):
        po.set_charge_levels(z[f"{key}_y0"][g])
        lm, st = po.compute_optimal_prices(z[f"{key}_w_ref"][g], 0.0)
        assert st["iter"] == z[f"{key}_iters"][g]
        assert np.allclose(lm, z[f"{key}_prices"][g], rtol=0, atol=1e-9)


# ----------------------------------------------------------------------------- GPU: CUDA path vs golden
@pytest.mark.gpu
@pytest.mark.parametrize("cost_type", [bo.WEIGHTED, bo.UNWEIGHTED, bo.EXP_UNWEIGHTED])
@pytest.mark.parametrize("N", [16, 24])
def test_bimpc_kernel_matches_golden(cost_type, N):
    from chargingstation.bimpc import BiMPC, BiMPCChargingCostType, BiMPCConstants
    from chargingstation.lompc import LoMPCConstants
    z = np.load(os.path.join(GOLD, "bimpc_golden.npz"))
    P = 12
    c = bo.example_consts(N, P)
    c.cost_type = cost_type
    key, ins = _bimpc_case(z, cost_type, N, P)
    b = BiMPC(N, P, BiMPCConstants(c.delta, c.c_g, c.u_g_max, c.u_b_max, c.x_max, BiMPCChargingCostType(cost_type), 5),
              LoMPCConstants(0.05, c.theta_s, 0.9, c.w_max_s, "small"), LoMPCConstants(0.025, c.theta_l, 0.9, c.w_max_l, "large"))
    ws, wl, ug, info = b.solve_bimpc_batch(*ins)
    assert (info["status"] == 0).all()
    for s in range(ins[0].shape[0]):
        par = [a[s] for a in ins]
        k = bo.kkt_certificate(c, par, ws[s], wl[s], ug[s])  # evaluates the point only (no solve)
        gold = z[f"{key}_objective"][s]
        assert k["max_violation"] <= 1e-8
        assert abs(k["objective"] - gold) <= 1e-7 * max(1.0, abs(gold))  # north-star bar: 1e-6
        assert np.max(np.abs(ug[s] - z[f"{key}_u_g"][s])) <= 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("ev", ["small", "large"])
@pytest.mark.parametrize("price_type", ["linear", "linear-convex"])
def test_price_loop_matches_golden(ev, price_type):
    """Three groups solved one after the other by ONE PriceSolver (the warm start chains through them)."""
    from chargingstation import settings
    from chargingstation.lompc import LoMPCConstants
    from chargingstation.price_solver import PriceSolver
    settings.PRINT_LEVEL = 0
    z = np.load(os.path.join(GOLD, "price_loop_golden.npz"))
    o = orc.small_ev_consts() if ev == "small" else orc.large_ev_consts()
    ps = PriceSolver(12, LoMPCConstants(o.delta, o.theta, o.y_max, o.w_max, o.ev_type), price_type)
    key = f"{ev}_{price_type}"
    for g in range(z[f"{key}_y0"].shape[0]):
        ps.set_charge_levels(z[f"{key}_y0"][g])
        lm, st = ps.compute_optimal_prices(z[f"{key}_w_ref"][g], 0.0)
        assert st["iter"] == z[f"{key}_iters"][g]
        scale = max(1.0, np.max(np.abs(z[f"{key}_prices"][g])))
        assert np.max(np.abs(lm - z[f"{key}_prices"][g])) <= 1e-7 * scale
        assert abs(st["price_after_reg"] - z[f"{key}_post"][g]) <= 1e-7 * max(1.0, abs(z[f"{key}_post"][g]))
        w0, p0 = ps.get_w0_price0(lm[: ps.r], 0.0)
        assert np.max(np.abs(w0 - z[f"{key}_w0"][g])) <= 1e-7 * o.w_max
        assert abs(p0 - z[f"{key}_price0"][g]) <= 1e-7 * max(1.0, abs(z[f"{key}_price0"][g]))


@pytest.mark.gpu
def test_station_closed_loop_matches_golden():
    from chargingstation import settings
    from chargingstation.bimpc import BiMPCChargingCostType, BiMPCConstants
    from chargingstation.charging_station import ChargingStation, ChargingStationConstants
    from chargingstation.demand_data import medium_term_demand_forecast
    from chargingstation.lompc import LoMPCConstants
    settings.PRINT_LEVEL = 0
    z = np.load(os.path.join(GOLD, "station_golden.npz"))
    Tf, N_bi, N_lo, M2, P = (int(v) for v in z["sizes"])
    dem = medium_term_demand_forecast(Tf + N_bi + 1, 0.25) * (M2 / 500)
    consts = ChargingStationConstants(Tf, N_bi, N_lo, M2, P, dem,
                                      BiMPCConstants(1e3, 1, 1, 0.3, 0.3, BiMPCChargingCostType.UNWEIGHTED, 5),
                                      LoMPCConstants(0.05, 10, 0.9, 0.25, "small"),
                                      LoMPCConstants(0.025, 50, 0.9, 0.15, "large"), "linear-convex")
    np.random.seed(0)
    cs = ChargingStation(consts)
    logs = cs.simulate()
    for t in range(Tf):
        assert np.array_equal(logs["statistics"]["Mp_s"][:, t], z["Mp_s"][t])
        assert np.array_equal(logs["statistics"]["niter_s"][:, t], z["niter_s"][t])
        assert np.array_equal(logs["statistics"]["niter_l"][:, t], z["niter_l"][t])
        assert abs(logs["inputs"]["u_g"][t] - z["u_g"][t][0]) <= 1e-5
        assert abs(logs["states"]["x"][t] - z["x_before"][t]) <= 1e-5
    assert np.max(np.abs(cs.y_s - z["y_s_final"])) <= 1e-5 and np.max(np.abs(cs.y_l - z["y_l_final"])) <= 1e-5
