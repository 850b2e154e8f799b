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
        k = bo.kkt_certificate(c, par, ws[s], wl[s], ug[s])
        assert k["max_violation"] <= 1e-8
        # north-star bar: objective <= 1e-6 relative
        assert abs(k["objective"] - io["objective"]) <= 1e-7 * max(1.0, abs(io["objective"]))
        assert abs(info["objective"][s] - k["objective"]) <= 1e-9 * max(1.0, abs(k["objective"]))
        assert np.max(np.abs(ug[s] - ugo)) <= 2e-5
        tol_w = 3e-2 if cost_type == bo.EXP_UNWEIGHTED else 1e-4
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
    ws, wl, ug, info = _mirror(c).solve_bimpc_batch(*stack([par]))
    assert info["status"][0] != 0
