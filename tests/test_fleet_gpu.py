"""The device-resident fleet loop (chargingstation.fleet) against S independent runs of the
single-station mirror (chargingstation.charging_station), which tests/test_station_gpu.py
pins to the CPU oracle; plus size-independent properties at a larger fleet."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _consts(Tf, N_bi, N_lo, M2, P, cost_type, price_type="linear-convex"):
    from chargingstation import settings
    from chargingstation.bimpc import BiMPCChargingCostType, BiMPCConstants
    from chargingstation.charging_station import ChargingStationConstants
    from chargingstation.demand_data import medium_term_demand_forecast
    from chargingstation.lompc import LoMPCConstants
    settings.PRINT_LEVEL = 0
    dem = medium_term_demand_forecast(Tf + N_bi + 1, 0.25) * (M2 / 500)
    cb = BiMPCConstants(1e3, 1, 1, 0.3, 0.3, BiMPCChargingCostType(cost_type), 5)
    cs = LoMPCConstants(0.05, 10, 0.9, 0.25, "small")
    cl = LoMPCConstants(0.025, 50, 0.9, 0.15, "large")
    return ChargingStationConstants(Tf, N_bi, N_lo, M2, P, dem, cb, cs, cl, price_type)


@pytest.mark.parametrize("price_type", ["linear-convex", "linear"])
@pytest.mark.parametrize("sizes", [(8, 4, 40, 6), (16, 12, 120, 12)])  # (N_bi, N_lo, M, P); N_lo = 12 runs the fused chain kernel
def test_fleet_reference_chain_equals_single_station_runs(price_type, sizes):
    from chargingstation.charging_station import ChargingStation
    from chargingstation.fleet import ChargingStationFleet
    Tf, S = 4, 3
    N_bi, N_lo, M, P = sizes
    consts = _consts(Tf, N_bi, N_lo, M, P, 1, price_type)
    scale = np.array([1.0, 0.97, 1.04])
    demand = scale[:, None] * consts.demand[None, :]
    fleet = ChargingStationFleet(consts, S, demand=demand, seed=100, rng="numpy", chain="reference")
    fleet.simulate()
    for s in range(S):
        import copy
        c = copy.copy(consts)
        c.demand = demand[s]
        np.random.seed(100 + s)
        one = ChargingStation(c)
        ref = one.simulate()
        got = fleet.station_logs(s)
        assert np.array_equal(got["statistics"]["Mp_s"], ref["statistics"]["Mp_s"])
        assert np.array_equal(got["statistics"]["Mp_l"], ref["statistics"]["Mp_l"])
        assert np.array_equal(got["statistics"]["niter_s"], ref["statistics"]["niter_s"])
        assert np.array_equal(got["statistics"]["niter_l"], ref["statistics"]["niter_l"])
        assert got["statistics"]["ncharged_s"] == ref["statistics"]["ncharged_s"]
        assert got["statistics"]["ncharged_l"] == ref["statistics"]["ncharged_l"]
        for grp, key in (("inputs", "u_g"), ("states", "x"), ("inputs", "w_hat_s"), ("inputs", "w_hat_l"),
                         ("inputs", "w_s"), ("inputs", "w_l"), ("bounds", "beta_s"), ("bounds", "beta_l"),
                         ("statistics", "gamma_sm"), ("statistics", "gamma_lm"), ("prices", "avg_price_s"),
                         ("prices", "avg_price_l")):
            assert np.max(np.abs(got[grp][key] - ref[grp][key]) / np.maximum(1.0, np.abs(ref[grp][key]))) <= 1e-8, (grp, key)
        for key in ("price_red_s", "price_red_l"):
            a, b = got["prices"][key], ref["prices"][key]
            assert np.array_equal(np.isnan(a), np.isnan(b))
            assert np.nanmax(np.abs(a - b) / np.maximum(1.0, np.abs(b))) <= 1e-8
        assert np.max(np.abs(fleet.y["s"][s].cpu().numpy() - one.y_s)) <= 1e-9
        assert abs(float(fleet.x[s]) - one.x) <= 1e-9


@pytest.mark.parametrize("chain", ["reference", "partition"])
def test_fleet_device_rng_properties(chain):
    """64 stations, device RNG: population is conserved, SoCs stay inside the partitions, the
    battery stays inside its (robustified) limits, every non-empty group converged, and the
    realised mean charge tracks the BiMPC plan within the robustness bound."""
    from chargingstation.fleet import ChargingStationFleet
    Tf, S, M, P = 5, 64, 100, 12
    consts = _consts(Tf, 16, 12, M, P, 2)
    rng = np.random.default_rng(4)
    demand = np.stack([np.roll(consts.demand, int(rng.integers(24))) * rng.uniform(0.9, 1.1) for _ in range(S)])
    fleet = ChargingStationFleet(consts, S, demand=demand, seed=7, rng="device", chain=chain)
    log = fleet.simulate()
    for k in ("s", "l"):
        Mp = log[f"Mp_{k}"].cpu().numpy()
        assert np.all(Mp.sum(axis=1) == M)
        niter = log[f"niter_{k}"].cpu().numpy()
        assert np.all(niter[Mp > 0] >= 0) and np.all(niter[Mp == 0] == -1)
        conv = (Mp > 0) & (niter < 999)  # a few groups may run into the reference's cap of 1000 iterations
        assert conv.sum() >= 0.98 * (Mp > 0).sum()
        err = np.abs(log[f"w_{k}"].cpu().numpy() - log[f"w_hat_{k}"].cpu().numpy())
        assert np.all(err[conv] <= log[f"beta_{k}"].cpu().numpy()[conv] + 1e-9)
        y = fleet.y[k].cpu().numpy()
        assert y.min() >= 0.3 and y.max() <= 0.95 * 0.9 + 1e-12
        red = log[f"price_red_{k}"].cpu().numpy()
        assert np.all(np.isnan(red[Mp == 0])) and np.all(red[Mp > 0] <= 1e-9)
    assert np.all(log["bimpc_status"].cpu().numpy() == 0)
    x = log["x"].cpu().numpy()
    assert x.min() >= -1e-9 and x.max() <= 0.3 + 1e-9
    # two runs with the same seed are bit-identical (deterministic reductions, counter-based RNG)
    again = ChargingStationFleet(consts, S, demand=demand, seed=7, rng="device", chain=chain)
    again.simulate()
    assert np.array_equal(again.y["s"].cpu().numpy(), fleet.y["s"].cpu().numpy())
    assert np.array_equal(again.x.cpu().numpy(), fleet.x.cpu().numpy())
