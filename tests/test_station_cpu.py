"""CPU tests of the host-side pieces of the closed loop (no device needed)."""
import numpy as np

from chargingstation.charging_station import assign_partitions, partition_edges
from chargingstation.demand_data import MEDIUM_TERM_LOAD_FORECAST_MW, medium_term_demand_forecast


def test_demand_forecast_semantics():
    """demand_data.py:21-37: the forecast is a mid-hour value; the half-hourly series puts it on the odd
    half hours and the mean of neighbouring hours (wrapping around midnight) on the full hours; the hourly
    series (interpolate=False) is every other sample starting at 00:00, i.e. those means."""
    f24 = np.asarray(MEDIUM_TERM_LOAD_FORECAST_MW, dtype=float)
    d = medium_term_demand_forecast(66, 0.25)
    assert d.shape == (66,)
    assert np.allclose(d[:24], 0.25 * (f24 + np.roll(f24, 1)) / 2, rtol=1e-15)
    assert np.array_equal(d[24:48], d[:24]) and np.array_equal(d[48:], d[:18])
    di = medium_term_demand_forecast(30, 1 / 3, interpolate=True)
    assert di.shape == (60,)
    assert np.allclose(di[1::2][:24], f24 / 3) and np.isclose(di[0], (f24[0] + f24[-1]) / 2 / 3)
    assert np.isclose(di[2], (f24[1] + f24[0]) / 2 / 3)


def test_partition_assignment_semantics():
    """charging_station.py:111-116: closed intervals, the LAST matching partition wins, an SoC outside
    every partition keeps its previous index."""
    edges = partition_edges(0.3, 0.9, 12)
    assert np.array_equal(edges, np.linspace(0.3, 0.9, 13))
    y = np.array([0.3, 0.34999, edges[1], 0.9, 0.95, 0.2, edges[5]])
    idx = np.full(y.shape, 7)
    assign_partitions(y, edges, idx)
    assert idx[0] == 0 and idx[1] == 0
    assert idx[2] == 1          # on a boundary: the upper partition wins
    assert idx[3] == 11         # y_max belongs to the last partition
    assert idx[4] == 7 and idx[5] == 7  # outside [0.3, 0.9]: unchanged
    assert idx[6] == 5


def test_station_constants_asserts():
    """charging_station.py:44-53."""
    import pytest
    from chargingstation.bimpc import BiMPCChargingCostType, BiMPCConstants
    from chargingstation.charging_station import ChargingStation, ChargingStationConstants
    from chargingstation.lompc import LoMPCConstants
    cb = BiMPCConstants(1e3, 1, 1, 0.3, 0.3, BiMPCChargingCostType.EXP_UNWEIGHTED, 5)
    cs = LoMPCConstants(0.05, 10, 0.9, 0.25, "small")
    cl = LoMPCConstants(0.025, 50, 0.9, 0.15, "large")
    dem = medium_term_demand_forecast(20, 0.25)
    with pytest.raises(AssertionError):  # horizon_bimpc < horizon_lompc
        ChargingStation(ChargingStationConstants(2, 4, 8, 10, 3, dem, cb, cs, cl, "linear"))
    with pytest.raises(AssertionError):  # demand too short
        ChargingStation(ChargingStationConstants(30, 16, 12, 10, 3, dem, cb, cs, cl, "linear"))
