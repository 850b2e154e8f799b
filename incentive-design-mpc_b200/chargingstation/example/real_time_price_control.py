"""Mirror of the reference's ``chargingstation/example/real_time_price_control.py``
(BASELINE.json configs[0]): one station, 500 + 500 EVs, 12 partitions, LoMPC horizon 12,
BiMPC horizon 16, 49 hours, "linear-convex" prices, the bundled demand profile scaled by
1/4 (real_time_price_control.py:11-79).  Run from ``incentive-design-mpc_b200/``:

    python -m chargingstation.example.real_time_price_control [--steps T] [--out logs.pkl]
"""
from __future__ import annotations

import argparse
import pickle
import time

import numpy as np

from chargingstation.bimpc import BiMPCChargingCostType, BiMPCConstants
from chargingstation.charging_station import ChargingStation, ChargingStationConstants
from chargingstation.demand_data import medium_term_demand_forecast
from chargingstation.lompc import LoMPCConstants

# Scenario of the reference's example (real_time_price_control.py:11-79); the module-level names are kept.
SIMULATION_LENGTH = 49   # hours simulated
HORIZON_LOMPC, HORIZON_BIMPC = 12, 16
NUM_EVS_PER_EV_TYPE, NUM_PARTITIONS = 500, 12
PRICE_TYPE = "linear-convex"   # or "linear"
DEMAND_SCALE = 1 / 4           # the reference also lists 1 / 3

# EV classes: (delta, theta [kWh], y_max, w_max)
_EV = {"small": (0.05, 10, 0.9, 0.25), "large": (0.025, 50, 0.9, 0.15)}
# normalised station: charging-cost weight, generation cost, generation / battery-rate / battery-size limits
_STATION = dict(delta=1e3, c_g=1, u_g_max=1, u_b_max=0.3, x_max=0.3,
                charging_cost_type=BiMPCChargingCostType.EXP_UNWEIGHTED, exp_rate=5)


def _get_lompc_consts() -> tuple[LoMPCConstants, LoMPCConstants]:
    return tuple(LoMPCConstants(*_EV[kind], kind) for kind in ("small", "large"))


def _get_normalized_bimpc_consts() -> BiMPCConstants:
    return BiMPCConstants(**_STATION)


def _get_unnormalized_external_demand(simulation_length: int = SIMULATION_LENGTH) -> np.ndarray:
    hours = simulation_length + HORIZON_BIMPC + 1
    return medium_term_demand_forecast(hours, DEMAND_SCALE, interpolate=False)


def get_chargingstation_consts(simulation_length: int = SIMULATION_LENGTH) -> ChargingStationConstants:
    small, large = _get_lompc_consts()
    return ChargingStationConstants(
        simulation_length=simulation_length, horizon_bimpc=HORIZON_BIMPC, horizon_lompc=HORIZON_LOMPC,
        nEVs_per_EV_type=NUM_EVS_PER_EV_TYPE, npartitions=NUM_PARTITIONS,
        demand=_get_unnormalized_external_demand(simulation_length), bimpc_consts=_get_normalized_bimpc_consts(),
        small_EV_consts=small, large_EV_consts=large, price_type=PRICE_TYPE)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=SIMULATION_LENGTH)
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--out", default="real-time-price-control_logs_" + PRICE_TYPE + ".pkl")
    args = ap.parse_args()
    if args.seed is not None:
        np.random.seed(args.seed)
    cs = ChargingStation(get_chargingstation_consts(args.steps))
    t0 = time.time()
    logs = cs.simulate()
    print(f"{args.steps} closed-loop steps in {time.time() - t0:.2f} s")
    with open(args.out, "wb") as file:
        pickle.dump(logs, file)


if __name__ == "__main__":
    main()
