"""Drop-in mirror of the reference's ``chargingstation/charging_station.py``: the closed
loop of one charging station (BiMPC -> price loop per partition -> EV response -> state
update), same constants dataclass, attribute names, ``simulate()`` and ``logs`` schema
(charging_station.py:16-433).  Every optimisation it triggers runs on the GPU: the BiMPC
through ``include/bimpc_b200.h``, the price loops and the EV responses through
``include/lompc_b200.h``.  Random draws use ``np.random`` in the reference's order
(initial SoCs small then large, charging_station.py:95-100; replacements small then
large, :333-346), so a seeded run visits the same EV population.

The fleet version (thousands of stations, device-resident state) is
``chargingstation.fleet.ChargingStationFleet``."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from chargingstation import settings
from chargingstation.bimpc import BiMPC, BiMPCConstants, BiMPCParameters
from chargingstation.lompc import LoMPCConstants
from chargingstation.price_solver import PriceSolver
from chargingstation.settings import (ADD_RESIDUAL_CHARGE_TO_BATTERY, MAX_INITIAL_SOC,
                                      MIN_FULL_CHARGE_FRACTION, MIN_INITIAL_SOC)


@dataclass
class ChargingStationConstants:
    """Scenario of one station (field names and order as in the reference, charging_station.py:16-41).

    simulation_length   closed-loop steps (hours)
    horizon_bimpc       planning horizon of the upper level, >= horizon_lompc
    horizon_lompc       horizon of every EV's own problem
    nEVs_per_EV_type    EVs of each class (small, large) plugged in at any time
    npartitions         SoC classes per EV type (one price vector each)
    demand              external demand [kWh per step], at least simulation_length + horizon_bimpc + 1 values
    bimpc_consts        BiMPCConstants, normalised by the total EV capacity
    small_EV_consts / large_EV_consts   LoMPCConstants of the two classes
    price_type          "linear" or "linear-convex"
    """

    simulation_length: int
    horizon_bimpc: int
    horizon_lompc: int
    nEVs_per_EV_type: int
    npartitions: int
    demand: np.ndarray
    bimpc_consts: BiMPCConstants
    small_EV_consts: LoMPCConstants
    large_EV_consts: LoMPCConstants
    price_type: str


def partition_edges(y0_min: float, y_max: float, P: int) -> np.ndarray:
    """Partition definition of charging_station.py:87-92."""
    return np.linspace(y0_min, y_max, P + 1)


def assign_partitions(y: np.ndarray, edges: np.ndarray, idx: np.ndarray) -> None:
    """charging_station.py:111-116: the LAST partition p with edges[p] <= y <= edges[p+1]
    wins; an SoC outside every partition keeps its previous index."""
    P = len(edges) - 1
    for p in range(P):
        idx[(y >= edges[p]) & (y <= edges[p + 1])] = p


class ChargingStation:
    def __init__(self, consts: ChargingStationConstants, device: int = 0) -> None:
        # charging_station.py:44-53
        assert consts.simulation_length >= 1
        assert (consts.horizon_bimpc >= consts.horizon_lompc) and (consts.horizon_lompc >= 1)
        assert consts.nEVs_per_EV_type >= 1
        assert consts.npartitions >= 1
        assert (len(consts.demand.shape) == 1) and (
            consts.demand.shape[0] >= consts.simulation_length + consts.horizon_bimpc + 1)
        self._set_constants(consts)
        self.bimpc = BiMPC(self.N_bi, self.P, self.consts_bi, self.consts_s, self.consts_l, device=device)
        self.price_solver_s = PriceSolver(self.N_lo, self.consts_s, self.price_type, device=device)
        self.price_solver_l = PriceSolver(self.N_lo, self.consts_l, self.price_type, device=device)
        # State variables = (EV SoCs, charge stored).
        self._init_states()
        self._init_logs(consts)

    def _set_constants(self, consts: ChargingStationConstants) -> None:
        self.Tf = consts.simulation_length
        self.N_bi = consts.horizon_bimpc
        self.N_lo = consts.horizon_lompc
        self.M_2 = consts.nEVs_per_EV_type
        self.P = consts.npartitions
        self.demand = consts.demand
        self.consts_bi = consts.bimpc_consts
        self.consts_s = consts.small_EV_consts
        self.consts_l = consts.large_EV_consts
        self.price_type = consts.price_type
        self.r = 2 * self.N_lo if self.price_type == "linear" else 3 * self.N_lo
        # Range of initial SoCs of EVs.
        self.y0_min = MIN_INITIAL_SOC
        self.y0_max = MAX_INITIAL_SOC
        self.y0_s_rng = partition_edges(self.y0_min, self.consts_s.y_max, self.P)
        self.y0_l_rng = partition_edges(self.y0_min, self.consts_l.y_max, self.P)
        # Total charge capacity of EVs.
        self.B = (self.consts_s.theta + self.consts_l.theta) * self.M_2

    def _draw_soc(self, n: int) -> np.ndarray:
        return self.y0_min + (self.y0_max - self.y0_min) * np.random.random((n,))

    def _init_states(self) -> None:
        self.y_s = self._draw_soc(self.M_2)
        self.y_l = self._draw_soc(self.M_2)
        self.x = 0  # Storage battery SoC, normalized wrt B.
        self.t = 0
        self.ncharged_s = 0
        self.ncharged_l = 0
        self.idx_s = np.zeros((self.M_2,), dtype=int)
        self.idx_l = np.zeros((self.M_2,), dtype=int)
        self._update_indices()

    def _update_indices(self) -> None:
        assign_partitions(self.y_s, self.y0_s_rng, self.idx_s)
        assign_partitions(self.y_l, self.y0_l_rng, self.idx_l)

    def _init_logs(self, consts: ChargingStationConstants) -> None:
        P, Tf = self.P, self.Tf
        z = lambda *shape, **kw: np.zeros(shape, **kw)  # noqa: E731
        self.logs = {
            "constants": consts,
            "inputs": {"w_s": z(P, Tf), "w_l": z(P, Tf), "w_hat_s": z(P, Tf), "w_hat_l": z(P, Tf), "u_g": z(Tf)},
            "states": {"x": z(Tf)},
            "bounds": {"beta_s": z(P, Tf), "beta_l": z(P, Tf)},
            "statistics": {"ncharged_s": 0, "ncharged_l": 0, "gamma_sm": z(P, Tf), "gamma_lm": z(P, Tf),
                           "niter_s": z(P, Tf, dtype=int), "niter_l": z(P, Tf, dtype=int),
                           "Mp_s": z(P, Tf, dtype=int), "Mp_l": z(P, Tf, dtype=int)},
            "prices": {"lmbd_r": z(Tf), "avg_price_s": z(P, Tf), "avg_price_l": z(P, Tf),
                       "price_red_s": z(P, Tf), "price_red_l": z(P, Tf)},
        }

    def simulate(self) -> dict:
        for _ in range(self.Tf):
            self._step()
        return self.logs

    def _step(self):
        if settings.PRINT_LEVEL >= 1:
            print(f"===== closed-loop step {self.t} =====")
        lmbd_r = 0  # charging_station.py:162
        w_hat_s, w_hat_l, u_g, stats_bi = self._get_bimpc_solution(lmbd_r)
        prices_s, prices_l, stats_s, stats_l = self._get_optimal_prices(w_hat_s, w_hat_l, lmbd_r)
        w0_s, w0_l, price0_s, price0_l = self._get_w0_price0(prices_s, prices_l, lmbd_r)
        self._update_logs(lmbd_r, (w_hat_s, w_hat_l, u_g, w0_s, w0_l), (stats_bi, stats_s, stats_l),
                          (price0_s, price0_l))
        self._update_state(w0_s, w0_l, u_g[0])
        self.t += 1

    # ------------------------------------------------------------------ upper level
    def _partition_stats(self, solver: PriceSolver, y: np.ndarray, idx: np.ndarray, lmbd_r: float):
        """Mp, beta, gamma_m per partition (charging_station.py:196-211)."""
        Mp = np.zeros((self.P,), dtype=int)
        beta, gamma_m = np.zeros((self.P,)), np.zeros((self.P,))
        for p in range(self.P):
            mask = idx == p
            Mp[p] = mask.sum()
            if Mp[p] > 0:
                solver.set_charge_levels(y[mask])
                _, beta[p] = solver.get_robustness_bounds(lmbd_r)
                gamma_m[p] = solver.get_gamma_sm()
        return Mp, beta, gamma_m

    def _get_bimpc_solution(self, lmbd_r: float) -> tuple[np.ndarray, np.ndarray, np.ndarray, dict]:
        Mp_s, beta_s, gamma_sm = self._partition_stats(self.price_solver_s, self.y_s, self.idx_s, lmbd_r)
        Mp_l, beta_l, gamma_lm = self._partition_stats(self.price_solver_l, self.y_l, self.idx_l, lmbd_r)
        Mp_s_, Mp_l_ = Mp_s / self.B, Mp_l / self.B
        demand = self.demand[self.t: self.t + self.N_bi] / self.B
        params = BiMPCParameters(Mp_s_, Mp_l_, beta_s, beta_l, gamma_sm, gamma_lm, self.x, demand)
        w_hat_s, w_hat_l, u_g = self.bimpc.solve_bimpc(params)
        stats_bi = {"Mp_s": Mp_s, "Mp_l": Mp_l, "beta_s": beta_s, "beta_l": beta_l,
                    "gamma_sm": gamma_sm, "gamma_lm": gamma_lm}
        if settings.PRINT_LEVEL >= 1:  # the report of charging_station.py:229-263, condensed
            load0 = self.consts_s.theta * Mp_s_ @ w_hat_s[:, 0] + self.consts_l.theta * Mp_l_ @ w_hat_l[:, 0]
            ub0 = u_g[0] - demand[0] - load0
            margin = self.consts_s.theta * Mp_s_ @ beta_s + self.consts_l.theta * Mp_l_ @ beta_l
            print("partition sizes  small", Mp_s.tolist(), " large", Mp_l.tolist())
            print(f"plan for this step: generation {u_g[0]:.6f} (limit {self.consts_bi.u_g_max:g}), demand "
                  f"{demand[0]:.6f}, EV load {load0:.6f}")
            print(f"battery: rate in [{ub0 - margin:.5f}, {ub0 + margin:.5f}] (limit {self.consts_bi.u_b_max:g}), state "
                  f"{self.x:.5f} -> [{self.x + ub0 - margin:.5f}, {self.x + ub0 + margin:.5f}] "
                  f"(capacity {self.consts_bi.x_max:g})")
        return w_hat_s, w_hat_l, u_g, stats_bi

    # ------------------------------------------------------------------ price loop
    def _get_optimal_prices(self, w_hat_s: np.ndarray, w_hat_l: np.ndarray, lmbd_r: float
                            ) -> tuple[np.ndarray, np.ndarray, list, list]:
        """charging_station.py:265-305.  The two PriceSolver objects are shared by the P
        partitions of their EV type, so partition p warm-starts from the prices of the last
        non-empty partition before it (``prev_prices``); the order of the calls is kept."""
        w_ref = {"s": w_hat_s[:, : self.N_lo], "l": w_hat_l[:, : self.N_lo]}  # reduced horizon
        prices = {"s": np.zeros((self.P, self.r)), "l": np.zeros((self.P, self.r))}
        stats = {"s": [], "l": []}
        fleet = {"s": (self.price_solver_s, self.y_s, self.idx_s, "Small"),
                 "l": (self.price_solver_l, self.y_l, self.idx_l, "Large")}
        for p in range(self.P):
            for key in ("s", "l"):
                solver, y, idx, name = fleet[key]
                y0p = y[idx == p]
                if len(y0p) == 0:
                    stats[key].append({})
                    continue
                solver.set_charge_levels(y0p)
                if settings.PRINT_LEVEL >= 1:
                    print(f"[{name.lower()} EVs, partition {p}] ", end="")
                lmbd_, stats_ = solver.compute_optimal_prices(w_ref[key][p, :], lmbd_r)
                prices[key][p, :] = lmbd_[: self.r]
                stats[key].append(stats_)
        return prices["s"], prices["l"], stats["s"], stats["l"]

    def _get_w0_price0(self, prices_s: np.ndarray, prices_l: np.ndarray, lmbd_r: float
                       ) -> tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
        """charging_station.py:307-327; the P partitions of a type go to the GPU as one batch
        of groups (the per-partition results are independent of each other)."""
        out = []
        for solver, y, idx, prices in ((self.price_solver_s, self.y_s, self.idx_s, prices_s),
                                       (self.price_solver_l, self.y_l, self.idx_l, prices_l)):
            order = np.argsort(idx, kind="stable")
            counts = np.bincount(idx, minlength=self.P)
            off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
            lm = np.zeros((self.P, 3 * self.N_lo))
            lm[:, : self.r] = prices
            w0_sorted, price0 = solver.get_w0_price0_batch(off, y[order], lm, np.full(self.P, float(lmbd_r)))
            w0 = np.zeros((self.M_2,))
            w0[order] = w0_sorted
            out.append((w0, np.where(counts > 0, price0, 0.0)))
        (w0_s, price0_s), (w0_l, price0_l) = out
        return w0_s, w0_l, price0_s, price0_l

    # ------------------------------------------------------------------ plant
    def _update_state(self, w0_s: np.ndarray, w0_l: np.ndarray, u0_g: float) -> None:
        # Update EV SoCs and indices (charging_station.py:329-349): an EV whose SoC passes
        # MIN_FULL_CHARGE_FRACTION * y_max leaves and a new one arrives.
        residual_charge = 0
        for key, w0, consts in (("s", w0_s, self.consts_s), ("l", w0_l, self.consts_l)):
            y = self.y_s if key == "s" else self.y_l
            y += w0
            full = MIN_FULL_CHARGE_FRACTION * consts.y_max
            mask = y > full
            residual_charge += consts.theta * np.sum(y[mask] - full)
            y[mask] = self._draw_soc(mask.sum())
            if key == "s":
                self.ncharged_s += mask.sum()
            else:
                self.ncharged_l += mask.sum()
        self._update_indices()
        if not ADD_RESIDUAL_CHARGE_TO_BATTERY:
            residual_charge = 0
        # Update battery charge state (charging_station.py:353-365).
        u0_b = u0_g + (-self.consts_s.theta * np.sum(w0_s) - self.consts_l.theta * np.sum(w0_l)
                       + residual_charge - self.demand[self.t]) / self.B
        self.x += u0_b
        if settings.PRINT_LEVEL >= 1:
            print(f"departed so far: {self.ncharged_s} small, {self.ncharged_l} large EVs\n")

    def _update_logs(self, lmbd_r: float, nu: tuple, stats: tuple, price0: tuple) -> None:
        """charging_station.py:371-433 (same keys, same conventions: -1 iterations and NaN
        price reduction for an empty partition)."""
        w_hat_s, w_hat_l, u_g, w0_s, w0_l = nu
        stats_bi, stats_s, stats_l = stats
        price0_s, price0_l = price0
        t, L = self.t, self.logs
        for key, w0, idx, w_hat, st, p0 in (("s", w0_s, self.idx_s, w_hat_s, stats_s, price0_s),
                                            ("l", w0_l, self.idx_l, w_hat_l, stats_l, price0_l)):
            for p in range(self.P):
                sel = w0[idx == p]
                if len(sel) > 0:
                    L["inputs"]["w_" + key][p, t] = np.mean(sel)
                if st[p]:
                    L["statistics"]["niter_" + key][p, t] = st[p]["iter"]
                    L["prices"]["price_red_" + key][p, t] = st[p]["price_after_reg"] - st[p]["price_before_reg"]
                else:
                    L["statistics"]["niter_" + key][p, t] = -1
                    L["prices"]["price_red_" + key][p, t] = np.nan
            L["inputs"]["w_hat_" + key][:, t] = w_hat[:, 0]
            L["bounds"]["beta_" + key][:, t] = stats_bi["beta_" + key]
            L["statistics"]["gamma_" + key + "m"][:, t] = stats_bi["gamma_" + key + "m"]
            L["statistics"]["Mp_" + key][:, t] = stats_bi["Mp_" + key]
            L["prices"]["avg_price_" + key][:, t] = p0
        L["inputs"]["u_g"][t] = u_g[0]
        L["states"]["x"][t] = self.x
        L["statistics"]["ncharged_s"] = self.ncharged_s
        L["statistics"]["ncharged_l"] = self.ncharged_l
        L["prices"]["lmbd_r"][t] = lmbd_r
