"""CPU ORACLE (test infrastructure, NOT a product path) for the closed loop of one charging
station: restates ``chargingstation/charging_station.py:156-370`` (one ``_step``: BiMPC ->
price loop per partition -> EV responses -> SoC / battery update) on the oracle solvers
(``bimpc_oracle``, ``price_oracle``, ``lompc_oracle``).  PARITY UNPINNED like its parts
(cvxpy/CLARABEL are not installable and the reference holds no closed-loop assertions); it
is slow (one Python QP solve per EV per price iteration) and meant for small instances.

Random draws follow the reference's order on ``np.random`` (charging_station.py:95-100,
333-346) so that a seeded run is comparable step by step."""
from __future__ import annotations

import numpy as np

from oracle import bimpc_oracle as bo
from oracle import lompc_oracle as orc
from oracle.price_oracle import PriceOracle

# settings.py:27-33
MIN_INITIAL_SOC, MAX_INITIAL_SOC = 0.3, 0.5
MIN_FULL_CHARGE_FRACTION = 0.95


class StationOracle:
    def __init__(self, Tf, N_bi, N_lo, M_2, P, demand, bi: bo.BiConsts, cs: orc.OracleConsts,
                 cl: orc.OracleConsts, price_type: str, fast: bool = False):
        """``fast``: the price loops solve their QPs with the C twin of the exact oracle (see PriceOracle)."""
        assert N_bi >= N_lo >= 1 and demand.shape[0] >= Tf + N_bi + 1  # charging_station.py:44-53
        self.Tf, self.N_bi, self.N_lo, self.M_2, self.P = Tf, N_bi, N_lo, M_2, P
        self.demand, self.bi, self.cs, self.cl = demand, bi, cs, cl
        self.r = 2 * N_lo if price_type == "linear" else 3 * N_lo
        self.ps = {"s": PriceOracle(N_lo, cs, price_type, fast=fast), "l": PriceOracle(N_lo, cl, price_type, fast=fast)}
        self.edges = {"s": np.linspace(MIN_INITIAL_SOC, cs.y_max, P + 1),
                      "l": np.linspace(MIN_INITIAL_SOC, cl.y_max, P + 1)}
        self.B = (cs.theta + cl.theta) * M_2
        draw = lambda n: MIN_INITIAL_SOC + (MAX_INITIAL_SOC - MIN_INITIAL_SOC) * np.random.random((n,))  # noqa
        self._draw = draw
        self.y = {"s": draw(M_2), "l": draw(M_2)}
        self.x = 0.0
        self.t = 0
        self.idx = {"s": np.zeros(M_2, dtype=int), "l": np.zeros(M_2, dtype=int)}
        self._update_indices()
        self.trace = []

    def _update_indices(self):  # charging_station.py:111-116
        for k in ("s", "l"):
            for p in range(self.P):
                m = (self.y[k] >= self.edges[k][p]) & (self.y[k] <= self.edges[k][p + 1])
                self.idx[k][m] = p

    def step(self):
        P, lmbd_r = self.P, 0.0
        st = {}
        for k in ("s", "l"):  # charging_station.py:196-211
            Mp, beta, gm = np.zeros(P, dtype=int), np.zeros(P), np.zeros(P)
            for p in range(P):
                m = self.idx[k] == p
                Mp[p] = m.sum()
                if Mp[p] > 0:
                    self.ps[k].set_charge_levels(self.y[k][m])
                    _, beta[p] = self.ps[k].get_robustness_bounds(lmbd_r)
                    gm[p] = self.ps[k].gamma_sm
            st[k] = (Mp, beta, gm)
        dem = self.demand[self.t: self.t + self.N_bi] / self.B
        w_hat_s, w_hat_l, u_g, info = bo.solve_ipm(self.bi, st["s"][0] / self.B, st["l"][0] / self.B, st["s"][1],
                                                   st["l"][1], st["s"][2], st["l"][2], self.x, dem)
        assert info["status"] == 0
        w_hat = {"s": w_hat_s, "l": w_hat_l}
        prices = {k: np.zeros((P, self.r)) for k in ("s", "l")}
        niter = {k: -np.ones(P, dtype=int) for k in ("s", "l")}
        for p in range(P):  # charging_station.py:273-304
            for k in ("s", "l"):
                y0p = self.y[k][self.idx[k] == p]
                if len(y0p) == 0:
                    continue
                self.ps[k].set_charge_levels(y0p)
                lm, stats = self.ps[k].compute_optimal_prices(w_hat[k][p, : self.N_lo], lmbd_r)
                prices[k][p] = lm[: self.r]
                niter[k][p] = stats["iter"]
        w0 = {k: np.zeros(self.M_2) for k in ("s", "l")}
        price0 = {k: np.zeros(P) for k in ("s", "l")}
        for p in range(P):  # charging_station.py:313-326
            for k in ("s", "l"):
                m = self.idx[k] == p
                if m.sum() > 0:
                    self.ps[k].set_charge_levels(self.y[k][m])
                    w0[k][m], price0[k][p] = self.ps[k].get_w0_price0(prices[k][p], lmbd_r)
        rec = {"u_g": u_g.copy(), "w_hat_s": w_hat_s.copy(), "w_hat_l": w_hat_l.copy(), "x_before": self.x,
               "prices_s": prices["s"].copy(), "prices_l": prices["l"].copy(), "niter_s": niter["s"],
               "niter_l": niter["l"], "w0_s": w0["s"].copy(), "w0_l": w0["l"].copy(),
               "price0_s": price0["s"], "price0_l": price0["l"], "Mp_s": st["s"][0], "Mp_l": st["l"][0],
               "y_s": self.y["s"].copy(), "y_l": self.y["l"].copy()}
        # plant update, charging_station.py:329-365 (ADD_RESIDUAL_CHARGE_TO_BATTERY = False)
        for k, c in (("s", self.cs), ("l", self.cl)):
            self.y[k] += w0[k]
            m = self.y[k] > MIN_FULL_CHARGE_FRACTION * c.y_max
            self.y[k][m] = self._draw(m.sum())
        self._update_indices()
        self.x += u_g[0] + (-self.cs.theta * np.sum(w0["s"]) - self.cl.theta * np.sum(w0["l"])
                            - self.demand[self.t]) / self.B
        rec["x_after"] = self.x
        self.trace.append(rec)
        self.t += 1
        return rec
