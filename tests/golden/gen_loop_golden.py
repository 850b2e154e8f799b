"""Generates tests/golden/bimpc_golden.npz, price_loop_golden.npz and station_golden.npz with the CPU
oracles (oracle/bimpc_oracle.py, price_oracle.py, station_oracle.py).  The reference itself cannot run
in this image (cvxpy/CLARABEL absent), so these are ORACLE vectors, each BiMPC
point with its solver-independent certificate (objective, largest constraint violation).

    python tests/golden/gen_loop_golden.py        # ~1 minute
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "incentive-design-mpc_b200")):
    sys.path.insert(0, p)
from bimpc_cases import draw_station  # noqa: E402
from oracle import bimpc_oracle as bo  # noqa: E402
from oracle import lompc_oracle as orc  # noqa: E402
from oracle.price_oracle import PriceOracle  # noqa: E402
from oracle.station_oracle import StationOracle  # noqa: E402


def bimpc_golden():
    out = {}
    for cost_type in (bo.WEIGHTED, bo.UNWEIGHTED, bo.EXP_UNWEIGHTED):
        for N, P in ((16, 12), (24, 12)):
            c = bo.example_consts(N, P)
            c.cost_type = cost_type
            rng = np.random.default_rng(1000 + 10 * cost_type + N)
            stations = [draw_station(rng, c) for _ in range(4)]
            key = f"c{cost_type}_N{N}_P{P}"
            for i, name in enumerate(("Mp_s", "Mp_l", "beta_s", "beta_l", "gamma_sm", "gamma_lm", "x0", "demand")):
                out[f"{key}_{name}"] = np.stack([np.asarray(s[i], dtype=float) for s in stations])
            ws, wl, ug, obj, viol, its = [], [], [], [], [], []
            for par in stations:
                a, b, u, info = bo.solve_ipm(c, *par)
                assert info["status"] == 0
                k = bo.kkt_certificate(c, par, a, b, u)
                ws.append(a), wl.append(b), ug.append(u), obj.append(k["objective"]), viol.append(k["max_violation"])
                its.append(info["iters"])
            out[f"{key}_w_hat_s"], out[f"{key}_w_hat_l"], out[f"{key}_u_g"] = np.stack(ws), np.stack(wl), np.stack(ug)
            out[f"{key}_objective"], out[f"{key}_max_violation"] = np.array(obj), np.array(viol)
            out[f"{key}_iters"] = np.array(its)
    np.savez_compressed(os.path.join(HERE, "bimpc_golden.npz"), **out)


def price_loop_golden():
    out = {}
    for ev, o in (("small", orc.small_ev_consts()), ("large", orc.large_ev_consts())):
        for price_type in ("linear", "linear-convex"):
            N, nev, G = 12, 7, 3
            rng = np.random.default_rng(7 + 10 * (ev == "large") + (price_type == "linear"))
            key = f"{ev}_{price_type}"
            y0 = 0.3 + 0.05 * rng.random((G, nev)) + 0.05 * np.arange(G)[:, None]
            w_ref = o.w_max * rng.random((G, N)) * 0.6
            po = PriceOracle(N, o, price_type)
            prices, iters, pre, post, w0s, p0s = [], [], [], [], [], []
            for g in range(G):  # one shared solver: the warm start chains through the groups
                po.set_charge_levels(y0[g])
                lm, st = po.compute_optimal_prices(w_ref[g], 0.0)
                prices.append(lm.copy()), iters.append(st["iter"]), pre.append(st["price_before_reg"])
                post.append(st["price_after_reg"])
                w0, p0 = po.get_w0_price0(lm[: po.r], 0.0)
                w0s.append(w0), p0s.append(p0)
            out[f"{key}_y0"], out[f"{key}_w_ref"] = y0, w_ref
            out[f"{key}_prices"], out[f"{key}_iters"] = np.stack(prices), np.array(iters)
            out[f"{key}_pre"], out[f"{key}_post"] = np.array(pre), np.array(post)
            out[f"{key}_w0"], out[f"{key}_price0"] = np.stack(w0s), np.array(p0s)
    np.savez_compressed(os.path.join(HERE, "price_loop_golden.npz"), **out)


def station_golden():
    from chargingstation.demand_data import medium_term_demand_forecast
    Tf, N_bi, N_lo, M2, P = 3, 8, 4, 24, 6
    out = {"sizes": np.array([Tf, N_bi, N_lo, M2, P])}
    dem = medium_term_demand_forecast(Tf + N_bi + 1, 0.25) * (M2 / 500)
    bi = bo.example_consts(N_bi, P)
    bi.cost_type = bo.UNWEIGHTED
    np.random.seed(0)
    so = StationOracle(Tf, N_bi, N_lo, M2, P, dem, bi, orc.small_ev_consts(), orc.large_ev_consts(), "linear-convex")
    for _ in range(Tf):
        so.step()
    for name in ("u_g", "w_hat_s", "w_hat_l", "x_before", "x_after", "niter_s", "niter_l", "Mp_s", "Mp_l", "price0_s",
                 "price0_l", "w0_s", "w0_l"):
        out[name] = np.stack([np.asarray(r[name]) for r in so.trace])
    out["y_s_final"], out["y_l_final"] = so.y["s"], so.y["l"]
    np.savez_compressed(os.path.join(HERE, "station_golden.npz"), **out)


if __name__ == "__main__":
    bimpc_golden()
    price_loop_golden()
    station_golden()
    print("written:", [f for f in os.listdir(HERE) if f.endswith(".npz")])
