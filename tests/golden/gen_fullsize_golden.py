"""Generates tests/golden/fullsize_station_golden.npz and price_loop_n24_golden.npz: ORACLE runs of the closed
loop at the reference's FULL sizes (BASELINE.json configs[0]: example/real_time_price_control.py:12-23 -
500 + 500 EVs, P = 12, N_lo = 12, N_bi = 16, 49 steps, "linear-convex", np.random.seed(0)) and at the north-star
horizon (N_lo = N_bi = 24), plus whole price loops at N = 24 with groups larger than one CTA pass.

The reference itself cannot run in this image (cvxpy/CLARABEL absent), so these are vectors
of oracle/station_oracle.py (charging_station.py:156-370 restated) with the price loops on the C twin of the exact
active-set LoMPC oracle (oracle/lompc_oracle.c::solve_exact_one, checked against the Python one to 1e-15 in
tests/test_oracle.py).  Per scenario the file holds a per-step summary of the whole run and, at selected steps,
everything a teacher-forced comparison needs: the state entering the step (SoCs, battery, warm-start prices), the
BiMPC plan, and the price loops' outputs.

    python tests/golden/gen_fullsize_golden.py        # ~10 minutes on 8 cores
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "incentive-design-mpc_b200")):
    sys.path.insert(0, p)
from oracle import bimpc_oracle as bo  # noqa: E402
from oracle import lompc_oracle as orc  # noqa: E402
from oracle.price_oracle import PriceOracle  # noqa: E402
from oracle.station_oracle import StationOracle  # noqa: E402

SCENARIOS = {  # name: (Tf, N_bi, N_lo, M2, P, cost_type, steps recorded in full)
    "cfg0_unw": (49, 16, 12, 500, 12, bo.UNWEIGHTED, (0, 1, 2, 3, 4, 8, 16, 24, 32, 40, 48)),
    "cfg0_exp": (49, 16, 12, 500, 12, bo.EXP_UNWEIGHTED, (0, 1, 2, 3, 4, 8, 16, 18, 24, 32, 40, 48)),
    "n24_unw": (6, 24, 24, 500, 12, bo.UNWEIGHTED, (0, 1, 2, 3, 4, 5)),
}


def station_scenario(name, out):
    from chargingstation.demand_data import medium_term_demand_forecast
    Tf, N_bi, N_lo, M2, P, cost_type, full = SCENARIOS[name]
    dem = medium_term_demand_forecast(Tf + N_bi + 1, 0.25, interpolate=False)  # example: DEMAND_SCALE = 1/4
    bi = bo.example_consts(N_bi, P)
    bi.cost_type = cost_type
    np.random.seed(0)
    so = StationOracle(Tf, N_bi, N_lo, M2, P, dem, bi, orc.small_ev_consts(), orc.large_ev_consts(), "linear-convex",
                       fast=True)
    out[f"{name}_sizes"] = np.array([Tf, N_bi, N_lo, M2, P, cost_type])
    out[f"{name}_full_steps"] = np.array(full)
    t0 = time.time()
    for t in range(Tf):
        prev = {k: so.ps[k].prev_prices.copy() for k in ("s", "l")}
        rng_state = np.random.get_state()
        rec = so.step()
        if t in full:
            for k in ("s", "l"):
                out[f"{name}_t{t}_y_{k}"] = rec["y_" + k]
                out[f"{name}_t{t}_prev_prices_{k}"] = prev[k]
                out[f"{name}_t{t}_w_hat_{k}"] = rec["w_hat_" + k]
                out[f"{name}_t{t}_prices_{k}"] = rec["prices_" + k]
                out[f"{name}_t{t}_w0_{k}"] = rec["w0_" + k]
            out[f"{name}_t{t}_u_g"] = rec["u_g"]
            out[f"{name}_t{t}_rng_keys"] = rng_state[1]  # np.random state entering the step (MT19937 key vector)
            out[f"{name}_t{t}_rng_pos"] = np.array([rng_state[2]])
        print(f"[{name}] step {t}: niter_s {rec['niter_s'].tolist()} niter_l {rec['niter_l'].tolist()} "
              f"x {rec['x_after']:.6f}  ({time.time() - t0:.0f} s)", flush=True)
    tr = so.trace
    out[f"{name}_u_g0"] = np.array([r["u_g"][0] for r in tr])
    out[f"{name}_x_before"] = np.array([r["x_before"] for r in tr])
    out[f"{name}_x_after"] = np.array([r["x_after"] for r in tr])
    for k in ("s", "l"):
        out[f"{name}_niter_{k}"] = np.stack([r["niter_" + k] for r in tr])
        out[f"{name}_Mp_{k}"] = np.stack([r["Mp_" + k] for r in tr])
        out[f"{name}_price0_{k}"] = np.stack([r["price0_" + k] for r in tr])
        out[f"{name}_w0sum_{k}"] = np.array([r["w0_" + k].sum() for r in tr])
        out[f"{name}_w_hat0_{k}"] = np.stack([r["w_hat_" + k][:, 0] for r in tr])
        out[f"{name}_y_final_{k}"] = so.y[k]
    out[f"{name}_lompc_solves"] = np.array([so.ps["s"].lompc_solves, so.ps["l"].lompc_solves])


def capped_groups(out, want=3):
    """Group instances on which the ORACLE loop runs into the reference's cap of 1000 price iterations
    (settings.py:14), harvested from runs of the example (EXP_UNWEIGHTED) with seeds 0, 1, ...: the group's
    SoCs, its BiMPC reference, the warm start entering the loop, and the oracle's result."""
    from chargingstation.demand_data import medium_term_demand_forecast
    Tf, N_bi, N_lo, M2, P, cost_type, _ = SCENARIOS["cfg0_exp"]
    dem = medium_term_demand_forecast(Tf + N_bi + 1, 0.25, interpolate=False)
    found = []
    for seed in range(32):
        bi = bo.example_consts(N_bi, P)
        bi.cost_type = cost_type
        np.random.seed(seed)
        so = StationOracle(Tf, N_bi, N_lo, M2, P, dem, bi, orc.small_ev_consts(), orc.large_ev_consts(), "linear-convex",
                           fast=True)
        for k in ("s", "l"):
            po = so.ps[k]
            orig = po.compute_optimal_prices

            def wrapped(w_ref, lmbd_r, _po=po, _orig=orig, _k=k, **kw):
                prev = _po.prev_prices.copy()
                lm, st = _orig(w_ref, lmbd_r, **kw)
                if st["iter"] >= 999:
                    found.append({"ev": _k, "seed": seed, "y0": _po.y0.copy(), "w_ref": np.array(w_ref, dtype=float),
                                  "prev": prev, "prices": lm.copy(), "pre": st["price_before_reg"],
                                  "post": st["price_after_reg"]})
                return lm, st

            po.compute_optimal_prices = wrapped
        for t in range(Tf):
            so.step()
            if len(found) >= want:
                break
        print(f"[capped] seed {seed}: {len(found)} capped groups so far", flush=True)
        if len(found) >= want:
            break
    out["capped_count"] = np.array([len(found)])
    for i, f in enumerate(found):
        out[f"capped_{i}_is_large"] = np.array([f["ev"] == "l"])
        for key in ("y0", "w_ref", "prev", "prices"):
            out[f"capped_{i}_{key}"] = f[key]
        out[f"capped_{i}_pre_post"] = np.array([f["pre"], f["post"]])


def price_loop_n24(out):
    """Whole loops (price_solver.py:79-174) at N = 24: three chained groups per case (one shared warm start),
    70 EVs each (more than one 64-thread CTA pass), inputs as test/test_price_solver.py:23-35."""
    N, nev, G = 24, 70, 3
    for ev, o in (("small", orc.small_ev_consts()), ("large", orc.large_ev_consts())):
        for price_type, lmbd_r in (("linear-convex", 0.0), ("linear", 0.0), ("linear-convex", 24.0)):
            rng = np.random.default_rng(240 + 10 * (ev == "large") + (price_type == "linear") + int(lmbd_r))
            key = f"{ev}_{price_type}_lr{int(lmbd_r)}"
            y0 = 0.3 + 0.04 * rng.random((G, nev)) + 0.05 * np.arange(G)[:, None]
            w_ref = o.w_max * rng.random((G, N)) * 0.6
            po = PriceOracle(N, o, price_type, fast=True)
            prices, iters, pre, post, p0s, w0s, dec_ac, dec_pr = [], [], [], [], [], [], [], []
            for g in range(G):
                po.set_charge_levels(y0[g])
                lm, st = po.compute_optimal_prices(w_ref[g], lmbd_r)
                prices.append(lm.copy()), iters.append(st["iter"]), pre.append(st["price_before_reg"])
                post.append(st["price_after_reg"])
                w0, p0 = po.get_w0_price0(lm[: po.r], lmbd_r)
                w0s.append(w0), p0s.append(p0)
                dec_ac.append(np.pad(st["dual_cost_decrease_actual"], (0, 1000))[:1000])
                dec_pr.append(np.pad(st["dual_cost_decrease_predicted"], (0, 1000))[:1000])
            print(f"[price_loop_n24] {key}: iters {iters}", flush=True)
            out[f"{key}_y0"], out[f"{key}_w_ref"] = y0, w_ref
            out[f"{key}_prices"], out[f"{key}_iters"] = np.stack(prices), np.array(iters)
            out[f"{key}_pre"], out[f"{key}_post"] = np.array(pre), np.array(post)
            out[f"{key}_w0"], out[f"{key}_price0"] = np.stack(w0s), np.array(p0s)
            n = max(iters) + 1
            out[f"{key}_dec_actual"], out[f"{key}_dec_predicted"] = np.stack(dec_ac)[:, :n], np.stack(dec_pr)[:, :n]


if __name__ == "__main__":
    which = sys.argv[1:] or ["price", "capped", "n24_unw", "cfg0_unw", "cfg0_exp"]
    if "price" in which:
        out = {}
        price_loop_n24(out)
        np.savez_compressed(os.path.join(HERE, "price_loop_n24_golden.npz"), **out)
    if "capped" in which:
        out = {}
        capped_groups(out)
        np.savez_compressed(os.path.join(HERE, "capped_groups_golden.npz"), **out)
    names = [w for w in which if w in SCENARIOS]
    if names:
        path = os.path.join(HERE, "fullsize_station_golden.npz")
        out = dict(np.load(path)) if os.path.exists(path) else {}
        for name in names:
            out = {k: v for k, v in out.items() if not k.startswith(name + "_")}
            station_scenario(name, out)
            np.savez_compressed(path, **out)
    print("written:", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))
