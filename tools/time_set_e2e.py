#!/usr/bin/env python
"""Host round-trip latency of the solve set (LoMPCSet.solve) against the batch size, in its transfer modes:
zero-copy (kernel reads / writes the mapped pinned blocks) vs staged (H2D copy, launch, D2H copy), CUDA graph vs plain
stream calls.   python tools/time_set_e2e.py"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "incentive-design-mpc_b200")):
    sys.path.insert(0, p)
EV = {"small": (0.05, 10.0, 0.9, 0.25), "large": (0.025, 50.0, 0.9, 0.15)}


def main():
    from chargingstation.lompc import LoMPC, LoMPCConstants, LoMPCSet
    N = 24
    solvers = [LoMPC(N, LoMPCConstants(*EV[ev], ev)) for ev in ("large", "small")]
    for Bs in (1, 32, 128, 512):
        for mapped, graph in (("1", "0"), ("1", "1"), ("0", "0"), ("0", "1")):
            os.environ["LOMPC_SET_MAPPED"] = mapped
            os.environ["LOMPC_SET_NO_GRAPH"] = "0" if graph == "1" else "1"
            sset = LoMPCSet(solvers, [Bs, Bs])
            for i, ev in enumerate(("large", "small")):
                delta, theta, y_max, w_max = EV[ev]
                rng = np.random.default_rng(2 + i)
                sset.lmbd[i][:] = theta * rng.random((Bs, 3 * N))
                sset.lmbd_r[i][:] = 3 * N * delta * rng.random(Bs)
                sset.gamma[i][:] = y_max * rng.random(Bs)
            for _ in range(20):
                sset.solve()
            ts = []
            for _ in range(300):
                t0 = time.perf_counter()
                sset.solve()
                ts.append(time.perf_counter() - t0)
            print(json.dumps({"QPs": 2 * Bs, "mapped": mapped == "1", "graph": graph == "1",
                              "us_median": float(np.median(ts)) * 1e6, "us_min": float(np.min(ts)) * 1e6}), flush=True)


if __name__ == "__main__":
    main()
