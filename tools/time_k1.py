#!/usr/bin/env python
"""Device timing of K1 at small and medium batches: the warp-cooperative kernel (variants 8 / 9), the register
thread kernel (4) and the any-N kernel (1), CUDA events around each launch, L2 flushed between launches.
    python tools/time_k1.py [--N 24] [--batches 1024,4096,...] [--reps 20]
Also times the solve set (one launch for small + large EVs) and its host round trip.  One JSON line per case."""
import argparse
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
    import torch
    from chargingstation.lompc import LoMPC, LoMPCConstants, LoMPCSet
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=24)
    ap.add_argument("--batches", default="512,2048,8192,32768,131072")
    ap.add_argument("--variants", default="8,4,1")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--set-batch", type=int, default=512, help="QPs per EV type of the solve-set case")
    args = ap.parse_args()
    N = args.N
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(fn, reps):
        for _ in range(3):
            fn()
        ms = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        return float(np.median(ms)), float(np.min(ms))

    for ev in ("small", "large"):
        delta, theta, y_max, w_max = EV[ev]
        for B in [int(x) for x in args.batches.split(",")]:
            rng = np.random.default_rng(B)
            lm = torch.from_numpy(theta * rng.random((B, 3 * N))).to(dev)
            lr = torch.from_numpy(3 * N * delta * rng.random(B)).to(dev)
            gam = torch.from_numpy(y_max * rng.random(B)).to(dev)
            out = (torch.empty((B, N), dtype=torch.float64, device=dev), torch.empty(B, dtype=torch.float64, device=dev))
            for var in [int(v) for v in args.variants.split(",")]:
                if var == 4 and N not in (12, 24):
                    continue
                s = LoMPC(N, LoMPCConstants(delta, theta, y_max, w_max, ev))
                s.set_kernel_variant(var)
                _, _, info = s.solve_lompc_batch(lm, lr, gam, return_info=True)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    s.solve_lompc_batch(lm, lr, gam, out=out)
                med, mn = timed(g.replay, args.reps)
                print(json.dumps({"case": "k1", "ev": ev, "N": N, "B": B, "variant": var, "us_median": med * 1e3,
                                  "us_min": mn * 1e3, "MQPs": B / med / 1e3,
                                  "iters_mean": float(info["iters"].double().mean()),
                                  "iters_max": int(info["iters"].max()), "bad": int((info["status"] != 0).sum())}),
                      flush=True)

    # the solve set: small + large in one launch
    Bs = args.set_batch
    solvers = [LoMPC(N, LoMPCConstants(*EV[ev], ev)) for ev in ("small", "large")]
    sset = LoMPCSet(solvers, [Bs, Bs])
    for i, ev in enumerate(("small", "large")):
        delta, theta, y_max, w_max = EV[ev]
        rng = np.random.default_rng(2 + i)
        sset.lmbd[i][:] = theta * rng.random((Bs, 3 * N))
        sset.lmbd_r[i][:] = 3 * N * delta * rng.random(Bs)
        sset.gamma[i][:] = y_max * rng.random(Bs)
    sset.solve(info=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        sset.solve_dev(torch.cuda.current_stream().cuda_stream)
    med, mn = timed(g.replay, args.reps)
    print(json.dumps({"case": "set_dev", "N": N, "B": 2 * Bs, "us_median": med * 1e3, "us_min": mn * 1e3,
                      "MQPs": 2 * Bs / med / 1e3, "iters_max": [int(x.max()) for x in sset.iters]}), flush=True)
    for label, env in (("graph", None), ("no_graph", "LOMPC_SET_NO_GRAPH"), ("mapped", "LOMPC_SET_MAPPED")):
        if env:
            os.environ.pop("LOMPC_SET_NO_GRAPH", None)
            os.environ[env] = "1"
            sset2 = LoMPCSet(solvers, [Bs, Bs])
            for i in range(2):
                sset2.lmbd[i][:], sset2.lmbd_r[i][:], sset2.gamma[i][:] = sset.lmbd[i], sset.lmbd_r[i], sset.gamma[i]
        else:
            sset2 = sset
        for _ in range(10):
            sset2.solve()
        assert all(np.array_equal(sset2.w[i], sset.w[i]) for i in range(2)), label
        ts = []
        for _ in range(200):
            t0 = time.perf_counter()
            sset2.solve()
            ts.append(time.perf_counter() - t0)
        print(json.dumps({"case": "set_host_" + label, "N": N, "B": 2 * Bs, "us_median": float(np.median(ts)) * 1e6,
                          "us_min": float(np.min(ts)) * 1e6, "h2d": sset2.h2d_bytes, "d2h": sset2.d2h_bytes}),
              flush=True)


if __name__ == "__main__":
    main()
