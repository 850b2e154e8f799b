#!/usr/bin/env python
"""A/B of two builds of the library on the batched BiMPC (K6): solves the same random stations and saves the results,
so that a change that must not alter the arithmetic can be checked to the bit.
    python tools/k6_ab.py new;  LOMPC_B200_LIB=/path/to/other/liblompc_b200.so python tools/k6_ab.py old
    python tools/k6_ab.py --compare new old"""
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "incentive-design-mpc_b200")):
    sys.path.insert(0, p)
OUT = os.path.join(ROOT, "gpurun_out")

if sys.argv[1] == "--compare":
    a, b = (np.load(os.path.join(OUT, f"k6_{t}.npz")) for t in sys.argv[2:4])
    bad = [k for k in a.files if not np.array_equal(a[k], b[k])]
    print("identical" if not bad else f"DIFFERENT: {bad}")
    sys.exit(1 if bad else 0)

from bimpc_cases import draw_station, stack  # noqa: E402
from oracle import bimpc_oracle as bo  # noqa: E402
from chargingstation.bimpc import BiMPC, BiMPCChargingCostType, BiMPCConstants  # noqa: E402
from chargingstation.lompc import LoMPCConstants  # noqa: E402

tag = sys.argv[1]
out = {}
os.makedirs(OUT, exist_ok=True)
for ct in (bo.WEIGHTED, bo.UNWEIGHTED, bo.EXP_UNWEIGHTED):
    for (N, P) in ((24, 12), (16, 12), (8, 3)):
        c = bo.example_consts(N, P)
        c.cost_type = ct
        cb = BiMPCConstants(c.delta, c.c_g, c.u_g_max, c.u_b_max, c.x_max, BiMPCChargingCostType(c.cost_type), c.exp_rate)
        b = BiMPC(c.N, c.P, cb, LoMPCConstants(0.05, c.theta_s, 0.9, c.w_max_s, "small"),
                  LoMPCConstants(0.025, c.theta_l, 0.9, c.w_max_l, "large"))
        rng = np.random.default_rng(7)
        args = stack([draw_station(rng, c) for _ in range(1776)])
        b.solve_bimpc_batch(*args)
        t0 = time.perf_counter()
        ws, wl, ug, info = b.solve_bimpc_batch(*args)
        print(tag, ct, N, P, f"{(time.perf_counter() - t0) * 1e3:.1f} ms  iters mean {info['iters'].mean():.2f} "
              f"bad {int((info['status'] != 0).sum())}", flush=True)
        for name, v in (("ws", ws), ("wl", wl), ("ug", ug), ("it", info["iters"]), ("obj", info["objective"])):
            # (digests, not arrays: gpurun brings back at most 64 MiB)
            out[f"{ct}_{N}_{P}_{name}"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(v).tobytes()).digest(), dtype=np.uint8)
np.savez(os.path.join(OUT, f"k6_{tag}.npz"), **out)
