"""Builds tests/golden/lompc_hard_cases.npz: inputs on which an earlier build of the
kernel stalled (near-degenerate active sets found by running 10^6-QP batches on a B200,
tools/dump_failures.py), with the oracle's exact optimum for each.

    python tests/golden/gen_hard_cases.py gpurun_out/failures.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import lompc_oracle as orc  # noqa: E402

src = np.load(sys.argv[1])
dst_path = os.path.join(os.path.dirname(__file__), "lompc_hard_cases.npz")
out = dict(np.load(dst_path)) if os.path.exists(dst_path) else {}
keys = sorted(set("_".join(k.split("_")[:3]) for k in src.files))
for key in keys:
    ev, Ns, _ = key.split("_")
    N = int(Ns[1:])
    consts = orc.small_ev_consts() if ev == "small" else orc.large_ev_consts()
    lm, lr, gam = src[key + "_lmbd"][:8], src[key + "_lmbd_r"][:8], src[key + "_gamma"][:8]
    if key + "_lmbd" in out:  # append new cases
        lm = np.concatenate([out[key + "_lmbd"], lm])
        lr = np.concatenate([out[key + "_lmbd_r"], lr])
        gam = np.concatenate([out[key + "_gamma"], gam])
    w = np.zeros((len(gam), N))
    cost = np.zeros(len(gam))
    for b in range(len(gam)):
        w[b], cost[b], _ = orc.solve_active_set(N, consts, lm[b], lr[b], gam[b])
        viol, _ = orc.kkt_certificate(N, consts, w[b], lm[b], lr[b], gam[b])
        assert viol < 1e-8, (key, b, viol)
    out[key + "_lmbd"], out[key + "_lmbd_r"], out[key + "_gamma"] = lm, lr, gam
    out[key + "_w"], out[key + "_cost"] = w, cost
    print(key, len(gam), "cases")
np.savez_compressed(dst_path, **out)
