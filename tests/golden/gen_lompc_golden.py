"""Generates tests/golden/lompc_golden.npz with the CPU oracle (oracle/lompc_oracle.py).

The reference itself cannot run in this image (cvxpy/CLARABEL absent, SURVEY.md
section 8c), so these are ORACLE vectors: inputs drawn with the distributions of
the reference's own scripts (test/test_lompc.py:34-36, charging_station.py:95-100)
and the exact optimum, each with its solver-independent KKT certificate.

    python tests/golden/gen_lompc_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import lompc_oracle as orc  # noqa: E402


def draw(rng, N, consts, B, mode):
    th = consts.theta
    if mode == 0:  # test_lompc.py:34-36
        lm = th * rng.random((B, 3 * N))
        lr = 3 * N * consts.delta * rng.random(B)
        gam = consts.y_max * rng.random(B)
    elif mode == 1:  # closed-loop scale: small prices, lmbd_r = 0, gamma = y_max - U(0.3, 0.5)
        lm = 0.05 * th * rng.random((B, 3 * N))
        lr = np.zeros(B)
        gam = consts.y_max - (0.3 + 0.2 * rng.random(B))
    elif mode == 2:  # "linear" price type: lmbd3 = 0
        lm = np.zeros((B, 3 * N))
        lm[:, :2 * N] = 0.05 * th * rng.random((B, 2 * N))
        lr = np.zeros(B)
        gam = consts.y_max * rng.random(B)
    else:  # unpriced (test_lompc.py:47-48 uses lmbd = 0)
        lm = np.zeros((B, 3 * N))
        lr = np.zeros(B)
        gam = consts.y_max * rng.random(B)
        gam[0] = consts.y_max
    return lm, lr, gam


def main():
    rng = np.random.default_rng(20240818)
    out = {}
    B = 24
    for consts in (orc.small_ev_consts(), orc.large_ev_consts()):
        for N in (12, 24):
            for mode in range(4):
                lm, lr, gam = draw(rng, N, consts, B, mode)
                w = np.zeros((B, N))
                cost = np.zeros(B)
                kkt = np.zeros(B)
                for b in range(B):
                    w[b], cost[b], _ = orc.solve_active_set(N, consts, lm[b], lr[b], gam[b])
                    kkt[b], _ = orc.kkt_certificate(N, consts, w[b], lm[b], lr[b], gam[b])
                key = f"{consts.ev_type}_N{N}_m{mode}"
                out[key + "_lmbd"], out[key + "_lmbd_r"], out[key + "_gamma"] = lm, lr, gam
                out[key + "_w"], out[key + "_cost"], out[key + "_kkt"] = w, cost, kkt
                print(key, "max kkt", kkt.max())
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "lompc_golden.npz"), **out)


if __name__ == "__main__":
    main()
