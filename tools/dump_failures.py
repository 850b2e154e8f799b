"""Runs large batches on the GPU and dumps the inputs of every QP whose status != 0."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "incentive-design-mpc_b200")); sys.path.insert(0, ROOT)
from bench import EV_CONSTS
from chargingstation.lompc import LoMPC, LoMPCConstants
import torch
out = {}
dev = torch.device("cuda:0")
for ev in ("small", "large"):
    delta, theta, y_max, w_max = EV_CONSTS[ev]
    for N in (24, 96):
        solver = LoMPC(N, LoMPCConstants(delta, theta, y_max, w_max, ev))
        for mode in (0, 1, 2, 3, 4):
            B = 1 << 20 if N == 24 else 1 << 17
            rng = np.random.default_rng(mode)
            if mode == 0:
                lm, lr, gam = theta * rng.random((B, 3 * N)), 3 * N * delta * rng.random(B), y_max * rng.random(B)
            elif mode == 1:
                lm = 0.05 * theta * rng.random((B, 3 * N)) * (rng.random((B, 3 * N)) < 0.5)
                lr, gam = np.zeros(B), y_max - (0.3 + 0.2 * rng.random(B))
            elif mode == 2:
                lm = np.zeros((B, 3 * N)); lm[:, :2 * N] = 0.05 * theta * rng.random((B, 2 * N))
                lr, gam = np.zeros(B), y_max * rng.random(B)
            elif mode == 3:
                lm, lr, gam = np.zeros((B, 3 * N)), np.zeros(B), y_max * rng.random(B)
            else:  # nearly degenerate but strictly convex stages: the optimistic phase of K1 hands over
                lm = 0.05 * theta * rng.random((B, 3 * N)) * (rng.random((B, 3 * N)) < 0.5)
                lr, gam = np.full(B, 1e-9), y_max - (0.3 + 0.2 * rng.random(B))
            t = [torch.from_numpy(x).to(dev) for x in (lm, lr, gam)]
            w, cost, info = solver.solve_lompc_batch(*t, return_info=True)
            st = info["status"].cpu().numpy(); it = info["iters"].cpu().numpy(); kk = info["kkt_res"].cpu().numpy()
            bad = np.flatnonzero(st != 0)
            print(ev, N, "mode", mode, "B", B, "failures", len(bad), "iters mean %.2f max %d" % (it.mean(), it.max()),
                  "kkt max %.1e" % kk.max(), flush=True)
            if len(bad):
                sel = bad[:32]
                key = f"{ev}_N{N}_m{mode}"
                out[key + "_lmbd"], out[key + "_lmbd_r"], out[key + "_gamma"] = lm[sel], lr[sel], gam[sel]
                out[key + "_w"] = w.cpu().numpy()[sel]; out[key + "_iters"] = it[sel]; out[key + "_kkt"] = kk[sel]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "failures.npz"), **out)
