"""Small driver for ncu: a few launches of the LoMPC solve kernel at one batch size.
    python tools/profile_solve.py --ev small --batch 262144 --reps 3 [--mode 0]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "incentive-design-mpc_b200"))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--ev", default="small")
ap.add_argument("--batch", type=int, default=262144)
ap.add_argument("--N", type=int, default=24)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--mode", type=int, default=0, help="0: test_lompc.py:34-36 prices; 1: closed-loop scale")
args = ap.parse_args()

import torch  # noqa: E402
from bench import EV_CONSTS  # noqa: E402
from chargingstation.lompc import LoMPC, LoMPCConstants  # noqa: E402

delta, theta, y_max, w_max = EV_CONSTS[args.ev]
N, B = args.N, args.batch
rng = np.random.default_rng(0)
if args.mode == 0:
    lm, lr, gam = theta * rng.random((B, 3 * N)), 3 * N * delta * rng.random(B), y_max * rng.random(B)
else:
    lm = 0.05 * theta * rng.random((B, 3 * N)) * (rng.random((B, 3 * N)) < 0.5)
    lr, gam = np.zeros(B), y_max - (0.3 + 0.2 * rng.random(B))
dev = torch.device("cuda:0")
solver = LoMPC(N, LoMPCConstants(delta, theta, y_max, w_max, args.ev))
solver.set_kernel_variant(args.variant)
lm, lr, gam = (torch.from_numpy(x).to(dev) for x in (lm, lr, gam))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for r in range(args.reps):
    e0.record()
    w, cost, info = solver.solve_lompc_batch(lm, lr, gam, return_info=True)
    e1.record()
    torch.cuda.synchronize()
    print(f"rep {r}: {e0.elapsed_time(e1):.3f} ms  {B / e0.elapsed_time(e1) / 1e3:.2f} MQP/s  "
          f"iters mean {info['iters'].double().mean():.2f} max {int(info['iters'].max())} "
          f"status max {int(info['status'].max())} kkt {float(info['kkt_res'].max()):.1e}")
