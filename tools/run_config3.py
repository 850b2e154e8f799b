"""BASELINE config 3: a 65,536-EV station/scenario batch (2,048 groups x 32 EVs, half small-EV and
half large-EV groups, N = 24) sharded by EV index over the ranks, groups deliberately straddling
ranks, with the aggregate-load all-reduce (NCCL over NVLink) inside the price loop.

    python tools/run_config3.py                                    # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/run_config3.py                   # 2 GPUs
Prints one JSON line from rank 0: iterations, wall time, QP solves/s, and a checksum that must
not depend on the number of ranks beyond summation-order noise."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "incentive-design-mpc_b200"))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import EV_CONSTS  # noqa: E402
from chargingstation.lompc import LoMPCConstants  # noqa: E402
from chargingstation.price_solver import PriceSolver  # noqa: E402
from chargingstation.sharded import compute_optimal_prices_sharded, shard_groups  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

N, G_PER_TYPE, EVS = 24, 1024, 32
MAX_IT = int(os.environ.get("CONFIG3_MAX_ITER", "200"))
rng = np.random.default_rng(3)
res = {}
t_total, qp_total = 0.0, 0
for ev in ("small", "large"):
    delta, theta, y_max, w_max = EV_CONSTS[ev]
    off = (np.arange(G_PER_TYPE + 1) * EVS).astype(np.int64)
    y0 = 0.3 + 0.2 * rng.random(off[-1])              # charging_station.py:95-100
    y0.sort()                                         # partitions group EVs of similar SoC (charging_station.py:111-116)
    w_ref = 0.5 * w_max * rng.random((G_PER_TYPE, N))  # test_price_solver.py:34 (scaled to stay reachable)
    ps = PriceSolver(N, LoMPCConstants(delta, theta, y_max, w_max, ev), "linear-convex", device=local_rank)
    loc_off, loc_y0, (lo, hi) = shard_groups(off, y0, rank, world)
    args = (ps, loc_off, loc_y0, w_ref, np.zeros(G_PER_TYPE), np.zeros((G_PER_TYPE, 3 * N)))
    compute_optimal_prices_sharded(*args, max_iter=3)  # warm-up (allocations, NCCL channels)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    prices, stats = compute_optimal_prices_sharded(*args, max_iter=MAX_IT)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    it = stats["iter"]
    # EV QPs solved = sum over groups of (iterations run while active + 1) * EVs; + 2 group QPs per iteration
    qps = int(np.sum((np.minimum(it, MAX_IT - 1) + 1) * EVS + 2 * (np.minimum(it, MAX_IT - 1) + 1)))
    res[ev] = {"seconds": dt, "total_iters": int(stats["total_iters"]), "iters_mean": float(it.mean()),
               "iters_max": int(it.max()), "converged_groups": int(np.sum(it < MAX_IT - 1)), "qp_solves": qps,
               "qp_per_s": qps / dt, "price_checksum": float(np.sum(prices)),
               "local_evs": int(hi - lo)}
    t_total += dt
    qp_total += qps
if rank == 0:
    print(json.dumps({"config": "BASELINE config 3 (65,536 EVs, 2,048 groups x 32, N=24, linear-convex prices)",
                      "n_gpus": world, "max_iter": MAX_IT, "seconds": t_total, "qp_solves": qp_total,
                      "qp_per_s": qp_total / t_total, "allreduce_bytes_per_iter": G_PER_TYPE * N * 8,
                      "per_type": res}))
if world > 1:
    dist.destroy_process_group()
