"""CPU tests of the multi-rank host logic (world_size 2, gloo): sharding of a group-sorted EV
batch, the placement of the all-reduces in the price loop and loop termination.  The compute
phases are played by an oracle-backed stand-in (tests/fake_shard_backend.py); the CUDA phases
themselves are covered by tests/test_price_gpu.py::test_sharded_*."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from chargingstation.sharded import (compute_optimal_prices_sharded,
                                     shard_bounds, shard_groups)
from oracle import lompc_oracle as orc
from oracle import price_oracle as po


def test_shard_bounds_partition():
    for B in (0, 1, 7, 64, 65537):
        for world in (1, 2, 3, 8):
            cuts = [shard_bounds(B, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == B
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_shard_groups_local_offsets():
    off = np.array([0, 5, 5, 22, 23, 32])
    y0 = np.arange(32, dtype=np.float64)
    seen = np.zeros(32, dtype=int)
    for world in (1, 2, 4):
        seen[:] = 0
        counts = np.zeros(5, dtype=int)
        for r in range(world):
            loc, y_loc, (lo, hi) = shard_groups(off, y0, r, world)
            assert loc[0] == 0 and loc[-1] == hi - lo and np.all(np.diff(loc) >= 0)
            assert np.array_equal(y_loc, y0[lo:hi])
            counts += np.diff(loc)
            for g in range(5):  # local group g really is the global group's slice on this rank
                assert np.array_equal(y_loc[loc[g]:loc[g + 1]], y0[max(off[g], lo):max(min(off[g + 1], hi), max(off[g], lo))])
            seen[lo:hi] += 1
        assert np.all(seen == 1) and np.array_equal(counts, np.diff(off))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _problem():
    o = orc.small_ev_consts()
    N = 6
    rng = np.random.default_rng(31)
    sizes = [3, 0, 4, 2]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    y0 = 0.3 + 0.05 * rng.random(off[-1])
    w_ref = o.w_max * rng.random((len(sizes), N))
    return o, N, off, y0, w_ref


def _worker(rank, world, port, out, pipelined=True, bad_y0=False):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "incentive-design-mpc_b200"), os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from fake_shard_backend import OracleShardBackend
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    o, N, off, y0, w_ref = _problem()
    if bad_y0:
        y0 = y0.copy()
        y0[-1] = o.y_max + 0.05  # owned by the LAST rank only
    loc_off, loc_y0, _ = shard_groups(off, y0, rank, world)
    G = len(off) - 1
    be = OracleShardBackend(N, o, "linear-convex")
    try:
        prices, stats = compute_optimal_prices_sharded(None, loc_off, loc_y0, w_ref, np.zeros(G), np.zeros((G, 3 * N)),
                                                       backend=be, max_iter=60, pipelined=pipelined)
        if rank == 0:
            out["prices"], out["iter"], out["total"] = prices, np.asarray(stats["iter"]), stats["total_iters"]
            out["async_calls"] = getattr(be, "async_calls", 0)
            out["local_sums_calls"] = getattr(be, "local_sums_calls", 0)
    except AssertionError:
        out[f"assert_{rank}"] = True
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("pipelined", [True, False])
def test_two_rank_gloo_price_loop_matches_single_process(pipelined):
    """Both host loops: pipelined (group_phase_async + poll, the host PIPELINE_DEPTH iterations ahead) and the
    synchronous one; same prices, same iteration counts, and the ranks leave the loop together (the run would
    hang in gloo otherwise)."""
    world = 2
    port = _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out, pipelined), nprocs=world, join=True)
        prices, iters = np.array(out["prices"]), np.array(out["iter"])
        assert (out["async_calls"] > 0) == pipelined
        assert out["local_sums_calls"] == 0  # ranks that must reduce never ask for the column sums to stay local
        assert out["total"] == max(iters[[0, 2, 3]])  # the loop ends when the last NON-EMPTY group has converged
    # reference: every group on its own, single process, the oracle's own loop
    o, N, off, y0, w_ref = _problem()
    for g in range(len(off) - 1):
        if off[g + 1] == off[g]:
            assert np.all(prices[g] == 0)
            continue
        ora = po.PriceOracle(N, o, "linear-convex")
        ora.set_charge_levels(y0[off[g]:off[g + 1]])
        lam, st = ora.compute_optimal_prices(w_ref[g], 0.0, max_iter=60)
        assert st["iter"] == iters[g]
        assert np.max(np.abs(lam - prices[g])) <= 1e-9 * max(1.0, np.max(np.abs(lam)))


@pytest.mark.timeout(300)
def test_bad_charge_level_on_one_rank_raises_on_every_rank():
    """y0 > y_max (the assert of price_solver.py:71) on an EV only the last rank owns: the statistics are
    validated AFTER their all-reduce, so both ranks raise instead of one of them hanging in the next collective."""
    world = 2
    port = _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out, True, True), nprocs=world, join=True)
        assert out.get("assert_0") and out.get("assert_1")


def test_single_process_asks_for_local_sums():
    """One process, no process group: compute_optimal_prices_sharded tells the backend that nothing has to be reduced
    (price_shard_local_sums on the CUDA backend: the group phase forms the column sums itself) - once, after begin and
    before start - and the result is the oracle's own loop."""
    from fake_shard_backend import OracleShardBackend
    o, N, off, y0, w_ref = _problem()
    G = len(off) - 1

    class Recording(OracleShardBackend):
        calls = []

        def begin(self, *a, **k):
            self.calls.append("begin")
            return super().begin(*a, **k)

        def local_sums(self):
            self.calls.append("local_sums")

        def start(self):
            self.calls.append("start")
            return super().start()

    be = Recording(N, o, "linear-convex")
    prices, stats = compute_optimal_prices_sharded(None, off, y0, w_ref, np.zeros(G), np.zeros((G, 3 * N)),
                                                   backend=be, max_iter=60)
    assert be.calls[:3] == ["begin", "local_sums", "start"] and be.calls.count("local_sums") == 1
    for g in range(G):
        if off[g + 1] == off[g]:
            continue
        ora = po.PriceOracle(N, o, "linear-convex")
        ora.set_charge_levels(y0[off[g]:off[g + 1]])
        lam, st = ora.compute_optimal_prices(w_ref[g], 0.0, max_iter=60)
        assert st["iter"] == np.asarray(stats["iter"])[g]
        assert np.max(np.abs(lam - prices[g])) <= 1e-9 * max(1.0, np.max(np.abs(lam)))
