"""Multi-GPU price loop: EVs sharded over ranks, one process per GPU.

The lower-level QPs are independent, so each rank solves its own slice of the
EV batch.  The only coupling is inside ``_get_w_err`` (reference
price_solver.py:196-214): the mean response of a (station, EV-type, partition)
group, ``w_avg = sum_i w_i / nEVs`` (:205,210), and ``set_charge_levels``' group
statistics (:66-77).  When a group's EVs straddle ranks those become ONE
``all_reduce`` each: MIN/MAX/SUM of ``[G]`` statistics once per solve, SUM of the
``[G, N]`` fp64 partial sums once per price iteration (393 KB for BASELINE config 3),
issued on the same stream as the kernels.  The convergence test, the price step
and the gamma_sc solve then run replicated on every rank (identical inputs ->
identical prices, no broadcast).  ``torch.distributed`` is the plumbing (NCCL on
GPUs); the compute is the C ABI's ``price_shard_*`` phases.

The loop is PIPELINED: the host never synchronises inside it.  An iteration is enqueued as
kernels + the all-reduce on the stream; the number of still-active groups comes back through a
pinned ring the device writes (``price_shard_group_phase_async`` / ``price_shard_poll``), and the
host only makes sure it is never more than ``PIPELINE_DEPTH`` iterations ahead of the GPU.
Converged groups are skipped on the device, so the few iterations enqueued beyond convergence
are no-ops; every rank sees the same counts (the group phase is replicated), so every rank
leaves the loop at the same iteration and the all-reduces stay matched.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from chargingstation import _native
from chargingstation.settings import (MAX_PRICE_SOLVER_ITERATIONS,
                                      PRICE_SOLVER_TOL_TYPE)


PIPELINE_DEPTH = 4  # iterations the host may run ahead of the device (< the 64-slot ring of the C ABI)


def shard_bounds(B: int, rank: int, world: int) -> tuple[int, int]:
    """Block partition of EV indices [0, B) over ranks (sizes differ by at most one)."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_groups(group_off: np.ndarray, y0: np.ndarray, rank: int, world: int):
    """Local view of a group-sorted EV batch: (local group_off [G+1], local y0, (lo, hi)).
    Groups that have no EV on this rank get an empty local range."""
    group_off = np.asarray(group_off, dtype=np.int64)
    lo, hi = shard_bounds(int(group_off[-1]), rank, world)
    local_off = (np.clip(group_off, lo, hi) - lo).astype(np.int32)
    return local_off, np.ascontiguousarray(y0[lo:hi]), (lo, hi)


class CudaShardBackend:
    """The five phases on one GPU through the C ABI (``price_shard_*``).

    With more than one rank (one GPU per rank, NVLink peer access) the per-iteration aggregate does not go
    through an all-reduce call at all: every rank's partial sums live in a CUDA-IPC region all ranks have mapped,
    and the group phase adds them up in rank order straight out of the peers' memory
    (``price_shard_attach_peers``; set ``LOMPC_SHARD_PEER=0`` to keep the ``dist.all_reduce``).  The regions are
    set up once per PriceSolver (one ``all_gather`` of the 64-byte IPC handles) and reused."""

    def __init__(self, price_solver, process_group=None):
        import torch
        self.torch = torch
        self.ps = price_solver
        self.lib = _native.load()
        self.dev = torch.device("cuda", price_solver.device)
        self.group = process_group

    def _ensure_peers(self, G: int) -> None:
        """Collective: every rank calls it with the same G."""
        import os
        import torch.distributed as dist
        torch, lib, ps = self.torch, self.lib, self.ps
        if not (dist.is_available() and dist.is_initialized()) or os.environ.get("LOMPC_SHARD_PEER", "1") == "0":
            return
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if world <= 1 or world > 8:
            return
        need = 1024 + 2 * G * ps.N * 8
        st = getattr(ps, "_peer_state", None)
        if st is not None and st["bytes"] >= need and st["world"] == world:
            return
        if st is not None:  # grow: detach, unmap the peers' regions, free the own one
            _native.raise_for(lib.price_shard_attach_peers(ps._h, 0, 0, None, 0))
            torch.cuda.synchronize(self.dev)
            dist.barrier(self.group)
            for r, p in enumerate(st["ptrs"]):
                if r != rank:
                    lib.lompc_ipc_close(ps.device, p)
            dist.barrier(self.group)
            lib.lompc_ipc_free(ps.device, st["ptrs"][rank])
        nbytes = max(need, 4 << 20)
        own = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        _native.raise_for(lib.lompc_ipc_alloc(ps.device, nbytes, C.byref(own), handle))
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=self.dev)
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine, group=self.group)
        ptrs = []
        for r in range(world):
            if r == rank:
                ptrs.append(own.value)
                continue
            raw = (C.c_ubyte * 64)(*gathered[r].cpu().tolist())
            p = C.c_void_p()
            _native.raise_for(lib.lompc_ipc_open(ps.device, raw, C.byref(p)))
            ptrs.append(p.value)
        arr = (C.c_void_p * world)(*ptrs)
        _native.raise_for(lib.price_shard_attach_peers(ps._h, rank, world, arr, nbytes))
        ps._peer_state = {"bytes": nbytes, "world": world, "ptrs": ptrs}
        dist.barrier(self.group)  # nobody raises a flag in a region that is not mapped everywhere yet

    def uses_peers(self) -> bool:
        return bool(self.lib.price_shard_uses_peers(self.ps._h))

    def _t(self, x, dtype=None):
        t = self.torch.from_numpy(np.ascontiguousarray(x))
        return (t if dtype is None else t.to(dtype)).to(self.dev).contiguous()

    def _stream(self):
        return self.torch.cuda.current_stream(self.dev).cuda_stream

    def begin(self, group_off, y0, w_ref, lmbd_r, prev_prices, max_iter, history):
        torch, ps = self.torch, self.ps
        N = ps.N
        G = len(group_off) - 1
        self.G, self.B = G, int(group_off[-1])
        self._ensure_peers(G)
        f64 = dict(dtype=torch.float64, device=self.dev)
        self.off = self._t(np.asarray(group_off, dtype=np.int32))
        self.y0 = self._t(np.asarray(y0, dtype=np.float64)) if self.B > 0 else torch.zeros((1,), **f64)
        self.w_ref = self._t(np.asarray(w_ref, dtype=np.float64).reshape(G, N))
        self.lmbd_r = self._t(np.asarray(lmbd_r, dtype=np.float64).reshape(G))
        self.prices = self._t(np.asarray(prev_prices, dtype=np.float64).reshape(G, 3 * N).copy())
        self.iters = torch.zeros((G,), dtype=torch.int32, device=self.dev)
        self.stat_min, self.stat_max = torch.zeros((G,), **f64), torch.zeros((G,), **f64)
        self.stat_sum, self.stat_cnt = torch.zeros((G,), **f64), torch.zeros((G,), **f64)
        self.w_sum, self.err_max = torch.zeros((G, N), **f64), torch.zeros((G,), **f64)
        cap = max_iter if history else 0
        self.hist_ac = torch.zeros((G, max(cap, 1)), **f64) if history else None
        self.hist_pred = torch.zeros((G, max(cap, 1)), **f64) if history else None
        rc = self.lib.price_shard_begin(
            ps._h, G, self.B, self.off.data_ptr(), self.y0.data_ptr(), self.w_ref.data_ptr(),
            self.lmbd_r.data_ptr(), ps.r, int(max_iter), 1 if PRICE_SOLVER_TOL_TYPE == "max" else 0,
            float(ps.eps_reg), float(ps.eps_tol), self.prices.data_ptr(), self.iters.data_ptr(),
            self.stat_min.data_ptr(), self.stat_max.data_ptr(), self.stat_sum.data_ptr(), self.stat_cnt.data_ptr(),
            self.w_sum.data_ptr(), self.err_max.data_ptr(),
            self.hist_ac.data_ptr() if history else None, self.hist_pred.data_ptr() if history else None, cap,
            self._stream())
        _native.raise_for(rc)
        return self.stat_min, self.stat_max, self.stat_sum, self.stat_cnt

    def local_sums(self) -> None:
        """This rank holds every EV (one process): the group phase forms the column sums itself, ``ev_phase``
        launches no column-sum kernel and leaves ``w_sum`` / ``err_max`` untouched.  Call after ``begin``."""
        _native.raise_for(self.lib.price_shard_local_sums(self.ps._h, 1))

    def start(self):
        _native.raise_for(self.lib.price_shard_start(self.ps._h, self._stream()))

    def ev_phase(self):
        _native.raise_for(self.lib.price_shard_ev_phase(self.ps._h, self._stream()))
        return self.w_sum, self.err_max

    def group_phase(self, it: int) -> int:
        n = C.c_int32(0)
        _native.raise_for(self.lib.price_shard_group_phase(self.ps._h, int(it), C.byref(n), self._stream()))
        return int(n.value)

    def group_phase_async(self, it: int) -> None:
        _native.raise_for(self.lib.price_shard_group_phase_async(self.ps._h, int(it), self._stream()))

    def poll(self, it: int, wait: bool = True):
        """Active groups after iteration ``it`` (None if not published yet and ``wait`` is False)."""
        n = C.c_int32(0)
        rc = self.lib.price_shard_poll(self.ps._h, int(it), 1 if wait else 0, C.byref(n))
        if rc < 0:
            _native.raise_for(rc)
        return int(n.value) if rc == 1 else None

    def finish(self, history):
        torch = self.torch
        pre = torch.zeros((self.G,), dtype=torch.float64, device=self.dev)
        post = torch.zeros((self.G,), dtype=torch.float64, device=self.dev)
        w_k = torch.zeros((self.G, self.ps.N), dtype=torch.float64, device=self.dev)
        _native.raise_for(self.lib.price_shard_finish(self.ps._h, pre.data_ptr(), post.data_ptr(), w_k.data_ptr(),
                                                      self._stream()))
        stats = {"iter": self.iters.cpu().numpy(), "price_before_reg": pre.cpu().numpy(),
                 "price_after_reg": post.cpu().numpy(), "w_k": w_k.cpu().numpy()}
        if history:
            stats["hist_ac"], stats["hist_pred"] = self.hist_ac.cpu().numpy(), self.hist_pred.cpu().numpy()
        return self.prices.cpu().numpy(), stats


def compute_optimal_prices_sharded(price_solver, group_off_local, y0_local, w_ref, lmbd_r, prev_prices,
                                   process_group=None, backend=None, history: bool = False,
                                   max_iter: int = MAX_PRICE_SOLVER_ITERATIONS, pipelined: bool = True):
    """``PriceSolver.compute_optimal_prices`` for G groups whose EVs are sharded over the ranks
    of ``process_group`` (None: single process, no collective).  Every rank passes the SAME
    w_ref / lmbd_r / prev_prices and its LOCAL (group_off, y0) from ``shard_groups``; every rank
    returns the same (prices [G, 3N], stats)."""
    import torch
    import torch.distributed as dist

    be = backend if backend is not None else CudaShardBackend(price_solver, process_group)
    distributed = process_group is not None or (dist.is_available() and dist.is_initialized()
                                                and dist.get_world_size() > 1)
    smin, smax, ssum, scnt = be.begin(group_off_local, y0_local, w_ref, lmbd_r, prev_prices, max_iter, history)
    if not distributed and hasattr(be, "local_sums"):
        be.local_sums()  # nothing to reduce: one launch less per iteration
    if distributed:  # set_charge_levels across ranks (price_solver.py:66-77)
        dist.all_reduce(smin, op=dist.ReduceOp.MIN, group=process_group)
        dist.all_reduce(smax, op=dist.ReduceOp.MAX, group=process_group)
        dist.all_reduce(ssum, op=dist.ReduceOp.SUM, group=process_group)
        dist.all_reduce(scnt, op=dist.ReduceOp.SUM, group=process_group)
    be.start()  # validates the REDUCED statistics: every rank raises together (price_solver.py:71)

    peer_exchange = bool(getattr(be, "uses_peers", lambda: False)())  # the sums travel over NVLink peer memory

    def reduce_sums(w_sum, err_max):  # w_avg += w_i over all ranks (price_solver.py:205)
        if distributed and not peer_exchange:
            dist.all_reduce(w_sum, op=dist.ReduceOp.SUM, group=process_group)
            if PRICE_SOLVER_TOL_TYPE == "max":
                dist.all_reduce(err_max, op=dist.ReduceOp.MAX, group=process_group)

    total = 0
    if pipelined and hasattr(be, "group_phase_async"):
        # no host synchronisation inside the loop: iteration `it` is enqueued while the device may still be
        # PIPELINE_DEPTH iterations behind; the loop ends at the first iteration whose published count is 0
        done_at = None
        for it in range(max_iter):
            reduce_sums(*be.ev_phase())
            be.group_phase_async(it)
            j = it - PIPELINE_DEPTH
            if j >= 0 and be.poll(j, wait=True) == 0:
                done_at = j
                break
        if done_at is None:  # drain the iterations still in flight
            last = min(max_iter, it + 1)
            for j in range(max(0, last - PIPELINE_DEPTH), last):
                if be.poll(j, wait=True) == 0:
                    done_at = j
                    break
        total = done_at if done_at is not None else max_iter
    else:
        for it in range(max_iter):
            reduce_sums(*be.ev_phase())
            total = it
            if be.group_phase(it) == 0:
                break
            total = it + 1
    prices, stats = be.finish(history)
    stats["total_iters"] = total
    return prices, stats
