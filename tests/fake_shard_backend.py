"""Oracle-backed stand-in for ``chargingstation.sharded.CudaShardBackend`` (TEST ONLY): the same
five phases on CPU tensors, so that the multi-rank host logic (sharding, placement of the
all-reduces, loop termination) can be exercised with gloo on a machine without GPUs."""
import numpy as np
import torch

from oracle import lompc_oracle as orc
from oracle import price_oracle as po


class OracleShardBackend:
    def __init__(self, N, consts, price_type):
        self.ora = po.PriceOracle(N, consts, price_type)
        self.N, self.consts = N, consts

    def begin(self, group_off, y0, w_ref, lmbd_r, prev_prices, max_iter, history):
        N = self.N
        self.off = np.asarray(group_off)
        G = len(self.off) - 1
        self.G, self.y0 = G, np.asarray(y0)
        self.w_ref = np.asarray(w_ref).reshape(G, N)
        self.lmbd_r = np.asarray(lmbd_r).reshape(G)
        self.prices = np.asarray(prev_prices, dtype=np.float64).reshape(G, 3 * N).copy()
        self.max_iter = max_iter
        smin, smax = torch.full((G,), 1e300, dtype=torch.float64), torch.full((G,), -1e300, dtype=torch.float64)
        ssum, scnt = torch.zeros(G, dtype=torch.float64), torch.zeros(G, dtype=torch.float64)
        for g in range(G):
            y = self.y0[self.off[g]:self.off[g + 1]]
            if len(y):
                smin[g], smax[g], ssum[g], scnt[g] = y.min(), y.max(), y.sum(), len(y)
        self.stats = (smin, smax, ssum, scnt)
        self.w_sum = torch.zeros((G, N), dtype=torch.float64)
        self.err_max = torch.zeros(G, dtype=torch.float64)
        return self.stats

    def start(self):
        smin, smax, ssum, scnt = (t.numpy() for t in self.stats)
        G, N, c = self.G, self.N, self.consts
        # the assert of price_solver.py:71 on the REDUCED statistics: every rank raises together
        assert np.all((scnt <= 0) | ((smin >= 0) & (smax <= c.y_max)))
        self.cnt = scnt.copy()
        self.skip = scnt <= 0
        self.y0_rng = np.where(self.skip, 0, (smax - smin) / 2)
        self.gamma_sc = np.where(self.skip, 0, c.y_max - (smax + smin) / 2)
        self.iters = np.full(G, self.max_iter - 1)
        self.w_k = np.zeros((G, N))
        self.dual = np.zeros(G)
        self.first = np.ones(G, dtype=bool)
        for g in range(G):
            if not self.skip[g]:
                self.w_k[g], self.dual[g] = orc.solve_lompc(N, c, self.prices[g], self.lmbd_r[g], self.gamma_sc[g])

    def ev_phase(self):
        N, c = self.N, self.consts
        self.w_sum.zero_()
        for g in range(self.G):
            if self.skip[g]:
                continue
            for b in range(self.off[g], self.off[g + 1]):
                w, _ = orc.solve_lompc(N, c, self.prices[g], self.lmbd_r[g], c.y_max - self.y0[b])
                self.w_sum[g] += torch.from_numpy(w)
        return self.w_sum, self.err_max

    def group_phase(self, it):
        ora, N = self.ora, self.N
        r = ora.r
        active = 0
        for g in range(self.G):
            if self.skip[g]:
                continue
            w_avg = self.w_sum[g].numpy() / self.cnt[g]
            A_bar, A_bar_inv = ora._metric(self.lmbd_r[g])
            v = w_avg - self.w_ref[g]
            err = np.sqrt(v @ A_bar @ v)
            tol = np.sqrt(N) * self.y0_rng[g] + ora.eps_tol
            if err <= tol:
                self.skip[g], self.iters[g] = True, it
                continue
            active += 1
            nxt, _ = ora.price_gradient_descent_step(A_bar_inv, self.w_ref[g], self.w_k[g], self.prices[g, :r])
            self.prices[g, :r] = nxt
            self.w_k[g], self.dual[g] = orc.solve_lompc(N, self.consts, self.prices[g], self.lmbd_r[g],
                                                         self.gamma_sc[g])
        return active

    def local_sums(self):  # (CudaShardBackend: price_shard_local_sums; here the sums are always formed in ev_phase)
        self.local_sums_calls = getattr(self, "local_sums_calls", 0) + 1

    # the pipelined interface of CudaShardBackend: "enqueue" = run now, the count is published per iteration
    def group_phase_async(self, it):
        if not hasattr(self, "_published"):
            self._published = {}
        self._published[it] = self.group_phase(it)
        self.async_calls = getattr(self, "async_calls", 0) + 1

    def poll(self, it, wait=True):
        return self._published.get(it)

    def finish(self, history):
        ora = self.ora
        pre, post = np.zeros(self.G), np.zeros(self.G)
        for g in range(self.G):
            if self.cnt[g] <= 0:
                continue
            pre[g] = ora.phi(self.w_k[g]) @ self.prices[g]
            self.prices[g, :ora.r] = po.regularize_closed_form(self.N, self.consts, ora.r, self.w_k[g],
                                                               self.prices[g, :ora.r])
            post[g] = ora.phi(self.w_k[g]) @ self.prices[g]
        return self.prices, {"iter": self.iters, "price_before_reg": pre, "price_after_reg": post, "w_k": self.w_k}
