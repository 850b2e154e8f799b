"""GPU parity tests of the warp-cooperative K1 (csrc/lompc_solve_warp.cuh: one QP per lane group,
time-parallel sweeps) and of the solve set (``lompc_set_*``: several LoMPC objects, one launch, one copy
each way) through the C ABI: against the exact active-set oracle, against the one-QP-per-thread kernels
(same algorithm, same decisions => same iteration counts) and through the error conventions of
lompc.py:78-90."""
import numpy as np
import pytest

from oracle import lompc_oracle as orc

pytestmark = pytest.mark.gpu

W_RTOL = 1e-9   # relative to w_max   (north-star bar: 1e-5)
C_RTOL = 1e-10  # relative to max(1,|cost|) (north-star bar: 1e-6)


def _consts(ev):
    from chargingstation.lompc import LoMPCConstants
    o = orc.small_ev_consts() if ev == "small" else orc.large_ev_consts()
    return o, LoMPCConstants(o.delta, o.theta, o.y_max, o.w_max, o.ev_type)


def _regimes(o, N, B, seed):
    """Thirds of the batch: test_lompc.py:34-36 prices, closed-loop-scale sparse prices, linear prices."""
    rng = np.random.default_rng(seed)
    th = o.theta
    lm = np.zeros((B, 3 * N))
    lr = np.zeros(B)
    gam = o.y_max * rng.random(B)
    a, b = B // 3, 2 * B // 3
    lm[:a] = th * rng.random((a, 3 * N))
    lr[:a] = 3 * N * o.delta * rng.random(a)
    lm[a:b] = 0.05 * th * rng.random((b - a, 3 * N)) * (rng.random((b - a, 3 * N)) < 0.5)
    lm[b:, :2 * N] = 0.05 * th * rng.random((B - b, 2 * N))
    return lm, lr, gam


@pytest.mark.parametrize("ev", ["small", "large"])
@pytest.mark.parametrize("N,variant", [(12, 8), (24, 8), (48, 8), (96, 8)])
def test_warp_kernel_against_oracle_and_thread_kernel(ev, N, variant):
    from chargingstation.lompc import LoMPC
    o, c = _consts(ev)
    B = 203  # not a multiple of the QPs per warp: the last warp has idle groups
    lm, lr, gam = _regimes(o, N, B, 300 + N)
    solver = LoMPC(N, c)
    solver.set_kernel_variant(variant)
    w, cost, info = solver.solve_lompc_batch(lm, lr, gam, return_info=True)
    assert np.all(info["status"] == 0), np.nonzero(info["status"])[0]
    assert info["kkt_res"].max() <= 1e-10
    assert w.min() >= 0.0 and w.max() <= o.w_max
    # the one-QP-per-thread kernel of the same horizon (register kernel for N = 12, 24; any-N otherwise)
    ref = LoMPC(N, c)
    ref.set_kernel_variant(4 if N in (12, 24) else 1)
    w_t, cost_t, info_t = ref.solve_lompc_batch(lm, lr, gam, return_info=True)
    assert np.max(np.abs(w - w_t)) <= 1e-12 * o.w_max
    assert np.max(np.abs(cost - cost_t) / np.maximum(1, np.abs(cost_t))) <= 1e-13
    # same decisions as the register kernel => the iteration counts agree except where a rounding-level tie
    # falls the other way (the any-N kernel has no optimistic phase: its counts differ)
    if N in (12, 24):
        assert np.mean(info["iters"] != info_t["iters"]) <= 0.02, (info["iters"], info_t["iters"])
    for b in range(0, B, 7 if N <= 24 else 29):
        wo, co, _ = orc.solve_active_set(N, o, lm[b], lr[b], gam[b])
        assert np.max(np.abs(w[b] - wo)) <= W_RTOL * o.w_max, (b, info["iters"][b])
        assert abs(cost[b] - co) <= C_RTOL * max(1, abs(co))
        viol, dist = orc.kkt_certificate(N, o, w[b], lm[b], lr[b], gam[b])
        assert dist <= 1e-8


def test_warp_kernel_is_the_small_batch_default_and_deterministic():
    from chargingstation import _native
    from chargingstation.lompc import LoMPC
    o, c = _consts("large")
    N, B = 24, 1000
    lm, lr, gam = _regimes(o, N, B, 11)
    auto, forced = LoMPC(N, c), LoMPC(N, c)
    forced.set_kernel_variant(8)
    w1, c1 = auto.solve_lompc_batch(lm, lr, gam)
    w2, c2 = forced.solve_lompc_batch(lm, lr, gam)
    assert np.array_equal(w1, w2) and np.array_equal(c1, c2)
    # a QP's result does not depend on its position in the batch (which lane group / warp solves it)
    perm = np.random.default_rng(1).permutation(B)
    w3, c3 = auto.solve_lompc_batch(lm[perm], lr[perm], gam[perm])
    assert np.array_equal(w3, w1[perm]) and np.array_equal(c3, c1[perm])
    # broadcast prices (lmbd_stride = 0, price_solver.py:203-204)
    w4, c4 = auto.solve_lompc_batch(lm[0], lr[0], gam)
    w5, c5 = auto.solve_lompc_batch(np.tile(lm[0], (B, 1)), np.full(B, lr[0]), gam)
    assert np.array_equal(w4, w5) and np.array_equal(c4, c5)
    assert _native.load().lompc_launch_count() > 0


def test_warp_kernel_edge_cases_and_status():
    from chargingstation.lompc import LoMPC
    for ev in ("small", "large"):
        o, c = _consts(ev)
        N = 24
        solver = LoMPC(N, c)
        solver.set_kernel_variant(8)
        z = np.zeros(3 * N)
        w, cost = solver.solve_lompc(z, 0.0, 0.0)
        assert np.all(w == 0.0) and cost == 0.0
        w, cost = solver.solve_lompc(z, 0.0, o.y_max)
        wo, co, _ = orc.solve_active_set(N, o, z, 0.0, o.y_max)
        assert np.max(np.abs(w - wo)) <= W_RTOL * o.w_max and abs(cost - co) <= C_RTOL * abs(co)
        big = np.concatenate([1e4 * np.ones(N), np.zeros(2 * N)])
        assert np.all(solver.solve_lompc(big, 0.0, 0.5)[0] == 0.0)
        neg = np.concatenate([np.zeros(N), 1e4 * np.ones(N), np.zeros(N)])
        assert np.all(solver.solve_lompc(neg, 0.0, 0.5)[0] == o.w_max)
        # prices of the size the closed-form regulariser can return (lmbd3 ~ 1e11 where w ~ 0)
        rng = np.random.default_rng(4)
        lm = 0.05 * o.theta * rng.random((8, 3 * N))
        lm[:, 2 * N::3] = 1e11
        gam = o.y_max * rng.random(8)
        w, cost, info = solver.solve_lompc_batch(lm, np.zeros(8), gam, return_info=True)
        assert np.all(info["status"] == 0)
        thread = LoMPC(N, c)
        thread.set_kernel_variant(4)
        w_t, cost_t = thread.solve_lompc_batch(lm, np.zeros(8), gam)
        assert np.max(np.abs(w - w_t)) <= 1e-11 * o.w_max
        assert np.max(np.abs(cost - cost_t) / np.maximum(1, np.abs(cost_t))) <= 1e-12
        # per-QP status and the reference's exceptions (lompc.py:78-90)
        lm = o.theta * rng.random((5, 3 * N))
        gam = np.array([0.1, 0.2, o.y_max + 0.01, 0.3, 0.4])
        with pytest.raises(AssertionError):
            solver.solve_lompc_batch(lm, 0.0, gam)
        lm2 = lm.copy()
        lm2[3, 7] = -1.0
        with pytest.raises(ValueError):
            solver.solve_lompc_batch(lm2, 0.0, np.full(5, 0.2))
        for variant in (0, 1, 4, 8):          # NaN parameters are invalid in every kernel (auto, registers, thread, warp)
            lm2[3, 7] = np.nan
            solver.set_kernel_variant(variant)
            with pytest.raises(ValueError):
                solver.solve_lompc_batch(lm2, 0.0, np.full(5, 0.2))
            lm2[3, 7] = 0.1
            with pytest.raises((ValueError, AssertionError)):
                solver.solve_lompc_batch(lm2, 0.0, np.array([0.2, 0.2, np.nan, 0.2, 0.2]))
        solver.set_kernel_variant(0)
        solver.set_solver_options(max_iter=1)
        with pytest.raises(RuntimeError):
            solver.solve_lompc_batch(lm, 0.0, np.full(5, 0.5))


def _make_set(N, Bs, Bl):
    from chargingstation.lompc import LoMPC, LoMPCSet
    os_, cs_ = _consts("small")
    ol, cl = _consts("large")
    small, large = LoMPC(N, cs_), LoMPC(N, cl)
    return (os_, ol), (small, large), LoMPCSet([small, large], [Bs, Bl])


@pytest.mark.parametrize("mapped", ["1", "0"])
@pytest.mark.parametrize("N", [12, 24])
def test_solve_set_matches_per_object_solves(N, mapped, monkeypatch):
    """Both transfer modes of the host round trip: zero-copy (the kernel reads / writes the pinned host blocks) and
    staged (H2D copy, launch, D2H copy)."""
    monkeypatch.setenv("LOMPC_SET_MAPPED", mapped)
    (os_, ol), (small, large), sset = _make_set(N, 301, 217)
    data = []
    for i, o in enumerate((os_, ol)):
        lm, lr, gam = _regimes(o, N, sset.batch_sizes[i], 40 + i)
        sset.lmbd[i][:] = lm
        sset.lmbd_r[i][:] = lr
        sset.gamma[i][:] = gam
        data.append((lm, lr, gam))
    sset.solve(info=True)
    for i, (solver, o) in enumerate(((small, os_), (large, ol))):
        lm, lr, gam = data[i]
        w, cost, info = solver.solve_lompc_batch(lm, lr, gam, return_info=True)
        assert np.array_equal(sset.w[i], w) and np.array_equal(sset.cost[i], cost)
        assert np.array_equal(sset.status[i], info["status"]) and np.array_equal(sset.iters[i], info["iters"])
        assert np.array_equal(sset.kkt_res[i], info["kkt_res"])
        for b in range(0, len(gam), 23):
            wo, co, _ = orc.solve_active_set(N, o, lm[b], lr[b], gam[b])
            assert np.max(np.abs(sset.w[i][b] - wo)) <= W_RTOL * o.w_max
            assert abs(sset.cost[i][b] - co) <= C_RTOL * max(1, abs(co))
    # a second call with new inputs in the same views (the captured graph is replayed)
    sset.gamma[0][:] = 0.5 * data[0][2]
    w_old = sset.w[0].copy()
    sset.solve()
    w, cost = small.solve_lompc_batch(data[0][0], data[0][1], 0.5 * data[0][2])
    assert np.array_equal(sset.w[0], w) and not np.array_equal(w, w_old)
    assert sset.h2d_bytes >= 8 * (3 * N + 2) * (301 + 217) and sset.d2h_bytes >= 8 * (N + 1) * (301 + 217)


@pytest.mark.parametrize("mapped", ["1", "0"])
def test_solve_set_error_conventions_and_async(mapped, monkeypatch):
    monkeypatch.setenv("LOMPC_SET_MAPPED", mapped)
    N = 24
    (os_, ol), (small, large), sset = _make_set(N, 64, 64)
    for i, o in enumerate((os_, ol)):
        lm, lr, gam = _regimes(o, N, 64, 7 + i)
        sset.lmbd[i][:], sset.lmbd_r[i][:], sset.gamma[i][:] = lm, lr, gam
    sset.solve_async()
    sset.wait()
    good = sset.w[1].copy()
    sset.gamma[1][5] = ol.y_max + 0.01       # lompc.py:87
    with pytest.raises(AssertionError):
        sset.solve()
    sset.gamma[1][5] = 0.2
    sset.lmbd[0][9, 3] = -0.5                # nonneg cv.Parameter, lompc.py:78-82
    with pytest.raises(ValueError):
        sset.solve()
    sset.lmbd[0][9, 3] = np.nan              # NaN is not a nonneg value either (cvxpy rejects it the same way)
    with pytest.raises(ValueError):
        sset.solve()
    sset.lmbd[0][9, 3] = 0.5
    sset.lmbd_r[1][2] = np.nan
    with pytest.raises(ValueError):
        sset.solve()
    sset.lmbd_r[1][2] = 0.0
    sset.solve()                             # the failure of an earlier call does not stick
    assert sset.w[1].shape == good.shape
    small.set_solver_options(max_iter=1)
    with pytest.raises(RuntimeError):
        sset.solve()


def test_solve_set_large_batches_take_the_thread_kernels():
    """Beyond 16,384 QPs (N = 12, 24) the set launches one one-QP-per-thread kernel per segment and reduces the status
    on the device; results equal the per-object path."""
    N = 24
    (os_, ol), (small, large), sset = _make_set(N, 70000, 70000)
    for i, o in enumerate((os_, ol)):
        rng = np.random.default_rng(60 + i)
        sset.lmbd[i][:] = o.theta * rng.random((70000, 3 * N))
        sset.lmbd_r[i][:] = 3 * N * o.delta * rng.random(70000)
        sset.gamma[i][:] = o.y_max * rng.random(70000)
    sset.solve()
    for i, (solver, o) in enumerate(((small, os_), (large, ol))):
        w, cost = solver.solve_lompc_batch(sset.lmbd[i].copy(), sset.lmbd_r[i].copy(), sset.gamma[i].copy())
        assert np.array_equal(sset.w[i], w) and np.array_equal(sset.cost[i], cost)
    sset.gamma[0][123] = os_.y_max + 0.5
    with pytest.raises(AssertionError):
        sset.solve()


def test_solve_set_device_entry_point_in_a_cuda_graph():
    import torch
    N = 24
    (os_, ol), (small, large), sset = _make_set(N, 512, 512)
    for i, o in enumerate((os_, ol)):
        rng = np.random.default_rng(2 + i)
        sset.lmbd[i][:] = o.theta * rng.random((512, 3 * N))
        sset.lmbd_r[i][:] = 3 * N * o.delta * rng.random(512)
        sset.gamma[i][:] = o.y_max * rng.random(512)
    sset.solve()
    w_host = [w.copy() for w in sset.w]
    for w in sset.w:
        w[:] = 0.0
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        sset.upload(stream.cuda_stream)
        stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            sset.solve_dev(torch.cuda.current_stream().cuda_stream)
        g.replay()
        sset.download(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    for i in range(2):
        assert np.array_equal(sset.w[i], w_host[i])
