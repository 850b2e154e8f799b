"""GPU parity tests of K1 (batched LoMPC solve) through the C ABI, against the
CPU oracle and the committed golden vectors.  Tolerances (north-star):
w <= 1e-5 relative to w_max, cost <= 1e-6 relative - the kernel is an exact
active-set method, so the tests actually hold it to 1e-9 / 1e-10."""
import os

import numpy as np
import pytest

from oracle import lompc_oracle as orc

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "lompc_golden.npz")
W_RTOL = 1e-9   # relative to w_max   (north-star bar: 1e-5)
C_RTOL = 1e-10  # relative to max(1,|cost|) (north-star bar: 1e-6)


def _consts(ev):
    from chargingstation.lompc import LoMPCConstants
    o = orc.small_ev_consts() if ev == "small" else orc.large_ev_consts()
    return o, LoMPCConstants(o.delta, o.theta, o.y_max, o.w_max, o.ev_type)


@pytest.mark.parametrize("ev", ["small", "large"])
@pytest.mark.parametrize("N", [12, 24])
def test_golden_vectors(ev, N):
    from chargingstation.lompc import LoMPC
    o, c = _consts(ev)
    z = np.load(GOLDEN)
    solver = LoMPC(N, c)
    for mode in range(4):
        key = f"{ev}_N{N}_m{mode}"
        w, cost, info = solver.solve_lompc_batch(z[key + "_lmbd"], z[key + "_lmbd_r"], z[key + "_gamma"],
                                                 return_info=True)
        assert np.all(info["status"] == 0)
        assert np.max(np.abs(w - z[key + "_w"])) <= W_RTOL * o.w_max, key
        assert np.max(np.abs(cost - z[key + "_cost"]) / np.maximum(1, np.abs(z[key + "_cost"]))) <= C_RTOL, key
        assert info["kkt_res"].max() <= 1e-10


@pytest.mark.parametrize("ev", ["small", "large"])
@pytest.mark.parametrize("N", [5, 12, 24, 48, 96])
def test_random_against_oracle(ev, N):
    """Seeded random inputs as test_lompc.py:34-36 plus the closed-loop regimes."""
    from chargingstation.lompc import LoMPC
    o, c = _consts(ev)
    rng = np.random.default_rng(100 + N)
    B = 96
    solver = LoMPC(N, c)
    th = o.theta
    lm = np.zeros((B, 3 * N))
    lr = np.zeros(B)
    gam = o.y_max * rng.random(B)
    lm[:32] = th * rng.random((32, 3 * N))
    lr[:32] = 3 * N * o.delta * rng.random(32)
    lm[32:64] = 0.05 * th * rng.random((32, 3 * N)) * (rng.random((32, 3 * N)) < 0.5)
    lm[64:, :2 * N] = 0.05 * th * rng.random((32, 2 * N))
    w, cost, info = solver.solve_lompc_batch(lm, lr, gam, return_info=True)
    assert np.all(info["status"] == 0)
    for b in range(0, B, 4):
        wo, co, _ = orc.solve_active_set(N, o, lm[b], lr[b], gam[b])
        assert np.max(np.abs(w[b] - wo)) <= W_RTOL * o.w_max, (b, info["iters"][b])
        assert abs(cost[b] - co) <= C_RTOL * max(1, abs(co))
        viol, dist = orc.kkt_certificate(N, o, w[b], lm[b], lr[b], gam[b])
        assert dist <= 1e-8


@pytest.mark.parametrize("ev", ["small", "large"])
def test_scalar_api_and_errors(ev):
    """solve_lompc keeps the reference's signature and error behaviour (lompc.py:84-90,137-156)."""
    from chargingstation.lompc import LoMPC
    o, c = _consts(ev)
    N = 12
    solver = LoMPC(N, c)
    rng = np.random.default_rng(5)
    lm = o.theta * rng.random(3 * N)
    w, cost = solver.solve_lompc(lm, 0.3, 0.4)
    assert isinstance(w, np.ndarray) and w.shape == (N,) and isinstance(cost, float)
    wo, co, _ = orc.solve_active_set(N, o, lm, 0.3, 0.4)
    assert np.max(np.abs(w - wo)) <= W_RTOL * o.w_max
    assert abs(cost - orc.lompc_cost(N, o, w, lm, 0.3, 0.4)) <= 1e-12 * max(1, abs(cost))
    with pytest.raises(AssertionError):
        solver.solve_lompc(lm, 0.0, o.y_max + 1e-3)  # lompc.py:87
    with pytest.raises(ValueError):
        solver.solve_lompc(-lm, 0.0, 0.1)  # nonneg cv.Parameter, lompc.py:78
    # batch entry point reports the same conditions through status / return code
    gam = np.array([0.1, o.y_max + 0.01])
    with pytest.raises(AssertionError):
        solver.solve_lompc_batch(lm, 0.0, gam)
    # phi / Dphi / get_price0 mirror lompc.py:164-187
    assert np.allclose(solver.phi(w), orc.phi(N, o, w))
    assert np.allclose(solver.Dphi(w), orc.Dphi(N, o, w))
    assert abs(solver.get_price0(w, lm, 0.3) - orc.get_price0(N, o, w, lm, 0.3)) < 1e-12
    assert solver.get_sc_modulus() == 2 * o.delta * o.theta ** 2
    assert np.array_equal(solver.get_input_mat(), np.tril(np.ones((N, N))))


def test_edge_cases():
    """gamma = 0 (nothing to charge), gamma = y_max, zero prices, huge prices, B = 0."""
    from chargingstation.lompc import LoMPC
    for ev in ("small", "large"):
        o, c = _consts(ev)
        N = 24
        solver = LoMPC(N, c)
        z = np.zeros(3 * N)
        w, cost = solver.solve_lompc(z, 0.0, 0.0)
        assert np.all(w == 0.0) and cost == 0.0
        w, cost = solver.solve_lompc(z, 0.0, o.y_max)
        wo, co, _ = orc.solve_active_set(N, o, z, 0.0, o.y_max)
        assert np.max(np.abs(w - wo)) <= W_RTOL * o.w_max and abs(cost - co) <= C_RTOL * abs(co)
        big = np.concatenate([1e4 * np.ones(N), np.zeros(2 * N)])
        w, cost = solver.solve_lompc(big, 0.0, 0.5)
        assert np.all(w == 0.0)
        neg = np.concatenate([np.zeros(N), 1e4 * np.ones(N), np.zeros(N)])
        w, cost = solver.solve_lompc(neg, 0.0, 0.5)
        assert np.all(w == o.w_max)
        w0, c0 = solver.solve_lompc_batch(np.zeros((0, 3 * N)), np.zeros(0), np.zeros(0))
        assert w0.shape == (0, N) and c0.shape == (0,)


def test_broadcast_prices_match_per_row():
    """lmbd_stride = 0 (one price vector for the whole group, price_solver.py:203-204)."""
    from chargingstation.lompc import LoMPC
    o, c = _consts("large")
    N = 24
    solver = LoMPC(N, c)
    rng = np.random.default_rng(9)
    lm = 0.05 * o.theta * rng.random(3 * N)
    gam = o.y_max - (0.3 + 0.2 * rng.random(333))
    w1, c1 = solver.solve_lompc_batch(lm, 0.0, gam)
    w2, c2 = solver.solve_lompc_batch(np.tile(lm, (333, 1)), np.zeros(333), gam)
    assert np.array_equal(w1, w2) and np.array_equal(c1, c2)


def test_full_size_properties():
    """BASELINE config sizes (65,536 QPs, N = 24): size-independent properties -
    feasibility, KKT residual reported by the kernel, cost identity, determinism."""
    from chargingstation.lompc import LoMPC
    for ev in ("small", "large"):
        o, c = _consts(ev)
        N, B = 24, 65536
        rng = np.random.default_rng(3)
        lm = o.theta * rng.random((B, 3 * N))
        lr = 3 * N * o.delta * rng.random(B)
        gam = o.y_max * rng.random(B)
        solver = LoMPC(N, c)
        w, cost, info = solver.solve_lompc_batch(lm, lr, gam, return_info=True)
        assert np.all(info["status"] == 0)
        assert w.min() >= 0.0 and w.max() <= o.w_max
        assert info["kkt_res"].max() <= 1e-10
        idx = rng.integers(0, B, 64)
        for b in idx:
            assert abs(cost[b] - orc.lompc_cost(N, o, w[b], lm[b], lr[b], gam[b])) <= 1e-11 * max(1, abs(cost[b]))
            viol, dist = orc.kkt_certificate(N, o, w[b], lm[b], lr[b], gam[b])
            assert dist <= 1e-8
        w2, cost2 = solver.solve_lompc_batch(lm, lr, gam)
        assert np.array_equal(w, w2) and np.array_equal(cost, cost2)


def test_torch_device_entry_point():
    import torch
    from chargingstation.lompc import LoMPC
    o, c = _consts("small")
    N, B = 24, 1024
    rng = np.random.default_rng(2)
    lm = o.theta * rng.random((B, 3 * N))
    lr = 3 * N * o.delta * rng.random(B)
    gam = o.y_max * rng.random(B)
    solver = LoMPC(N, c)
    w_h, c_h = solver.solve_lompc_batch(lm, lr, gam)
    dev = torch.device("cuda:0")
    w_d, c_d, info = solver.solve_lompc_batch(torch.from_numpy(lm).to(dev), torch.from_numpy(lr).to(dev),
                                              torch.from_numpy(gam).to(dev), return_info=True)
    torch.cuda.synchronize()
    assert np.array_equal(w_d.cpu().numpy(), w_h) and np.array_equal(c_d.cpu().numpy(), c_h)
    assert int(info["status"].max()) == 0


def test_hard_cases():
    """Near-degenerate inputs on which an earlier build stalled (tests/golden/gen_hard_cases.py)."""
    from chargingstation.lompc import LoMPC
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "lompc_hard_cases.npz"))
    keys = sorted(set("_".join(k.split("_")[:3]) for k in z.files))
    assert len(keys) >= 8
    for key in keys:
        ev, Ns, _ = key.split("_")
        N = int(Ns[1:])
        o, c = _consts(ev)
        solver = LoMPC(N, c)
        w, cost, info = solver.solve_lompc_batch(z[key + "_lmbd"], z[key + "_lmbd_r"], z[key + "_gamma"],
                                                 return_info=True)
        assert np.all(info["status"] == 0), (key, info["status"], info["kkt_res"])
        assert np.max(np.abs(w - z[key + "_w"])) <= W_RTOL * o.w_max, key
        assert np.max(np.abs(cost - z[key + "_cost"]) / np.maximum(1, np.abs(z[key + "_cost"]))) <= C_RTOL, key


@pytest.mark.parametrize("ev", ["small", "large"])
def test_optimistic_phase_hands_over_to_the_safeguarded_loop(ev):
    """K1 runs its first iterations without evaluating the objective when every stage cost of every QP of the
    warp is strictly convex.  (a) nearly degenerate but convex stages (lmbd_r = 1e-9, sparse prices): the
    optimistic phase is chosen, converges slowly and hands over; (b) the same rows with exact zeros interleaved
    (mixed warps: the warp votes for the safeguarded loop).  Both must reach the oracle's optimum."""
    from chargingstation.lompc import LoMPC
    o, c = _consts(ev)
    N, B = 24, 256
    rng = np.random.default_rng(77)
    lm = 0.05 * o.theta * rng.random((B, 3 * N)) * (rng.random((B, 3 * N)) < 0.5)
    gam = o.y_max - (0.3 + 0.2 * rng.random(B))
    solver = LoMPC(N, c)
    for lr in (np.full(B, 1e-9), np.where(np.arange(B) % 3 == 0, 0.0, 1e-9)):
        w, cost, info = solver.solve_lompc_batch(lm, lr, gam, return_info=True)
        assert np.all(info["status"] == 0), info["iters"].max()
        assert info["kkt_res"].max() <= 1e-10
        for b in range(0, B, 16):
            wo, co, _ = orc.solve_active_set(N, o, lm[b], lr[b], gam[b])
            assert np.max(np.abs(w[b] - wo)) <= W_RTOL * o.w_max, (b, info["iters"][b])
            assert abs(cost[b] - co) <= C_RTOL * max(1, abs(co))


def test_async_host_api_matches_blocking_call():
    """solve_lompc_batch(wait=False) on two independent handles + wait() == the blocking calls."""
    from chargingstation.lompc import LoMPC
    N, B = 24, 300
    rng = np.random.default_rng(21)
    solvers, ins, ref = {}, {}, {}
    for ev in ("small", "large"):
        c, lc = _consts(ev)
        solvers[ev] = LoMPC(N, lc)
        ins[ev] = (c.theta * rng.random((B, 3 * N)), 3 * N * c.delta * rng.random(B), c.y_max * rng.random(B))
        ref[ev] = solvers[ev].solve_lompc_batch(*ins[ev])
    outs = {ev: (np.empty((B, N)), np.empty(B)) for ev in solvers}
    for ev in solvers:
        solvers[ev].solve_lompc_batch(*ins[ev], out=outs[ev], wait=False)
    for ev in solvers:
        solvers[ev].wait()
        assert np.array_equal(outs[ev][0], ref[ev][0]) and np.array_equal(outs[ev][1], ref[ev][1])
    bad = ins["small"][2].copy()
    bad[7] = 0.95  # gamma > y_max (lompc.py:87) surfaces at wait()
    solvers["small"].solve_lompc_batch(ins["small"][0], ins["small"][1], bad, wait=False)
    with pytest.raises(AssertionError):
        solvers["small"].wait()


@pytest.mark.parametrize("ev", ["small", "large"])
def test_robustness_bound_of_the_reference_script(ev):
    """test/test_lompc.py:60-86 plots, for 100 SoC spreads, the A_bar-distance between the mean response of
    10 EVs and the response at the mid-point gamma against the bound sqrt(N) * Gamma_bar (and the first-step
    bound scaled by min(1, 1/sqrt(kappa))).  Here the same quantities are ASSERTED (one batch per spread)."""
    from chargingstation.lompc import LoMPC
    o, c = _consts(ev)
    N, nEVs = 12, 10
    solver = LoMPC(N, c)
    rng = np.random.default_rng(11)
    A = solver.get_input_mat()
    lmbd = o.theta * rng.random(3 * N)
    kappa = 3 * N * rng.random() + 1e-5
    lmbd_r = o.delta * kappa
    A_bar = A.T @ A + kappa * np.eye(N)
    gamma_max_arr = o.y_max * np.arange(1, 0, -0.01)
    gam = gamma_max_arr[:, None] * rng.random((len(gamma_max_arr), nEVs))
    gam_ref = (gam.max(axis=1) + gam.min(axis=1)) / 2
    w, _ = solver.solve_lompc_batch(lmbd, lmbd_r, np.concatenate([gam.ravel(), gam_ref]))
    w_avg = w[: gam.size].reshape(len(gamma_max_arr), nEVs, N).mean(axis=1)
    d = w_avg - w[gam.size:]
    w_err = np.sqrt(np.einsum("ij,jk,ik->i", d, A_bar, d))
    w0_err = np.abs(d[:, 0])
    bound = np.sqrt(N) * gamma_max_arr / 2
    assert np.all(w_err <= bound + 1e-12)
    assert np.all(w0_err <= bound * min(1.0, 1.0 / np.sqrt(kappa)) + 1e-12)


@pytest.mark.parametrize("N", [12, 24])
@pytest.mark.parametrize("ev", ["small", "large"])
def test_bulk_copy_rows_kernel_is_bit_identical(ev, N):
    """Kernel variant 9 moves every price row and every result row with the bulk-copy engine (cp.async.bulk through
    shared memory) instead of per-thread loads and stores; the arithmetic is the register kernel's, so the results
    must be the SAME BITS (w, iteration counts, statuses, residuals; the cost to rounding) - on dense and sparse prices (optimistic and safeguarded phases), on the hard cases, on a
    batch that ends in the middle of a CTA, with statuses of invalid rows - and the golden vectors must hold."""
    from chargingstation.lompc import LoMPC
    o, c = _consts(ev)
    rng = np.random.default_rng(900 + N)
    thread = LoMPC(N, c)
    thread.set_kernel_variant(4)
    bulk = LoMPC(N, c)
    bulk.set_kernel_variant(9)
    for B, sparse in ((1, False), (63, False), (64, True), (1000, False), (4099, True), (70001, False)):
        lm = o.theta * rng.random((B, 3 * N))
        if sparse:
            lm *= rng.random((B, 3 * N)) < 0.5
        lr = 3 * N * o.delta * rng.random(B) * (rng.random(B) < 0.7)
        gam = o.y_max * rng.random(B)
        w_t, c_t, i_t = thread.solve_lompc_batch(lm, lr, gam, return_info=True)
        w_b, c_b, i_b = bulk.solve_lompc_batch(lm, lr, gam, return_info=True)
        assert np.array_equal(w_b, w_t), (B, sparse, np.abs(w_b - w_t).max(), np.flatnonzero(np.any(w_b != w_t, axis=1))[:8])
        # (the cost is re-summed in each kernel's epilogue: same formula, the compiler's choice of fused operations)
        assert np.max(np.abs(c_b - c_t) / np.maximum(1.0, np.abs(c_t))) <= 1e-12, (B, sparse)
        assert np.array_equal(i_b["iters"], i_t["iters"]) and np.array_equal(i_b["status"], i_t["status"])
        assert np.array_equal(i_b["kkt_res"], i_t["kkt_res"])
        for b in range(0, B, max(1, B // 6)):
            wo, co, _ = orc.solve_active_set(N, o, lm[b], lr[b], gam[b])
            assert np.max(np.abs(w_b[b] - wo)) <= W_RTOL * o.w_max and abs(c_b[b] - co) <= C_RTOL * max(1, abs(co))
    # invalid rows: same per-QP status, same exception (lompc.py:78-90)
    lm = o.theta * rng.random((130, 3 * N))
    lm[77, 5] = -1.0
    with pytest.raises(ValueError):
        bulk.solve_lompc_batch(lm, 0.0, np.full(130, 0.2))
    with pytest.raises(AssertionError):
        bulk.solve_lompc_batch(np.abs(lm), 0.0, np.where(np.arange(130) == 129, o.y_max + 0.01, 0.2))
    # broadcast prices are not rows of the batch: served by the per-thread loads of the default shape
    w_b, c_b = bulk.solve_lompc_batch(np.abs(lm[0]), 0.0, np.linspace(0.0, o.y_max, 130))
    w_t, c_t = thread.solve_lompc_batch(np.abs(lm[0]), 0.0, np.linspace(0.0, o.y_max, 130))
    assert np.array_equal(w_b, w_t) and np.array_equal(c_b, c_t)
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "lompc_hard_cases.npz"))
    for key in sorted(set("_".join(k.split("_")[:3]) for k in z.files)):
        if key.split("_")[0] != ev or int(key.split("_")[1][1:]) != N:
            continue
        w, cost, info = bulk.solve_lompc_batch(z[key + "_lmbd"], z[key + "_lmbd_r"], z[key + "_gamma"], return_info=True)
        assert np.all(info["status"] == 0), key
        assert np.max(np.abs(w - z[key + "_w"])) <= W_RTOL * o.w_max, key
