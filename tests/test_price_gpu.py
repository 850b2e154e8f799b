"""GPU parity tests of the price loop (K2-K5) through the C ABI against
oracle/price_oracle.py.  The loop's break test (price_solver.py:125) is
discontinuous, so parity is judged per iteration (teacher forcing on the
oracle's price iterates) and, for whole loops, on iteration counts plus prices
when the counts agree."""
import numpy as np
import pytest

from oracle import lompc_oracle as orc
from oracle import price_oracle as po

pytestmark = pytest.mark.gpu


def _consts(ev):
    from chargingstation.lompc import LoMPCConstants
    o = orc.small_ev_consts() if ev == "small" else orc.large_ev_consts()
    return o, LoMPCConstants(o.delta, o.theta, o.y_max, o.w_max, o.ev_type)


@pytest.mark.parametrize("ev", ["small", "large"])
@pytest.mark.parametrize("price_type", ["linear", "linear-convex"])
def test_price_step_matches_nnls_oracle(ev, price_type):
    """_price_gradient_descent_step (price_solver.py:216-246): exact NNQP optimum."""
    from chargingstation.price_solver import PriceSolver
    o, c = _consts(ev)
    rng = np.random.default_rng(21)
    for N in (12, 24):
        ps = PriceSolver(N, c, price_type)
        ora = po.PriceOracle(N, o, price_type)
        for trial in range(6):
            lmbd_r = [0.0, 0.0, float(N), 3.0 * N][trial % 4]
            w = o.w_max * rng.random(N) * (rng.random(N) < 0.8)
            w_ref = o.w_max * rng.random(N)
            lam = 0.05 * o.theta * rng.random(ps.r) * (rng.random(ps.r) < 0.6)
            A_bar, A_bar_inv = ora._metric(lmbd_r)
            P, q = ora.price_step_matrices(A_bar_inv, w_ref, w, lam)
            x = po.nnqp_exact(P, q)
            for kw in ({"lmbd_r": lmbd_r}, {}):  # explicit kappa and kappa recovered from A_bar_inv
                lam_next, dec = ps._price_gradient_descent_step(A_bar_inv, w_ref, w, lam, **kw)
                assert np.max(np.abs(lam_next - x)) <= 1e-9 * max(1.0, np.max(np.abs(x)))
                assert po.nnqp_kkt(P, q, lam_next) <= 1e-8 * max(1.0, np.max(np.abs(q)))
                dec_o = (lam @ P @ lam + q @ lam) - (x @ P @ x + q @ x)
                assert abs(dec - dec_o) <= 1e-9 * max(1.0, abs(dec_o))


@pytest.mark.parametrize("ev", ["small", "large"])
@pytest.mark.parametrize("price_type", ["linear", "linear-convex"])
def test_regularizer_closed_form_vs_highs(ev, price_type):
    """_regularize_prices: same LP objective as HiGHS, feasible, non-negative (price_regularizer.py:68-85)."""
    from chargingstation.price_solver import PriceSolver
    o, c = _consts(ev)
    N = 12
    rng = np.random.default_rng(22)
    ps = PriceSolver(N, c, price_type)
    ora = po.PriceOracle(N, o, price_type)
    for trial in range(5):
        w = o.w_max * rng.random(N) * (rng.random(N) < 0.8)
        lam = 0.05 * o.theta * rng.random(ps.r)
        x = ps._regularize_prices(w, lam)
        x_lp = ora.regularize_prices(w, lam)
        phi = ora.phi(w)[:ps.r]
        D = ora.Dphi(w)[:ps.r]
        assert np.all(x >= 0)
        assert np.max(np.abs(D.T @ x - D.T @ lam)) <= 1e-10 * max(1, np.max(np.abs(D.T @ lam)))
        assert abs(phi @ x - phi @ x_lp) <= 1e-9 * max(1, abs(phi @ x_lp))
        assert np.allclose(x, po.regularize_closed_form(N, o, ps.r, w, lam), rtol=1e-13, atol=1e-15)


def test_price_regularizer_class_invariants():
    """test/test_price_regularizer.py:6-28 turned into assertions: feasibility and complementarity."""
    from chargingstation.price_regularizer import PriceRegularizer
    N, r = 12, 24
    A = np.block([np.eye(N), -np.eye(N)])
    cvec = np.ones((r,))
    reg = PriceRegularizer(N, r)
    rng = np.random.default_rng(23)
    for _ in range(50):
        b = 200 * (rng.random((N,)) - 0.5)
        x = reg.solve_price_regularization(A, b, cvec)
        assert np.linalg.norm(A @ x - b) <= 1e-12
        assert x[:N] @ x[N:] == 0.0 and np.all(x >= 0)
        x_lp = po.solve_price_regularization_lp(A, b, cvec)
        assert abs(cvec @ x - cvec @ x_lp) <= 1e-9
    with pytest.raises(NotImplementedError):
        reg.solve_price_regularization(rng.random((N, r)), np.ones(N), cvec)


@pytest.mark.parametrize("ev", ["small", "large"])
def test_get_w_err_and_w0_price0(ev):
    from chargingstation.price_solver import PriceSolver
    o, c = _consts(ev)
    N, nEVs = 12, 37
    rng = np.random.default_rng(24)
    for price_type, lmbd_r in (("linear-convex", 0.0), ("linear", 6.0)):
        ps = PriceSolver(N, c, price_type)
        ora = po.PriceOracle(N, o, price_type)
        y0 = 0.3 + 0.2 * rng.random(nEVs)
        ps.set_charge_levels(y0)
        ora.set_charge_levels(y0)
        assert ps.get_gamma_sc() == ora.gamma_sc and ps.get_gamma_sm() == ora.gamma_sm
        assert ps.get_robustness_bounds(lmbd_r) == ora.get_robustness_bounds(lmbd_r)
        lam = np.zeros(3 * N)
        lam[:ps.r] = 0.05 * o.theta * rng.random(ps.r)
        w_ref = o.w_max * rng.random(N)
        A_bar, _ = ora._metric(lmbd_r)
        e_max, e0, e_avg = ps._get_w_err(lam, lmbd_r, w_ref, A_bar)
        o_max, o0, o_avg, _ = ora.get_w_err(lam, lmbd_r, w_ref, A_bar)
        assert abs(e_max - o_max) <= 1e-10 and abs(e0 - o0) <= 1e-10 and abs(e_avg - o_avg) <= 1e-10
        w0, price0 = ps.get_w0_price0(lam[:ps.r], lmbd_r)
        w0_o, price0_o = ora.get_w0_price0(lam[:ps.r], lmbd_r)
        assert np.max(np.abs(w0 - w0_o)) <= 1e-10 * o.w_max and abs(price0 - price0_o) <= 1e-10 * max(1, abs(price0_o))
    with pytest.raises(AssertionError):
        ps.set_charge_levels(np.array([0.1, 0.95]))  # price_solver.py:71


@pytest.mark.parametrize("ev,price_type,lmbd_r", [("small", "linear-convex", 0.0), ("large", "linear-convex", 0.0),
                                                  ("small", "linear", 0.0), ("large", "linear", 12.0)])
def test_compute_optimal_prices_against_oracle(ev, price_type, lmbd_r):
    """Whole loop (price_solver.py:79-174), inputs as test/test_price_solver.py:23-35."""
    from chargingstation.price_solver import PriceSolver
    o, c = _consts(ev)
    N, nEVs = 12, 10
    rng = np.random.default_rng(25)
    ps = PriceSolver(N, c, price_type)
    ora = po.PriceOracle(N, o, price_type)
    y0 = (1 / 36.0) * o.y_max * rng.random(nEVs)
    w_ref = o.w_max * rng.random(N)
    ps.set_charge_levels(y0)
    ora.set_charge_levels(y0)
    trace = []
    lam_o, st_o = ora.compute_optimal_prices(w_ref, lmbd_r, trace=trace)
    lam, st = ps.compute_optimal_prices(w_ref, lmbd_r)
    assert set(st) == {"iter", "price_before_reg", "price_after_reg", "dual_cost_decrease_actual",
                       "dual_cost_decrease_predicted"}
    assert lam.shape == (3 * N,) and np.all(lam >= 0)
    # teacher forcing: at every oracle price iterate the device errors and the next step agree
    A_bar, A_bar_inv = ora._metric(lmbd_r)
    for t in trace[:: max(1, len(trace) // 6)]:
        e_max, e0, e_avg = ps._get_w_err(t["lmbd"], lmbd_r, w_ref, A_bar)
        assert abs(e_avg - t["w_avg_err"]) <= 1e-9 and abs(e_max - t["w_err_max"]) <= 1e-9
        nxt, _ = ps._price_gradient_descent_step(A_bar_inv, w_ref, t["w_k"], t["lmbd"][:ps.r], lmbd_r=lmbd_r)
        nxt_o, _ = ora.price_gradient_descent_step(A_bar_inv, w_ref, t["w_k"], t["lmbd"][:ps.r])
        assert np.max(np.abs(nxt - nxt_o)) <= 1e-8 * max(1, np.max(np.abs(nxt_o)))
    # whole loop
    assert st["iter"] == st_o["iter"]
    assert np.max(np.abs(lam - lam_o)) <= 1e-6 * max(1, np.max(np.abs(lam_o)))
    assert abs(st["price_before_reg"] - st_o["price_before_reg"]) <= 1e-6 * max(1, abs(st_o["price_before_reg"]))
    assert abs(st["price_after_reg"] - st_o["price_after_reg"]) <= 1e-6 * max(1, abs(st_o["price_after_reg"]))
    n = st["iter"]
    assert len(st["dual_cost_decrease_actual"]) == n == len(st_o["dual_cost_decrease_actual"])
    assert np.allclose(st["dual_cost_decrease_actual"], st_o["dual_cost_decrease_actual"], rtol=1e-5, atol=1e-7)
    assert np.allclose(st["dual_cost_decrease_predicted"], st_o["dual_cost_decrease_predicted"], rtol=1e-5, atol=1e-7)
    # invariants the reference prints: dual-cost decrease >= 0 (price_solver.py:135-138), response at the
    # regularised prices unchanged, warm start stored
    assert np.all(st["dual_cost_decrease_predicted"] >= -1e-9)
    assert np.array_equal(ps.prev_prices, lam[:ps.r])
    # second call warm-starts from prev_prices (price_solver.py:103-104,166)
    lam2, st2 = ps.compute_optimal_prices(w_ref, lmbd_r)
    lam2_o, st2_o = ora.compute_optimal_prices(w_ref, lmbd_r)
    assert st2["iter"] == st2_o["iter"]
    assert np.max(np.abs(lam2 - lam2_o)) <= 1e-6 * max(1, np.max(np.abs(lam2_o)))


def test_batch_equals_single_groups():
    """G groups in one device loop == each group on its own (no cross-group coupling)."""
    from chargingstation.price_solver import PriceSolver
    o, c = _consts("large")
    N = 12
    rng = np.random.default_rng(26)
    ps = PriceSolver(N, c, "linear-convex")
    sizes = [5, 0, 17, 1, 9]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    y0 = 0.3 + 0.05 * rng.random(off[-1])
    G = len(sizes)
    w_ref = o.w_max * rng.random((G, N))
    prev = np.zeros((G, 3 * N))
    prices, stats = ps.compute_optimal_prices_batch(off, y0, w_ref, np.zeros(G), prev)
    for g, n in enumerate(sizes):
        if n == 0:
            assert np.all(prices[g] == 0)
            continue
        single = PriceSolver(N, c, "linear-convex")
        single.set_charge_levels(y0[off[g]:off[g + 1]])
        lam, st = single.compute_optimal_prices(w_ref[g], 0.0)
        assert st["iter"] == stats["iter"][g]
        assert np.array_equal(lam, prices[g])
    w0, p0 = ps.get_w0_price0_batch(off, y0, prices, np.zeros(G))
    assert w0.shape == (off[-1],) and p0.shape == (G,) and p0[1] == 0.0


@pytest.mark.parametrize("ev", ["small", "large"])
def test_sharded_phases_two_emulated_ranks(ev):
    """The price_shard_* phases with EVs split over two 'ranks' (two handles on one GPU; the
    all-reduces are emulated by summing the ranks' buffers) reproduce the unsharded loop."""
    import torch
    from chargingstation.price_solver import PriceSolver
    from chargingstation.sharded import CudaShardBackend, shard_groups
    o, c = _consts(ev)
    N = 12
    rng = np.random.default_rng(27)
    sizes = [9, 0, 14, 1, 6, 11]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    G = len(sizes)
    y0 = 0.3 + 0.05 * rng.random(off[-1])
    w_ref = o.w_max * rng.random((G, N))
    prev = np.zeros((G, 3 * N))
    ref = PriceSolver(N, c, "linear-convex")
    prices_ref, stats_ref = ref.compute_optimal_prices_batch(off, y0, w_ref, np.zeros(G), prev, max_iter=300)
    world = 2
    bes = []
    for rank in range(world):
        ps = PriceSolver(N, c, "linear-convex")
        be = CudaShardBackend(ps)
        loc_off, loc_y0, _ = shard_groups(off, y0, rank, world)
        be.begin(loc_off, loc_y0, w_ref, np.zeros(G), prev, 300, False)
        bes.append(be)
    # all-reduce MIN / MAX / SUM / SUM of the group statistics
    smin = torch.minimum(bes[0].stat_min, bes[1].stat_min)
    smax = torch.maximum(bes[0].stat_max, bes[1].stat_max)
    ssum, scnt = bes[0].stat_sum + bes[1].stat_sum, bes[0].stat_cnt + bes[1].stat_cnt
    for be in bes:
        be.stat_min.copy_(smin); be.stat_max.copy_(smax); be.stat_sum.copy_(ssum); be.stat_cnt.copy_(scnt)
        be.start()
    for it in range(300):
        sums = [be.ev_phase()[0] for be in bes]
        tot = sums[0] + sums[1]
        for be in bes:
            be.w_sum.copy_(tot)
        act = [be.group_phase(it) for be in bes]
        assert act[0] == act[1]
        if act[0] == 0:
            break
    outs = [be.finish(False) for be in bes]
    assert np.array_equal(outs[0][0], outs[1][0])  # replicated group phase: bit-identical on both ranks
    assert np.array_equal(outs[0][1]["iter"], stats_ref["iter"])
    # different summation order across ranks: last-bit differences only
    assert np.max(np.abs(outs[0][0] - prices_ref)) <= 1e-8 * max(1.0, np.max(np.abs(prices_ref)))


@pytest.mark.parametrize("ev", ["small", "large"])
def test_phase_split_loop_variants_agree_bitwise(ev):
    """The phase-split loop has a plain form (price_shard_ev_phase with its column-sum kernel, the synchronous
    price_shard_group_phase with separate gamma_sc-solve and bookkeeping launches, everything on one stream) and the
    pipelined one-process form (column sums inside the group step - price_shard_local_sums -, gamma_sc solve and
    bookkeeping in one launch on a side stream beside the next EV phase).  Same kernels' arithmetic in the same order:
    prices, iteration counts and the decrease histories (price_solver.py:135-137) must agree to the last bit."""
    from chargingstation.price_solver import PriceSolver
    from chargingstation.sharded import CudaShardBackend, compute_optimal_prices_sharded
    o, c = _consts(ev)
    N = 24
    rng = np.random.default_rng(91)
    sizes = [32, 7, 0, 41, 1, 32, 19]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    G = len(sizes)
    y0 = np.concatenate([np.sort(0.3 + 0.03 * g + 0.04 * rng.random(n)) for g, n in enumerate(sizes)])
    w_ref = o.w_max * rng.random((G, N)) * 0.7
    prev, zG, max_iter = np.zeros((G, 3 * N)), np.zeros(G), 80
    # plain form, driven phase by phase
    be = CudaShardBackend(PriceSolver(N, c, "linear-convex"))
    be.begin(off, y0, w_ref, zG, prev, max_iter, True)
    be.start()
    for it in range(max_iter):
        be.ev_phase()
        if be.group_phase(it) == 0:
            break
    prices_a, st_a = be.finish(True)
    # pipelined one-process form through the public function, and through price_solve_dev (loop mode 1)
    prices_b, st_b = compute_optimal_prices_sharded(PriceSolver(N, c, "linear-convex"), off, y0, w_ref, zG, prev,
                                                    history=True, max_iter=max_iter)
    ps = PriceSolver(N, c, "linear-convex")
    ps.set_loop_mode(1)
    prices_c, st_c = ps.compute_optimal_prices_batch(off, y0, w_ref, zG, prev, history=True, max_iter=max_iter)
    assert st_a["iter"].max() > 3  # the loop did iterate
    for prices_x, st_x in ((prices_b, st_b), (prices_c, st_c)):
        assert np.array_equal(st_x["iter"], st_a["iter"])
        assert np.array_equal(prices_x, prices_a)
        nonempty = np.asarray(sizes) > 0  # (an empty group is never solved: its w_k row is not defined)
        assert np.array_equal(st_x["w_k"][nonempty], st_a["w_k"][nonempty])
        assert np.array_equal(st_x["hist_ac"], st_a["hist_ac"]) and np.array_equal(st_x["hist_pred"], st_a["hist_pred"])


_REF_SCENARIOS = (  # the four families of test/test_price_solver.py:38-108: (name, nEVs, N, price_type, lmbd_r, max_charge)
    [("single", 1, 12, pt, 0.0, 1 / 3.0) for pt in ("linear", "linear-convex")] +
    [("multiple", 100, 12, "linear-convex", 0.0, 1 / 36.0)] +
    [("horizon", 10, N, "linear-convex", 0.0, 1 / 36.0) for N in (12, 24)] +
    [("robustness", 10, 12, "linear-convex", float(lr), 1 / 36.0) for lr in (0, 12, 24, 36)])


@pytest.mark.parametrize("ev", ["small", "large"])
@pytest.mark.parametrize("scenario", _REF_SCENARIOS, ids=lambda s: f"{s[0]}-{s[1]}ev-N{s[2]}-{s[3]}-lr{s[4]:g}")
def test_reference_convergence_scenarios(ev, scenario):
    """test/test_price_solver.py runs these cases and prints `w0-error | w0 error bound`
    (price_solver.py:150-164); here the printed invariants are asserted: the loop converges before
    the cap, the mean response tracks w_ref within the tolerance, the first-step error respects its
    bound, and the regularisation leaves the response unchanged while not raising the price."""
    from chargingstation.price_solver import PriceSolver
    _, nEVs, N, price_type, lmbd_r, max_charge = scenario
    o, c = _consts(ev)
    rng = np.random.default_rng(1000 + nEVs + N + int(lmbd_r))
    ps = PriceSolver(N, c, price_type)
    y0 = max_charge * o.y_max * rng.random(nEVs)
    w_ref = o.w_max * rng.random(N)
    ps.set_charge_levels(y0)
    lam, st = ps.compute_optimal_prices(w_ref, lmbd_r)
    assert st["iter"] < 999
    tol, w0_bound = ps.get_robustness_bounds(lmbd_r)
    w_err_max, w0_err, w_avg_err = ps._get_w_err(lam, lmbd_r, w_ref, None)
    assert w_avg_err <= tol + 1e-7        # converged iterate; regularisation keeps w*(lmbd) (price_solver.py:249-250)
    assert w0_err <= w0_bound + 1e-9      # price_solver.py:162-164
    assert st["price_after_reg"] <= st["price_before_reg"] + 1e-9 * max(1.0, abs(st["price_before_reg"]))
    assert np.all(lam >= 0) and (price_type == "linear-convex" or np.all(lam[2 * N:] == 0))


@pytest.mark.parametrize("ev", ["small", "large"])
def test_parametric_loop_degenerate_groups(ev):
    """The parametric loop (one warp per group, EVs between two pivots with the same active set interpolated,
    price_set_loop_mode 2) on the groups where its pivot bookkeeping is most fragile: all SoCs equal to the last
    bit (gamma_sc rounds onto the largest gamma), two distinct SoCs, one and two EVs, a wide group.  Same iteration
    counts and prices as the thread-per-EV loop and as the oracle loop."""
    from chargingstation.price_solver import PriceSolver
    o, c = _consts(ev)
    N = 12
    rng = np.random.default_rng(77)
    base = 0.7627601805992867
    cases = {
        "equal_to_the_last_bit": np.array([base, np.nextafter(base, 1.0), np.nextafter(np.nextafter(base, 1.0), 1.0)] * 4),
        "identical": np.full(9, 0.41),
        "two_values": np.array([0.35] * 5 + [0.39] * 4),
        "one_ev": np.array([0.52]),
        "two_evs": np.array([0.31, 0.33]),
        "wide": 0.3 + 0.3 * rng.random(37),
    }
    for name, y0 in cases.items():
        w_ref = o.w_max * rng.random(N) * 0.6
        ora = po.PriceOracle(N, o, "linear-convex", fast=True)
        ora.set_charge_levels(y0)
        lam_o, st_o = ora.compute_optimal_prices(w_ref, 0.0, max_iter=300)
        out = {}
        for mode in (3, 2):
            ps = PriceSolver(N, c, "linear-convex")
            ps.set_loop_mode(mode)
            off = np.array([0, len(y0)], dtype=np.int32)
            prices, st = ps.compute_optimal_prices_batch(off, y0, w_ref[None], np.zeros(1), np.zeros((1, 3 * N)), max_iter=300)
            out[mode] = (prices[0], int(st["iter"][0]))
            assert ps.last_pivot_overflows() == 0
        assert out[2][1] == out[3][1] == st_o["iter"], (name, out[2][1], out[3][1], st_o["iter"])
        small = np.abs(lam_o) <= 1e3
        assert np.max(np.abs(out[2][0] - lam_o)[small]) <= 1e-6 * max(1.0, np.max(np.abs(lam_o[small]))), name


@pytest.mark.parametrize("N", [48, 96])
@pytest.mark.parametrize("ev", ["small", "large"])
def test_parametric_loop_long_horizons(ev, N):
    """compute_optimal_prices (price_solver.py:79-174) at the long horizons of BASELINE configs[4]: the device-resident
    parametric loop (16 / 32 lanes per QP) against the CPU oracle loop - iteration counts equal, prices within 1e-6 -
    and against the phase-split loop (mode 1, the only path these horizons had before)."""
    from chargingstation import settings
    from chargingstation.lompc import LoMPCConstants
    from chargingstation.price_solver import PriceSolver
    settings.PRINT_LEVEL = 0
    o = orc.small_ev_consts() if ev == "small" else orc.large_ev_consts()
    c = LoMPCConstants(o.delta, o.theta, o.y_max, o.w_max, o.ev_type)
    rng = np.random.default_rng(N + (ev == "large"))
    G = 6
    counts = np.array([1, 2, 9, 33, 47, 70])
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    spans = [0.0, 0.02, 0.05, 0.3, 0.04, 0.45]
    y0 = np.concatenate([0.3 + sp * rng.random(n) for n, sp in zip(counts, spans)])
    w_ref = o.w_max * rng.random((G, N)) * 0.6
    out = {}
    for mode in (2, 1):
        ps = PriceSolver(N, c, "linear-convex")
        ps.set_loop_mode(mode)
        prices, st = ps.compute_optimal_prices_batch(off, y0, w_ref, np.zeros(G), np.zeros((G, 3 * N)), max_iter=60)
        out[mode] = (prices, st["iter"].copy(), st["price_before_reg"].copy(), st["price_after_reg"].copy())
        if mode == 2:
            assert ps.last_nnqp_cap_hits() == 0
            assert int(ps._lib.price_last_qp_solves(ps._h)) < int(np.sum((counts + 1) * (st["iter"] + 1)))  # it did interpolate
    assert np.array_equal(out[2][1], out[1][1]), (out[2][1], out[1][1])
    for g in range(G):
        ora = po.PriceOracle(N, o, "linear-convex", fast=True)
        ora.set_charge_levels(y0[off[g]: off[g + 1]])
        lam_o, st_o = ora.compute_optimal_prices(w_ref[g], 0.0, max_iter=60)
        assert out[2][1][g] == st_o["iter"], (g, out[2][1][g], st_o["iter"])
        small = np.abs(lam_o) <= 1e3
        for mode in (2, 1):
            err = np.abs(out[mode][0][g] - lam_o)
            assert np.max(err[small]) <= 1e-6 * max(1.0, np.max(np.abs(lam_o[small]))), (mode, g)
            assert np.all(err[~small] <= 1e-3 * np.abs(lam_o[~small])), (mode, g)
        assert abs(out[2][3][g] - st_o["price_after_reg"]) <= 1e-6 * max(1.0, abs(st_o["price_after_reg"]))
