"""Host builds of kernel bodies for CPU-side algorithm tests (test infrastructure only)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "incentive-design-mpc_b200", "csrc")


def _build(name: str, deps: list[str]) -> str:
    src = os.path.join(HERE, name + ".cpp")
    out = os.path.join(HERE, "lib" + name + ".so")
    deps = [src] + deps
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-I" + CSRC,
                               "-o", out, src, "-lm"])
    return out


def _p(x):
    return x.ctypes.data_as(C.c_void_p)


def bimpc_solve(c, Mp_s, Mp_l, beta_s, beta_l, gamma_sm, gamma_lm, x0, demand, omega, tol=1e-9, max_iter=100):
    """Runs csrc/bimpc_solve.cuh on the host for S stations.  `c` is an oracle BiConsts."""
    lib = C.CDLL(_build("bimpc_hostsim", [os.path.join(CSRC, "bimpc_solve.cuh")]))
    fn = lib.bimpc_hostsim_solve
    fn.restype = C.c_int
    fn.argtypes = ([C.c_int, C.c_int] + [C.c_double] * 5 + [C.c_int] + [C.c_double] * 4 + [C.c_int] +
                   [C.c_void_p] * 15 + [C.c_double, C.c_int])
    arrs = [np.ascontiguousarray(np.atleast_2d(np.asarray(v, dtype=np.float64)))
            for v in (Mp_s, Mp_l, beta_s, beta_l, gamma_sm, gamma_lm)]
    S = arrs[0].shape[0]
    x0 = np.ascontiguousarray(np.atleast_1d(np.asarray(x0, dtype=np.float64)))
    demand = np.ascontiguousarray(np.atleast_2d(np.asarray(demand, dtype=np.float64)))
    omega = np.ascontiguousarray(omega, dtype=np.float64)
    ws = np.zeros((S, c.P, c.N))
    wl = np.zeros((S, c.P, c.N))
    ug = np.zeros((S, c.N))
    st = np.zeros(S, dtype=np.int32)
    it = np.zeros(S, dtype=np.int32)
    obj = np.zeros(S)
    rc = fn(c.N, c.P, c.delta, c.c_g, c.u_g_max, c.u_b_max, c.x_max, c.cost_type, c.theta_s, c.theta_l,
            c.w_max_s, c.w_max_l, S, _p(omega), *[_p(v) for v in arrs], _p(x0), _p(demand), _p(ws), _p(wl),
            _p(ug), _p(st), _p(it), _p(obj), tol, max_iter)
    assert rc == 0, rc
    return ws, wl, ug, {"status": st, "iters": it, "objective": obj}
