"""ctypes wrapper of oracle/liblompc_oracle.so (the C restatement of the
reference's cvxpy->CLARABEL solve).  ORACLE / CPU-BASELINE INFRASTRUCTURE ONLY:
imported by tests/ and by bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblompc_oracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            subprocess.check_call(["make", "-s", "-C", _HERE])
        lib = C.CDLL(_LIB_PATH)
        lib.oracle_solve_lompc_batch.restype = C.c_int
        lib.oracle_solve_lompc_batch.argtypes = [
            C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int64,
            C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_double, C.c_int,
            C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.oracle_solve_lompc_exact_batch.restype = C.c_int
        lib.oracle_solve_lompc_exact_batch.argtypes = [
            C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int64,
            C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int,
            C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.oracle_lompc_cost.restype = C.c_double
        lib.oracle_max_threads.restype = C.c_int
        _lib = lib
    return _lib


def max_threads() -> int:
    return int(load().oracle_max_threads())


def solve_lompc_batch(N, consts, lmbd, lmbd_r, gamma, tol=1e-8, max_iter=200, nthreads=0):
    """IPM solve of B QPs on `nthreads` host threads (0 = all).  Returns
    (w[B,N], cost[B], iters[B], threads_used)."""
    lib = load()
    gamma = np.ascontiguousarray(np.atleast_1d(gamma), dtype=np.float64)
    B = gamma.shape[0]
    lmbd = np.ascontiguousarray(lmbd, dtype=np.float64)
    lm_stride = 0 if lmbd.ndim == 1 else 3 * N
    lmbd_r = np.ascontiguousarray(np.atleast_1d(lmbd_r), dtype=np.float64)
    lr_stride = 0 if lmbd_r.shape[0] == 1 else 1
    w = np.empty((B, N))
    cost = np.empty(B)
    iters = np.empty(B, dtype=np.int32)
    used = lib.oracle_solve_lompc_batch(
        N, consts.delta, consts.theta, consts.y_max, consts.w_max,
        1 if consts.ev_type == "large" else 0, B, lmbd.ctypes.data, lm_stride,
        lmbd_r.ctypes.data, lr_stride, gamma.ctypes.data, tol, max_iter, nthreads,
        w.ctypes.data, cost.ctypes.data, iters.ctypes.data)
    if used < 0:
        raise ValueError("oracle_solve_lompc_batch: bad arguments")
    return w, cost, iters, used


def solve_lompc_exact_batch(N, consts, lmbd, lmbd_r, gamma, max_iter=10000, nthreads=0):
    """The exact active-set oracle (C twin of ``lompc_oracle.solve_active_set``) for B QPs.
    Returns (w[B,N], cost[B], iters[B])."""
    lib = load()
    gamma = np.ascontiguousarray(np.atleast_1d(gamma), dtype=np.float64)
    B = gamma.shape[0]
    lmbd = np.ascontiguousarray(lmbd, dtype=np.float64)
    lm_stride = 0 if lmbd.ndim == 1 else 3 * N
    lmbd_r = np.ascontiguousarray(np.atleast_1d(lmbd_r), dtype=np.float64)
    lr_stride = 0 if lmbd_r.shape[0] == 1 else 1
    w = np.empty((B, N))
    cost = np.empty(B)
    iters = np.empty(B, dtype=np.int32)
    used = lib.oracle_solve_lompc_exact_batch(
        N, consts.delta, consts.theta, consts.y_max, consts.w_max,
        1 if consts.ev_type == "large" else 0, B, lmbd.ctypes.data, lm_stride,
        lmbd_r.ctypes.data, lr_stride, gamma.ctypes.data, max_iter, nthreads,
        w.ctypes.data, cost.ctypes.data, iters.ctypes.data)
    if used < 0:
        raise ValueError("oracle_solve_lompc_exact_batch: bad arguments")
    if np.any(iters < 0):
        raise RuntimeError("oracle_solve_lompc_exact_batch: Cholesky breakdown")
    return w, cost, iters
