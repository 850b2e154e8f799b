"""Drop-in mirror of the reference's ``chargingstation/lompc.py`` (class and
method names, argument meaning, return types and error behaviour as in
lompc.py:12-187), with the cvxpy/CLARABEL solve replaced by the batched
sm_100a kernel behind ``include/lompc_b200.h``.

Added (not in the reference): ``LoMPC.solve_lompc_batch`` - the same solve for
a whole batch of (lmbd, lmbd_r, gamma) triples in one kernel launch; the scalar
``solve_lompc`` is that call with B = 1."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from chargingstation import _native
from chargingstation.settings import (MAX_BAT_CHARGE_RATE, MAX_MAX_BAT_SOC,
                                      MIN_MAX_BAT_SOC)


@dataclass
class LoMPCConstants:
    """
    delta:      Relative weight of charging cost.
    theta:      Battery capacity [kWh].
    y_max:      Maximum allowed state of charge (SoC) as a fraction of capacity.
    w_max:      Maximum fraction of charge replenished per time step (normalized charging rate).
    ev_type:    EV type, either "small" or "large".
    """

    delta: float
    theta: float
    y_max: float
    w_max: float
    ev_type: str


def _is_torch_cuda(x) -> bool:
    return type(x).__module__.startswith("torch") and getattr(x, "is_cuda", False)


class LoMPC:
    def __init__(self, N: int, consts: LoMPCConstants, device: int = 0) -> None:
        """
        Inputs:
            N:      LoMPC horizon length.
            consts: LoMPC constants.
            device: CUDA device ordinal (extension; the reference is CPU-only).
        """
        # lompc.py:36-38
        assert (consts.y_max >= MIN_MAX_BAT_SOC) and (consts.y_max <= MAX_MAX_BAT_SOC)
        assert (consts.w_max >= 0) and (consts.w_max <= MAX_BAT_CHARGE_RATE)
        assert (consts.ev_type == "small") or (consts.ev_type == "large")
        self._set_constants(N, consts)
        self._lib = _native.load()
        self._h = C.c_void_p()
        rc = self._lib.lompc_create(
            self.N, float(self.delta), float(self.theta), float(self.y_max), float(self.w_max),
            _native.EV_LARGE if self.ev_type == "large" else _native.EV_SMALL, int(device),
            C.byref(self._h))
        _native.raise_for(rc)
        self.device = int(device)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._lib.lompc_destroy(h)
            self._h = C.c_void_p()

    def _set_constants(self, N: int, consts: LoMPCConstants) -> None:
        # lompc.py:59-71
        self.N = N
        self.delta = consts.delta
        self.theta = consts.theta
        self.y_max = consts.y_max
        self.w_max = consts.w_max
        self.ev_type = consts.ev_type
        # Scaling factor for the quadratic electricity cost.
        self.q_scale = 3 * self.theta / (4 * self.w_max)
        # LoMPC input matrix, y = A w.
        self.A = np.tril(np.ones((self.N, self.N)))
        # Strong convexity modulus.
        self.m = 2 * self.delta * self.theta ** 2

    # ------------------------------------------------------------------ solves
    def set_solver_options(self, max_iter: int = 200, tol: float = 1e-11) -> None:
        _native.raise_for(self._lib.lompc_set_options(self._h, int(max_iter), float(tol)))

    def set_kernel_variant(self, variant: int = 0) -> None:
        """0 = automatic; 1 = any-N shared-memory kernel; 2, 3 = the two register-kernel shapes (tuning)."""
        _native.raise_for(self._lib.lompc_set_kernel_variant(self._h, int(variant)))

    def solve_lompc(self, lmbd: np.ndarray, lmbd_r: float, gamma: float) -> tuple[np.ndarray, float]:
        """
        Inputs:
            lmbd:   Unit price (incentive) vector.
            lmbd_r: Robustness price parameter.
            gamma:  Fraction of battery capacity remaining to be charged.
        Outputs:
            w_opt:      Optimal w vector.
            cost_opt:   Optimal cost.
        """
        lmbd = np.ascontiguousarray(lmbd, dtype=np.float64)
        assert lmbd.shape == (3 * self.N,)
        # lompc.py:87 asserts first; then cvxpy validates the nonneg Parameters (lompc.py:78-82).
        assert gamma <= self.y_max
        if np.any(lmbd < 0) or lmbd_r < 0 or gamma < 0:
            raise ValueError("Parameter value must be nonnegative.")
        w, cost = self.solve_lompc_batch(lmbd[None, :], np.array([lmbd_r], dtype=np.float64),
                                         np.array([gamma], dtype=np.float64))
        return w[0], float(cost[0])

    def solve_lompc_batch(self, lmbd, lmbd_r, gamma, return_info: bool = False, out=None, wait: bool = True):
        """Batched ``solve_lompc``.

        lmbd:   [B, 3N] or [3N] (one price vector broadcast to the batch, the
                ``_get_w_err`` case of price_solver.py:203-204).
        lmbd_r: [B] or scalar.   gamma: [B].
        numpy inputs  -> numpy outputs (host entry point, copies included);
        torch CUDA fp64 tensors -> torch CUDA tensors (device entry point,
        asynchronous on the current torch stream, no status check).
        Returns (w[B, N], cost[B]) and, with ``return_info``, a dict with
        ``status``, ``iters`` and ``kkt_res`` per QP.  ``out=(w, cost)`` reuses
        preallocated (e.g. pinned) output buffers.  ``wait=False`` (numpy path) only
        enqueues the copies and the kernel on this object's stream - call ``wait()``
        before reading the outputs; independent LoMPC objects overlap that way."""
        if _is_torch_cuda(gamma):
            return self._solve_batch_torch(lmbd, lmbd_r, gamma, return_info, out)
        N = self.N
        gamma = np.ascontiguousarray(np.atleast_1d(gamma), dtype=np.float64)
        B = gamma.shape[0]
        lmbd = np.ascontiguousarray(lmbd, dtype=np.float64)
        if lmbd.ndim == 1:
            assert lmbd.shape == (3 * N,)
            lm_stride = 0
        else:
            assert lmbd.shape == (B, 3 * N)
            lm_stride = 3 * N
        lmbd_r = np.ascontiguousarray(np.atleast_1d(lmbd_r), dtype=np.float64)
        if lmbd_r.shape[0] == 1 and B != 1:
            lr_stride = 0
        else:
            assert lmbd_r.shape == (B,)
            lr_stride = 1 if B > 1 else 0
        if out is not None:
            w, cost = out
            assert w.shape == (B, N) and cost.shape == (B,) and w.dtype == np.float64
            assert w.flags.c_contiguous and cost.flags.c_contiguous
        else:
            w = np.empty((B, N), dtype=np.float64)
            cost = np.empty((B,), dtype=np.float64)
        if return_info:
            status = np.empty((B,), dtype=np.int32)
            iters = np.empty((B,), dtype=np.int32)
            kkt = np.empty((B,), dtype=np.float64)
            ptrs = (status.ctypes.data, iters.ctypes.data, kkt.ctypes.data)
        else:
            ptrs = (None, None, None)  # the C side still fetches and checks the status
        fn = self._lib.lompc_solve_batch_host if wait else self._lib.lompc_solve_batch_host_async
        rc = fn(self._h, B, lmbd.ctypes.data, lm_stride, lmbd_r.ctypes.data, lr_stride,
                gamma.ctypes.data, w.ctypes.data, cost.ctypes.data, *ptrs)
        _native.raise_for(rc)
        if not wait:
            self._pending = (lmbd, lmbd_r, gamma, w, cost, ptrs)  # keep the host buffers alive
        if return_info:
            return w, cost, {"status": status, "iters": iters, "kkt_res": kkt}
        return w, cost

    def _solve_batch_torch(self, lmbd, lmbd_r, gamma, return_info, out=None):
        import torch

        N = self.N
        B = gamma.shape[0]
        dev = gamma.device
        assert gamma.dtype == torch.float64 and gamma.is_contiguous()
        assert lmbd.dtype == torch.float64 and lmbd.is_contiguous() and lmbd.device == dev
        lm_stride = 0 if lmbd.dim() == 1 else 3 * N
        assert lmbd.shape == ((3 * N,) if lm_stride == 0 else (B, 3 * N))
        if not torch.is_tensor(lmbd_r):
            lmbd_r = torch.full((1,), float(lmbd_r), dtype=torch.float64, device=dev)
        assert lmbd_r.dtype == torch.float64 and lmbd_r.device == dev
        lr_stride = 0 if lmbd_r.numel() == 1 else 1
        if out is not None:
            w, cost = out
            assert w.shape == (B, N) and w.is_contiguous() and cost.shape == (B,)
        else:
            w = torch.empty((B, N), dtype=torch.float64, device=dev)
            cost = torch.empty((B,), dtype=torch.float64, device=dev)
        if return_info:
            status = torch.empty((B,), dtype=torch.int32, device=dev)
            iters = torch.empty((B,), dtype=torch.int32, device=dev)
            kkt = torch.empty((B,), dtype=torch.float64, device=dev)
            ptrs = (status.data_ptr(), iters.data_ptr(), kkt.data_ptr())
        else:
            ptrs = (None, None, None)
        stream = torch.cuda.current_stream(dev).cuda_stream
        rc = self._lib.lompc_solve_batch_dev(
            self._h, B, lmbd.data_ptr(), lm_stride, lmbd_r.data_ptr(), lr_stride,
            gamma.data_ptr(), w.data_ptr(), cost.data_ptr(), *ptrs, stream)
        _native.raise_for(rc)
        if return_info:
            return w, cost, {"status": status, "iters": iters, "kkt_res": kkt}
        return w, cost

    def wait(self) -> None:
        """Completes a ``solve_lompc_batch(..., wait=False)``; raises like the blocking call."""
        rc = self._lib.lompc_host_wait(self._h)
        self._pending = None
        _native.raise_for(rc)

    # --------------------------------------------------------------- accessors
    def get_sc_modulus(self) -> float:
        return self.m

    def get_input_mat(self) -> np.ndarray:
        return self.A

    def get_price0(self, w: np.ndarray, lmbd: np.ndarray, lmbd_r: float) -> float:
        # lompc.py:164-170
        price0 = (
            self.theta * (w[0] * lmbd[0] + (self.w_max - w[0]) * lmbd[self.N])
            + self.q_scale * w[0] ** 2 * lmbd[2 * self.N]
            + self.theta ** 2 * w[0] ** 2 * lmbd_r
        )
        return price0

    def phi(self, w: np.ndarray) -> np.ndarray:
        assert w.shape == (self.N,)
        # Linear + quadratic prices are given by: lmbd @ phi(w).  (lompc.py:172-177)
        return np.hstack(
            (self.theta * w, self.theta * (self.w_max - w), self.q_scale * (w * w))
        )

    def Dphi(self, w: np.ndarray) -> np.ndarray:
        assert w.shape == (self.N,)
        # lompc.py:179-187
        return np.vstack(
            (self.theta * np.eye(self.N), -self.theta * np.eye(self.N),
             2 * self.q_scale * np.diag(w))
        )
