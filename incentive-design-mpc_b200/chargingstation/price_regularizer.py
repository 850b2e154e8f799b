"""Drop-in mirror of the reference's ``chargingstation/price_regularizer.py``.

The reference builds a cvxpy LP ``min c'x, Ax = b, x >= 0`` and solves it with
cvxpy's default solver (price_regularizer.py:20-85).  Every constraint matrix
it is ever given has the pattern ``A = [diag(a_0), ..., diag(a_{r/N-1})]``
(``Dphi(w)'`` in price_solver.py:248-255, ``[I, -I]`` in
test/test_price_regularizer.py:24), for which the LP separates into N one-row
LPs with a closed-form optimum; that is what the CUDA kernel behind
``price_lp_rows_dev`` evaluates.  Other matrices raise ``NotImplementedError``.
"""
from __future__ import annotations

import numpy as np

from chargingstation import _native


class PriceRegularizer:
    """
    Solves the LP:
    min  c.T @ x,
    s.t. A @ x == b,
         x >= 0.
    When c = phi(w), A = D phi(w).T, and b = D phi(w).T @ lmbd,
    where w = w*(lmbd), the LP minimizes total price without
    affecting the incentive controllability property.
    """

    def __init__(self, N: int, r: int, device: int = 0) -> None:
        """
        Inputs:
            N:  Horizon length.
            r:  Unit price (incentive) vector length.
        """
        assert (N >= 0) and (r >= 0)  # price_regularizer.py:26
        self.N = N
        self.r = r
        self.device = int(device)
        self._lib = _native.load()

    def solve_price_regularization(self, A: np.ndarray, b: np.ndarray, c: np.ndarray) -> np.ndarray:
        """
        A x = b should be feasible.
        Inputs:
            A:  Constraint matrix.
            b:  Constraint vector.
            c:  Cost vector.
        Outputs:
            x_opt:  Optimal regularized vector.
        """
        import torch

        N, r = self.N, self.r
        A = np.asarray(A, dtype=np.float64)
        b = np.asarray(b, dtype=np.float64)
        c = np.asarray(c, dtype=np.float64)
        assert A.shape == (N, r) and b.shape == (N,) and c.shape == (r,)
        if N == 0 or r == 0:
            return np.zeros((r,))
        if r % N != 0:
            raise NotImplementedError("constraint matrix must be [diag(a_0), ..., diag(a_{r/N-1})]")
        nb = r // N
        diags = np.stack([np.diagonal(A[:, j * N:(j + 1) * N]) for j in range(nb)])
        rebuilt = np.hstack([np.diag(diags[j]) for j in range(nb)])
        if not np.array_equal(rebuilt, A):
            raise NotImplementedError("constraint matrix must be [diag(a_0), ..., diag(a_{r/N-1})]")
        dev = torch.device("cuda", self.device)
        a_d = torch.from_numpy(np.ascontiguousarray(diags)).to(dev)
        b_d = torch.from_numpy(np.ascontiguousarray(b)).to(dev)
        c_d = torch.from_numpy(np.ascontiguousarray(c.reshape(nb, N))).to(dev)
        x_d = torch.empty((nb, N), dtype=torch.float64, device=dev)
        rc = self._lib.price_lp_rows_dev(self.device, N, nb, a_d.data_ptr(), b_d.data_ptr(), c_d.data_ptr(),
                                         x_d.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        _native.raise_for(rc)
        return x_d.cpu().numpy().reshape(r)
