"""Conjugate gradients on the GPU with the reference's signature and stopping rules
(fealpy/solver/cg.py:14-123): `cg(A, b, x0=None, M=None, *, batch_first=False, atol=1e-12,
rtol=1e-8, maxit=10000, returninfo=False)`.
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib
from ..sparse import CSRTensor


def _minv_diag(M, n, device):
    """z = M @ r for the preconditioners the accelerated path fuses: a 1-D tensor (the diagonal
    of M), or a diagonal CSRTensor such as the reference's Jacobi `CSRTensor(diags, 1/diag)`
    (solver/iterative_solver_manger.py:273-280)."""
    if M is None:
        return None
    if isinstance(M, torch.Tensor) and M.ndim == 1 and M.shape[0] == n:
        return M.to(device=device, dtype=torch.float64).contiguous()
    if isinstance(M, CSRTensor) and M.nnz == n and M.sparse_shape == (n, n):
        rows = M.row_indices()
        if bool((rows == M.col).all()):
            return M.values.contiguous()
    raise NotImplementedError("only diagonal preconditioners (1-D tensor or diagonal CSRTensor) are on the accelerated path")


def cg(A, b, x0=None, M=None, *, batch_first=False, atol=1e-12, rtol=1e-8, maxit=10000, returninfo=False):
    assert isinstance(b, torch.Tensor), "b must be a Tensor"
    if x0 is not None:
        assert isinstance(x0, torch.Tensor), "x0 must be a Tensor if not None"
    if b.ndim not in (1, 2):
        raise ValueError("b must be a 1D or 2D dense tensor")
    if x0 is not None and x0.shape != b.shape:
        raise ValueError("x0 and b must have the same shape")
    if not isinstance(A, CSRTensor):
        raise TypeError("fealpy_b200.solver.cg needs a fealpy_b200 CSRTensor (assemble with BilinearForm.assembly())")
    if b.ndim == 2:
        raise NotImplementedError("batched right-hand sides are not on the accelerated path yet")
    if b.device.type != "cuda" or b.dtype != torch.float64:
        raise RuntimeError("fealpy_b200.solver.cg needs float64 CUDA tensors; there is no CPU fallback")
    n = A.sparse_shape[0]
    if A.sparse_shape != (n, n) or b.shape[0] != n:
        raise ValueError("shape mismatch between A and b")
    lib = _lib.load()
    x = torch.zeros_like(b) if x0 is None else x0.clone().contiguous()   # inputs are never mutated
    bb = b.contiguous()
    minv = _minv_diag(M, n, b.device)
    ws = _lib.workspace(lib.fb2_cg_workspace_bytes(n, A.nnz), b.device)
    niter, resid = C.c_int(0), C.c_double(0.0)
    _lib.call("fb2_cg", n, A.nnz, _lib.ptr(A.crow), _lib.ptr(A.col), _lib.ptr(A.values), _lib.ptr(bb), _lib.ptr(x), _lib.ptr(minv),
              float(atol), float(rtol), -1 if maxit is None else int(maxit), 0, _lib.ptr(ws), C.byref(niter), C.byref(resid),
              _lib.stream())
    info = {"residual": resid.value, "niter": niter.value}
    return (x, info) if returninfo else x
