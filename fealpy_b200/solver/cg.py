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


def _cg_operator(A, b, x0, M, atol, rtol, maxit, returninfo):
    """The reference recurrence (solver/cg.py:76-123) for operator-valued A (and M): the products are Python-level
    `A @ p` calls, the vector updates and inner products are the library's kernels (fb2_dot, fb2_bcg_update_xr/p with a
    batch of one); the stopping test runs on the device and is polled every few iterations."""
    if b.device.type != "cuda" or b.dtype != torch.float64:
        raise RuntimeError("fealpy_b200.solver.cg needs float64 CUDA tensors; there is no CPU fallback")
    n, dev = b.shape[0], b.device
    bb = b.contiguous()
    pws = _lib.partial_ws(dev)
    sc = torch.zeros(4, dtype=torch.float64, device=dev)         # rTr | pAp | rTr_new | tmp
    rTr, pAp, rTr_new, tmp = (sc[k:k + 1] for k in range(4))

    def dot(u, v, out):
        _lib.call("fb2_dot", n, _lib.ptr(u), _lib.ptr(v), _lib.ptr(out), _lib.ptr(pws), _lib.stream())

    def precond(v):
        if M is None:
            return v
        if isinstance(M, torch.Tensor) and M.ndim == 1:        # the diagonal of M
            return (M * v).contiguous()
        return (M @ v).contiguous()

    info = {"residual": 0.0, "niter": 0}
    dot(bb, bb, tmp)
    b_norm = float(tmp.item()) ** 0.5
    if b_norm < 1e-15:
        x = torch.zeros_like(bb)
        return (x, info) if returninfo else x
    x = torch.zeros_like(bb) if x0 is None else x0.clone().contiguous()
    r = (bb - (A @ x)).contiguous()
    z = precond(r)
    p = z.clone()
    dot(r, z, rTr)
    # stopping test on the device (fb2_bcg_check); after it fires the update kernels do nothing, so x stays the iterate of
    # the stopping iteration while the host polls the state block only every _BCG_CHECK_EVERY iterations
    state = torch.zeros(4, dtype=torch.float64, device=dev)
    mit = -1 if maxit is None else int(maxit)
    done = False
    while not done:
        for _ in range(_BCG_CHECK_EVERY):
            Ap = (A @ p).contiguous()
            dot(p, Ap, pAp)
            _lib.call("fb2_bcg_update_xr", n, 1, _lib.ptr(x), _lib.ptr(r), _lib.ptr(p), _lib.ptr(Ap), _lib.ptr(rTr), _lib.ptr(pAp),
                      _lib.ptr(state), _lib.stream())
            z = precond(r)
            dot(r, z, rTr_new)
            _lib.call("fb2_bcg_check", 1, _lib.ptr(rTr_new), float(atol), float(rtol) * b_norm, mit, _lib.ptr(state), _lib.stream())
            # p = z + beta p ; fb2_bcg_update_p forms z = minv .* r itself, so a general z goes in as "r" with minv = NULL
            _lib.call("fb2_bcg_update_p", n, 1, _lib.ptr(p), _lib.ptr(z), None, _lib.ptr(rTr_new), _lib.ptr(rTr), _lib.ptr(state),
                      _lib.stream())
            rTr, rTr_new = rTr_new, rTr
        st = state.tolist()
        done = st[0] != 0.0
        info["residual"], info["niter"] = st[2], int(st[1])
    return (x, info) if returninfo else x


def cg(A, b, x0=None, M=None, *, batch_first=False, atol=1e-12, rtol=1e-8, maxit=10000, returninfo=False):
    assert isinstance(b, torch.Tensor), "b must be a Tensor"
    if x0 is not None:
        assert isinstance(x0, torch.Tensor), "x0 must be a Tensor if not None"
    if b.ndim not in (1, 2):
        raise ValueError("b must be a 1D or 2D dense tensor")
    if x0 is not None and x0.shape != b.shape:
        raise ValueError("x0 and b must have the same shape")
    if b.ndim == 2:
        if not isinstance(A, CSRTensor):
            raise NotImplementedError("batched right-hand sides need an assembled CSRTensor on the accelerated path")
        return _cg_batched(A, b, x0, M, batch_first, atol, rtol, maxit, returninfo)
    minv, fused = None, isinstance(A, CSRTensor)
    if fused:
        try:
            minv = _minv_diag(M, A.sparse_shape[0], b.device)
        except NotImplementedError:
            if not hasattr(M, "__matmul__"):
                raise
            fused = False
    if not fused:
        # any A / M with __matmul__ (solver/cg.py:10-14): matrix-free forms, DirichletBCOperator, user operators
        if not hasattr(A, "__matmul__"):
            raise TypeError("A must support `A @ x` (a CSRTensor, a BilinearForm or any operator with __matmul__)")
        return _cg_operator(A, b, x0, M, atol, rtol, maxit, returninfo)
    if b.device.type != "cuda" or b.dtype != torch.float64:
        raise RuntimeError("fealpy_b200.solver.cg needs float64 CUDA tensors; there is no CPU fallback")
    n = A.sparse_shape[0]
    if A.sparse_shape != (n, n) or b.shape[0] != n:
        raise ValueError("shape mismatch between A and b")
    lib = _lib.load()
    x = torch.zeros_like(b) if x0 is None else x0.clone().contiguous()   # inputs are never mutated
    bb = b.contiguous()
    ws = _lib.workspace(lib.fb2_cg_workspace_bytes(n, A.nnz), b.device)
    niter, resid = C.c_int(0), C.c_double(0.0)
    blk_row, tile, max_row = A.spmv_plan()        # cached per pattern
    _lib.call("fb2_cg", n, A.nnz, _lib.ptr(A.crow), _lib.ptr(A.col), _lib.ptr(A.values), _lib.ptr(bb), _lib.ptr(x), _lib.ptr(minv),
              float(atol), float(rtol), -1 if maxit is None else int(maxit), 0, _lib.ptr(blk_row), tile, max_row,
              _lib.ptr(ws), C.byref(niter), C.byref(resid), _lib.stream())
    info = {"residual": resid.value, "niter": niter.value}
    return (x, info) if returninfo else x


_BCG_CHECK_EVERY = 8


def _cg_batched(A, b, x0, M, batch_first, atol, rtol, maxit, returninfo):
    """b of shape (dof, batch) (or (batch, dof) with batch_first): per-column alpha/beta, joint stopping
    test on sqrt(sum_k r_k.z_k) -- solver/cg.py:58-121.  Host-driven loop over fb2_csr_spmm + fb2_bcg_*."""
    if b.device.type != "cuda" or b.dtype != torch.float64:
        raise RuntimeError("fealpy_b200.solver.cg needs float64 CUDA tensors; there is no CPU fallback")
    if batch_first:
        b = b.swapaxes(0, 1)
        x0 = None if x0 is None else x0.swapaxes(0, 1)
    n, B = b.shape
    if A.sparse_shape != (n, n):
        raise ValueError("shape mismatch between A and b")
    dev = b.device
    bb = b.contiguous()
    info = {"residual": 0.0, "niter": 0}
    if float(torch.linalg.norm(bb)) < 1e-15:
        x = torch.zeros_like(bb)
    else:
        minv = _minv_diag(M, n, dev)
        pws = _lib.partial_ws(dev)
        x = torch.zeros_like(bb) if x0 is None else x0.contiguous().clone()
        r = bb - (A @ x)
        z = r if minv is None else minv[:, None] * r
        p = z.clone()
        rTr = torch.empty(B, dtype=torch.float64, device=dev)
        rTr_new = torch.empty_like(rTr)
        pAp = torch.empty_like(rTr)
        _lib.call("fb2_bcg_dots", n, B, _lib.ptr(r), _lib.ptr(z), _lib.ptr(rTr), _lib.ptr(pws), _lib.stream())
        b_norm = float(torch.linalg.norm(bb))
        # the stopping test runs on the device (fb2_bcg_check); once it fires the update kernels do nothing, so the host
        # polls the 4-double state block only every CHECK_EVERY iterations instead of synchronising in each one
        state = torch.zeros(4, dtype=torch.float64, device=dev)
        Ap = torch.empty_like(bb)
        z = r if minv is None else torch.empty_like(r)
        mit = -1 if maxit is None else int(maxit)
        done = False
        while not done:
            for _ in range(_BCG_CHECK_EVERY):
                A.matmul(p, out=Ap)
                _lib.call("fb2_bcg_dots", n, B, _lib.ptr(p), _lib.ptr(Ap), _lib.ptr(pAp), _lib.ptr(pws), _lib.stream())
                _lib.call("fb2_bcg_update_xr", n, B, _lib.ptr(x), _lib.ptr(r), _lib.ptr(p), _lib.ptr(Ap), _lib.ptr(rTr), _lib.ptr(pAp),
                          _lib.ptr(state), _lib.stream())
                if minv is not None:
                    torch.mul(minv[:, None], r, out=z)
                _lib.call("fb2_bcg_dots", n, B, _lib.ptr(r), _lib.ptr(z), _lib.ptr(rTr_new), _lib.ptr(pws), _lib.stream())
                _lib.call("fb2_bcg_check", B, _lib.ptr(rTr_new), float(atol), float(rtol) * b_norm, mit, _lib.ptr(state), _lib.stream())
                _lib.call("fb2_bcg_update_p", n, B, _lib.ptr(p), _lib.ptr(r), _lib.ptr(minv), _lib.ptr(rTr_new), _lib.ptr(rTr),
                          _lib.ptr(state), _lib.stream())
                rTr, rTr_new = rTr_new, rTr
            st = state.tolist()                       # one device->host read per CHECK_EVERY iterations
            done = st[0] != 0.0
            info["residual"], info["niter"] = st[2], int(st[1])
    if batch_first:
        x = x.swapaxes(0, 1)
    return (x, info) if returninfo else x
