"""Reference-element Lagrange bases and the small tables the kernels consume (host, numpy).

Semantics follow the reference backend kernels multi_index_matrix / simplex_shape_function /
simplex_grad_shape_function (fealpy/backend/numpy_backend.py:356-365, 423-477) and the
pre-contraction of the 'fast' diffusion variant (fem/scalar_diffusion_integrator.py:65-79).
These depend only on (TD, p, q): they are computed once on the host in float64 and uploaded.
"""
from functools import lru_cache
from itertools import combinations_with_replacement

import numpy as np

from .quadrature import simplex_quadrature


def number_of_local_dofs(TD: int, p: int) -> int:
    return (p + 1) * (p + 2) // 2 if TD == 2 else (p + 1) * (p + 2) * (p + 3) // 6


@lru_cache(maxsize=None)
def multi_index_matrix(p: int, TD: int) -> np.ndarray:
    """(ldof, TD+1) exponents, rows in descending lexicographic order."""
    rows = [tuple(p - sum(t) if k == 0 else t[k - 1] for k in range(TD + 1))
            for t in _compositions_tail(p, TD)]
    return np.array(rows, dtype=np.int32)


def _compositions_tail(p, TD):
    # all (a_1..a_TD) with sum <= p, ordered so that (a_0, a_1, ...) is descending lexicographic
    seps = sorted(combinations_with_replacement(range(p + 1), TD), reverse=True)
    out = []
    for s in seps:
        ext = (0,) + tuple(s) + (p,)
        full = tuple(ext[k + 1] - ext[k] for k in range(TD + 1))
        out.append(full[1:])
    return out


def _power_table(bc, p):
    # A[..., m, b] = prod_{t<m} (p*lam_b - t) / m!
    A = np.ones(bc.shape[:-1] + (p + 1, bc.shape[-1]))
    fact = 1.0
    for m in range(1, p + 1):
        A[..., m, :] = A[..., m - 1, :] * (p * bc - (m - 1))
    for m in range(1, p + 1):
        fact *= m
        A[..., m, :] = A[..., m, :] * (1.0 / fact)
    return A


def shape_function(bc, p):
    """phi (..., ldof)"""
    bc = np.asarray(bc, dtype=np.float64)
    if p == 1:
        return bc
    TD = bc.shape[-1] - 1
    mi = multi_index_matrix(p, TD)
    A = _power_table(bc, p)
    phi = np.ones(bc.shape[:-1] + (mi.shape[0],))
    for b in range(TD + 1):
        phi = phi * A[..., mi[:, b], b]
    return phi


def grad_shape_function(bc, p):
    """R (..., ldof, TD+1) = d phi_i / d lambda_b"""
    bc = np.asarray(bc, dtype=np.float64)
    TD = bc.shape[-1] - 1
    mi = multi_index_matrix(p, TD)
    A = _power_table(bc, p)
    dA = np.zeros_like(A)
    fact = 1.0
    for m in range(1, p + 1):
        fact *= m
        s = np.zeros_like(bc)
        for skip in range(m):
            term = np.full_like(bc, float(p))
            for t in range(m):
                if t != skip:
                    term = term * (p * bc - t)
            s = s + term
        dA[..., m, :] = s * (1.0 / fact)
    R = np.zeros(bc.shape[:-1] + (mi.shape[0], TD + 1))
    for b in range(TD + 1):
        v = dA[..., mi[:, b], b]
        for b2 in range(TD + 1):
            if b2 != b:
                v = v * A[..., mi[:, b2], b2]
        R[..., b] = v
    return R


@lru_cache(maxsize=None)
def host_tables(TD: int, p: int, q: int):
    """All host tables for (TD, p, q):
       ws (NQ,), bcs (NQ,TD+1), phi (NQ,l), R (NQ,l,TD+1), Mm (l,l), M4 (l,l,NV,NV), Ms (l,l,NG)."""
    qf = simplex_quadrature(TD, q)
    bcs, ws = qf.get_quadrature_points_and_weights()
    phi = shape_function(bcs, p)
    R = grad_shape_function(bcs, p)
    Mm = np.einsum("q,qi,qj->ij", ws, phi, phi)
    M4 = np.einsum("q,qik,qjl->ijkl", ws, R, R)
    NV = TD + 1
    L = phi.shape[1]
    NG = NV * (NV + 1) // 2
    Ms = np.zeros((L, L, NG))
    t = 0
    for k in range(NV):
        for l in range(k, NV):
            Ms[:, :, t] = M4[:, :, k, l] if k == l else M4[:, :, k, l] + M4[:, :, l, k]
            t += 1
    return dict(ws=np.ascontiguousarray(ws), bcs=np.ascontiguousarray(bcs), phi=np.ascontiguousarray(phi),
                R=np.ascontiguousarray(R), Mm=np.ascontiguousarray(Mm), M4=np.ascontiguousarray(M4),
                Ms=np.ascontiguousarray(Ms))


_DEVICE_TABLES = {}


def device_tables(TD: int, p: int, q: int, device):
    """host_tables uploaded once per (TD, p, q, device) as float64 torch tensors"""
    import torch
    key = (TD, p, q, str(device))
    if key not in _DEVICE_TABLES:
        h = host_tables(TD, p, q)
        _DEVICE_TABLES[key] = {k: torch.from_numpy(v).to(device) for k, v in h.items()}
    return _DEVICE_TABLES[key]
