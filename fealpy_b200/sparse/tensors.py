"""CSRTensor / COOTensor containers with CUDA torch tensors inside.

Mirrors the surface of fealpy.sparse used downstream of assembly
(sparse/csr_tensor.py:16-104,174-184,411-452; sparse/coo_tensor.py:16-43,137-157,184-213):
`.crow/.col/.values` (+ scipy aliases `.indptr/.indices/.data`), `.shape/.nnz/.itype/.dtype`,
`@`, `.to_scipy()`, `.tocoo()/.tocsr()`, `.diags()`, `.toarray()`, `COOTensor.coalesce()`.
`crow` is int64, `col` keeps the mesh itype (int32), `values` float64 -- as the reference returns.
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib


def _bits(n: int) -> int:
    b = 1
    while (1 << b) < n:
        b += 1
    return b


class CSRTensor:
    def __init__(self, crow: torch.Tensor, col: torch.Tensor, values, spshape=None):
        if not isinstance(crow, torch.Tensor) or not isinstance(col, torch.Tensor):
            raise TypeError("crow and col must be tensors")
        if crow.ndim != 1 or col.ndim != 1:
            raise ValueError("crow and col must be 1-D")
        if values is not None and values.shape[-1] != col.shape[0]:
            raise ValueError(f"values must have the same size as col ({col.shape[0]}) in the last dimension")
        self._crow, self._col, self._values = crow, col, values
        if spshape is None:
            ncol = int(col.max().item()) + 1 if col.numel() else 0
            spshape = (crow.shape[0] - 1, ncol)
        if crow.shape[0] != spshape[0] + 1:
            raise ValueError("crow.size must equal nrow + 1")
        self._spshape = tuple(int(v) for v in spshape)
        self._diag_inv = None

    # --- data access (sparse/csr_tensor.py:80-104) -----------------------------------------
    crow = property(lambda s: s._crow)
    col = property(lambda s: s._col)
    values = property(lambda s: s._values)
    indptr = crow
    indices = col
    data = values
    nnz = property(lambda s: s._col.shape[0])
    shape = property(lambda s: (() if s._values is None else tuple(s._values.shape[:-1])) + s._spshape)
    sparse_shape = property(lambda s: s._spshape)
    itype = property(lambda s: s._col.dtype)
    dtype = property(lambda s: None if s._values is None else s._values.dtype)
    ftype = dtype
    device = property(lambda s: s._col.device)

    def __repr__(self):
        return f"CSRTensor(shape={self.shape}, nnz={self.nnz}, device={self.device})"

    # --- products (sparse/csr_tensor.py:411-452 -> csr_spmm) --------------------------------
    def matmul(self, other, out=None):
        """A @ other; `out` (optional, beyond the reference signature): a contiguous float64 result buffer of the right shape,
        so that solver loops allocate nothing per product"""
        if isinstance(other, torch.Tensor):
            if self._values is None:
                raise ValueError("Cannot multiply a CSRTensor without values")
            if self._values.ndim != 1:
                raise NotImplementedError("batched CSR values are not on the accelerated path")
            n, m = self._spshape
            if other.shape[0] != m:
                raise ValueError(f"shape mismatch: {self._spshape} @ {tuple(other.shape)}")
            if other.dtype != torch.float64 or self._col.dtype != torch.int32:
                raise TypeError("fealpy_b200 SpMV needs float64 values/vectors and int32 column indices")
            x = other.contiguous()
            if out is not None and not (out.dtype == torch.float64 and out.is_contiguous() and out.device == x.device
                                        and tuple(out.shape) == (n,) + tuple(x.shape[1:])):
                raise ValueError("out must be a contiguous float64 tensor of the result's shape on the operand's device")
            if x.ndim == 1:
                y = torch.empty(n, dtype=torch.float64, device=x.device) if out is None else out
                blk_row, tile, max_row = self.spmv_plan()
                _lib.call("fb2_csr_spmv", n, self.nnz, _lib.ptr(self._crow), _lib.ptr(self._col), _lib.ptr(self._values),
                          _lib.ptr(x), _lib.ptr(y), _lib.ptr(blk_row), tile, max_row, _lib.stream())
                return y
            if x.ndim == 2:
                y = torch.empty((n, x.shape[1]), dtype=torch.float64, device=x.device) if out is None else out
                _lib.call("fb2_csr_spmm", n, _lib.ptr(self._crow), _lib.ptr(self._col), _lib.ptr(self._values), _lib.ptr(x),
                          _lib.ptr(y), x.shape[1], _lib.stream())
                return y
            raise ValueError("dense operand must be 1-D or 2-D")
        raise TypeError(f"unsupported operand for @: {type(other).__name__}")

    __matmul__ = matmul

    SPMV_TILE = int(__import__('os').environ.get('FB2_SPMV_TILE', '2560'))

    def spmv_plan(self):
        """(blk_row, tile, max_row): row-aligned nnz tiling of the streaming SpMV kernels.  It depends on the pattern only, so
        it is cached ON the `crow` tensor: every matrix assembled on one space (shared crow / col, see
        BilinearForm.share_pattern) reuses it."""
        if getattr(self, "_plan", None) is None:
            cached = getattr(self._crow, "_fb2_spmv_plan", None)
            if cached is not None and cached[1] == self.SPMV_TILE:
                self._plan = cached
                return self._plan
            lib = _lib.load()
            n = self._spshape[0]
            nblk = lib.fb2_spmv_plan_blocks(self.nnz, self.SPMV_TILE)
            blk_row = torch.empty(nblk + 2, dtype=torch.int32, device=self.device)
            mr = C.c_int32(0)
            _lib.call("fb2_spmv_plan_build", n, _lib.ptr(self._crow), self.SPMV_TILE, _lib.ptr(blk_row), self.nnz, C.byref(mr), _lib.stream())
            self._plan = (blk_row, self.SPMV_TILE, mr.value)
            try:
                self._crow._fb2_spmv_plan = self._plan
            except AttributeError:
                pass
        return self._plan

    # --- conversions ----------------------------------------------------------------------
    def row_indices(self):
        n = self._spshape[0]
        counts = self._crow[1:] - self._crow[:-1]
        return torch.repeat_interleave(torch.arange(n, device=self.device, dtype=self._col.dtype), counts)

    def tocoo(self):
        idx = torch.stack([self.row_indices(), self._col], dim=0)
        return COOTensor(idx, self._values, self._spshape, is_coalesced=True)

    def tocsr(self):
        return self

    def to_scipy(self):
        from scipy.sparse import csr_matrix as sp_csr
        return sp_csr((self._values.cpu().numpy(), self._col.cpu().numpy(), self._crow.cpu().numpy()), shape=self._spshape)

    def toarray(self):
        out = torch.zeros(self._spshape, dtype=self._values.dtype, device=self.device)
        out.index_put_((self.row_indices().long(), self._col.long()), self._values, accumulate=True)
        return out

    def diags(self):
        """the diagonal entries as a CSRTensor of the same shape (sparse/csr_tensor.py:467-475 -> partial(), :224-239);
        the reference builds its Jacobi preconditioner from it: CSRTensor(d.crow, d.col, 1/d.values, A.shape)"""
        n = self._spshape[0]
        rows = self.row_indices()
        mask = rows == self._col
        cnt = torch.bincount(rows[mask].long(), minlength=n)
        crow = torch.zeros(n + 1, dtype=torch.int64, device=self.device)
        crow[1:] = torch.cumsum(cnt, dim=0)
        return CSRTensor(crow, self._col[mask].clone(), None if self._values is None else self._values[..., mask].clone(),
                         self._spshape)

    def diagonal(self):
        """main diagonal as a dense vector (absent entries are 0)"""
        n = min(self._spshape)
        rows = self.row_indices()
        mask = rows == self._col
        d = torch.zeros(n, dtype=self._values.dtype, device=self.device)
        d.index_add_(0, rows[mask].long(), self._values[mask])
        return d

    def values_context(self):
        return dict(dtype=self._values.dtype, device=self._values.device)

    def copy(self):
        return CSRTensor(self._crow.clone(), self._col.clone(), None if self._values is None else self._values.clone(),
                         self._spshape)


class COOTensor:
    def __init__(self, indices: torch.Tensor, values, spshape=None, *, is_coalesced=None):
        if not isinstance(indices, torch.Tensor):
            raise TypeError(f"indices must be a Tensor, but got {type(indices)}")
        if indices.ndim != 2:
            raise ValueError(f"indices must be a 2D tensor, but got {indices.ndim}D")
        if values is not None and values.shape[-1] != indices.shape[1]:
            raise ValueError("values must have the same size as indices in the last dimension")
        self._indices, self._values = indices, values
        self.is_coalesced = is_coalesced
        if spshape is None:
            spshape = tuple((indices.max(dim=1).values + 1).tolist())
        elif len(spshape) != indices.shape[0]:
            raise ValueError("length of sparse shape must match the size of indices in dim-0")
        self._spshape = tuple(int(v) for v in spshape)

    indices = property(lambda s: s._indices)
    values = property(lambda s: s._values)
    row = property(lambda s: s._indices[0])
    col = property(lambda s: s._indices[1])
    nnz = property(lambda s: s._indices.shape[1])
    shape = property(lambda s: (() if s._values is None else tuple(s._values.shape[:-1])) + s._spshape)
    sparse_shape = property(lambda s: s._spshape)
    itype = property(lambda s: s._indices.dtype)
    dtype = property(lambda s: None if s._values is None else s._values.dtype)
    device = property(lambda s: s._indices.device)

    def add(self, other, alpha=1.0):
        """concatenation only (sparse/coo_tensor.py:329-342)"""
        if not isinstance(other, COOTensor):
            raise TypeError("only COOTensor + COOTensor is supported")
        if other._spshape != self._spshape:
            raise ValueError("sparse shape mismatch")
        idx = torch.cat([self._indices, other._indices], dim=1)
        vals = torch.cat([self._values, other._values * alpha], dim=-1)
        return COOTensor(idx, vals, self._spshape)

    def _sorted_unique(self):
        """K2: (crow, col, values, perm, seg_start) of the coalesced matrix."""
        if self._spshape.__len__() != 2:
            raise NotImplementedError("only 2-D sparse tensors")
        if self._values is None or self._values.ndim != 1 or self._values.dtype != torch.float64:
            raise NotImplementedError("coalesce needs 1-D float64 values on the accelerated path")
        lib = _lib.load()
        n = self.nnz
        nrow, ncol = self._spshape
        dev = self.device
        cbits = _bits(max(ncol, 2))
        kbits = cbits + _bits(max(nrow, 2))
        idx = self._indices.contiguous()
        ib = idx.element_size()
        keys = torch.empty(n, dtype=torch.int64, device=dev)
        perm = torch.empty(n, dtype=torch.int32, device=dev)
        _lib.call("fb2_coo_keys_from_coo", _lib.ptr(idx[0]), _lib.ptr(idx[1]), ib, n, cbits, _lib.ptr(keys), _lib.stream())
        ws = _lib.workspace(lib.fb2_coo_workspace_bytes(n), dev)
        nnz = C.c_int64(0)
        _lib.call("fb2_coo_symbolic", _lib.ptr(keys), _lib.ptr(perm), n, kbits, _lib.ptr(ws), C.byref(nnz), _lib.stream())
        nnz = nnz.value
        crow = torch.empty(nrow + 1, dtype=torch.int64, device=dev)
        col = torch.empty(nnz, dtype=idx.dtype, device=dev)
        seg = torch.empty(nnz + 1, dtype=torch.int64, device=dev)
        _lib.call("fb2_coo_fill", _lib.ptr(keys), n, cbits, nrow, _lib.ptr(ws), _lib.ptr(crow), _lib.ptr(col), ib, _lib.ptr(seg),
                  _lib.stream())
        vals = torch.empty(nnz, dtype=torch.float64, device=dev)
        _lib.call("fb2_coo_reduce", _lib.ptr(perm), _lib.ptr(seg), nnz, _lib.ptr(self._values.contiguous()), _lib.ptr(vals),
                  _lib.stream())
        return crow, col, vals

    def coalesce(self, accumulate=True):
        """sort by (row, col), sum duplicates left to right (sparse/coo_tensor.py:184-213)"""
        if self.is_coalesced or self.nnz == 0:
            return self
        crow, col, vals = self._sorted_unique()
        return CSRTensor(crow, col, vals, self._spshape).tocoo()

    def tocsr(self):
        """sparse/coo_tensor.py:137-157"""
        if self.nnz == 0:
            crow = torch.zeros(self._spshape[0] + 1, dtype=torch.int64, device=self.device)
            return CSRTensor(crow, self._indices[1], self._values, self._spshape)
        crow, col, vals = self._sorted_unique()
        if not self.is_coalesced and col.shape[0] != self.nnz:
            raise ValueError("tocsr() on a COOTensor with duplicates: call coalesce() first (as the reference requires)")
        return CSRTensor(crow, col, vals, self._spshape)

    def tocoo(self):
        return self

    def to_scipy(self):
        from scipy.sparse import coo_matrix as sp_coo
        return sp_coo((self._values.cpu().numpy(), self._indices.cpu().numpy()), shape=self._spshape)

    def toarray(self):
        out = torch.zeros(self._spshape, dtype=self._values.dtype, device=self.device)
        out.index_put_((self._indices[0].long(), self._indices[1].long()), self._values, accumulate=True)
        return out


def csr_matrix(arg, shape=None):
    data, indices, indptr = arg
    return CSRTensor(indptr, indices, data, shape)


def coo_matrix(arg, shape=None):
    data, (row, col) = arg
    return COOTensor(torch.stack([row, col], dim=0), data, shape)
