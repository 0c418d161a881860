"""LinearForm + ScalarSourceIntegrator + DirichletBC on the GPU ("next" rows f1/f2 of SURVEY.md section 8).

Mirrors fem/linear_form.py:36-86, fem/scalar_source_integrator.py:13-57 and
fem/dirichlet_bc.py:12-235 (single-space variant): same constructor signatures and
`assembly()/apply()/apply_matrix()/apply_vector()` entry points.
"""
from __future__ import annotations

import ctypes as C
from collections.abc import Sequence

import torch

from .. import _lib
from ..basis import device_tables, host_tables, number_of_local_dofs
from ..sparse import CSRTensor
from .bilinear_form import adjacency
from .integrators import Integrator, _Variant, _check_space, process_coef


class ScalarSourceIntegrator(Integrator):
    """(f, v); element vectors (NC, ldof)"""

    def __init__(self, source=None, q=None, *, region=None, batched=False, method=None):
        super().__init__()
        self.source, self.q, self.batched = source, q, batched
        self.set_region(region)
        self.assembly = _Variant(self, {None: self._assembly_default})
        self.assembly.set(method)

    def _assembly_default(self, space, indices=None):
        if indices is not None:
            raise NotImplementedError("chunked (indices=...) assembly is not needed on the GPU path")
        mesh = _check_space(space)
        TD, p, NC = mesh.TD, space.p, mesh.number_of_cells()
        q = p + 3 if self.q is None else self.q
        tabs = device_tables(TD, p, q, mesh.device)
        L = number_of_local_dofs(TD, p)
        NQ = tabs["ws"].shape[0]
        if "phiw" not in tabs:
            tabs["phiw"] = (tabs["ws"][:, None] * tabs["phi"]).contiguous()
        if self.source is None:
            raise ValueError("ScalarSourceIntegrator needs a source")
        kind, val = process_coef(self.source, mesh, host_tables(TD, p, q)["bcs"], self.batched)
        if kind == "matrix":
            raise RuntimeError("source must be scalar-valued")
        code = {"scalar": 0, "cell": 1, "quad": 2}[kind]
        out = torch.empty((NC, L), dtype=torch.float64, device=mesh.device)
        _lib.call("fb2_elem_source", TD, NC, L, NQ, _lib.ptr(mesh.node), _lib.ptr(mesh.cell), _lib.ptr(tabs["phiw"]), code,
                  val if code == 0 else 1.0, None if code == 0 else _lib.ptr(val), _lib.ptr(out), _lib.stream())
        return out


class VectorSourceIntegrator(Integrator):
    """(f, v) for a vector field f on a TensorFunctionSpace; element vectors (NC, ldof * GD) in the space's dof layout
    (fem/vector_source_integrator.py:12-58 with functional.linear_integral:57-59: F[c, I] = |K| sum_q w_q phi_I(q) . f(c, q)).
    `source`: a callable (cartesian or barycentric) or a tensor giving (NC, NQ, GD) values."""

    def __init__(self, source=None, q=None, *, index=slice(None), batched=False):
        super().__init__()
        if index != slice(None):
            raise NotImplementedError("index= sub-selection is not on the accelerated path")
        if batched:
            raise NotImplementedError("batched sources are not on the accelerated path")
        self.source, self.q, self.index, self.batched = source, q, index, batched

    def assembly(self, space, indices=None):
        if indices is not None:
            raise NotImplementedError("chunked (indices=...) assembly is not needed on the GPU path")
        sspace = getattr(space, "scalar_space", None)
        if sspace is None:
            raise RuntimeError("VectorSourceIntegrator needs a TensorFunctionSpace")
        mesh = _check_space(sspace)
        TD, p, NC, GD = mesh.TD, sspace.p, mesh.number_of_cells(), space.dof_numel
        q = p + 3 if self.q is None else self.q
        tabs = device_tables(TD, p, q, mesh.device)
        L = number_of_local_dofs(TD, p)
        NQ = tabs["ws"].shape[0]
        if "phiw" not in tabs:
            tabs["phiw"] = (tabs["ws"][:, None] * tabs["phi"]).contiguous()
        f = self.source
        if f is None:
            raise ValueError("VectorSourceIntegrator needs a source")
        if callable(f):
            bcs = host_tables(TD, p, q)["bcs"]
            if getattr(f, "coordtype", "barycentric") == "barycentric":
                f = f(torch.as_tensor(bcs, dtype=torch.float64, device=mesh.device), index=slice(None))
            else:
                f = f(mesh.bc_to_point(bcs))
        if not isinstance(f, torch.Tensor) or f.ndim != 3 or f.shape[-1] != GD:
            raise NotImplementedError("VectorSourceIntegrator on the accelerated path takes a vector field of shape (NC, NQ, GD)")
        f = f.to(device=mesh.device, dtype=torch.float64).expand(NC, NQ, GD)
        parts = []
        for a in range(GD):
            fa = f[..., a].contiguous()
            out = torch.empty((NC, L), dtype=torch.float64, device=mesh.device)
            _lib.call("fb2_elem_source", TD, NC, L, NQ, _lib.ptr(mesh.node), _lib.ptr(mesh.cell), _lib.ptr(tabs["phiw"]), 2, 1.0,
                      _lib.ptr(fa), _lib.ptr(out), _lib.stream())
            parts.append(out)
        if space.dof_priority:                       # local index a * ldof + i
            return torch.cat(parts, dim=1).contiguous()
        return torch.stack(parts, dim=2).reshape(NC, L * GD).contiguous()      # local index i * GD + a


class LinearForm:
    def __init__(self, space, batch_size: int = 0):
        if isinstance(space, (tuple, list)):
            if len(space) != 1:
                raise ValueError("LinearForm should have only one space.")
            space = space[0]
        if batch_size:
            raise NotImplementedError("batched forms are not on the accelerated path")
        self.space, self.integrators, self._cursor, self._V = space, {}, 0, None

    @property
    def shape(self):
        return (self.space.number_of_global_dofs(),)

    def add_integrator(self, *I, region=None, splitter=None, group=None):
        if len(I) == 0:
            return self
        if len(I) == 1 and isinstance(I[0], Sequence):
            I = tuple(I[0])
        if region is not None:
            raise NotImplementedError("region / sub-domain integration is not on the accelerated path")
        for it in I:
            self.integrators[f"_group_{self._cursor}"] = it
            self._cursor += 1
        return self

    def assembly(self, *, format="dense"):
        if format != "dense":
            raise ValueError(f"Unsupported format {format}.") if format != "coo" else NotImplementedError("format='coo'")
        space = self.space
        sym = adjacency(space)          # a load vector needs the dof -> (cell, i) lists only, not the CSR pattern
        fe = None
        for it in self.integrators.values():
            v = it.assembly(space)
            if not isinstance(v, torch.Tensor) or v.shape != (sym["NC"], sym["L"]):
                raise ValueError(f"Output of source integrators should be (NC, ldof), but got {tuple(getattr(v, 'shape', ()))}.")
            fe = v if fe is None else fe + v          # out of place: an integrator may return a cached block
        F = torch.empty(sym["gdof"], dtype=torch.float64, device=fe.device)
        _lib.call("fb2_gather_vector", sym["gdof"], _lib.ptr(sym["adj_ptr"]), _lib.ptr(sym["adj_pair"]), _lib.ptr(fe.contiguous()),
                  _lib.ptr(F), _lib.stream())
        self._V = F
        return F


class DirichletBC:
    """u = gd on the boundary dofs: F <- F - A uh, F[bd] = uh[bd]; boundary rows/columns of A removed,
    unit diagonal (fem/dirichlet_bc.py:101-229).  The matrix comes back as canonical CSR (sorted
    columns); the reference returns the same matrix with unsorted int64 columns after its COO round trip."""

    def __init__(self, space, gd=None, *, threshold=None, method=None):
        if isinstance(space, tuple):
            raise NotImplementedError("multi-space Dirichlet conditions are not on the accelerated path")
        self.space, self.gd, self.threshold, self.method = space, gd, threshold, method
        self.bctype = "Dirichlet"
        self.gdof = space.number_of_global_dofs()
        if isinstance(threshold, torch.Tensor):
            self.is_boundary_dof = threshold
        else:
            self.is_boundary_dof = space.is_boundary_dof(threshold=threshold, method=method)
        self.boundary_dof_index = self.is_boundary_dof.nonzero().reshape(-1)
        self._mask = self.is_boundary_dof.to(torch.uint8).contiguous()

    def check_matrix(self, matrix):
        if not isinstance(matrix, CSRTensor):
            raise ValueError("The type of matrix must be COOTensor or CSRTensor.")
        if len(matrix.shape) != 2:
            raise ValueError("The matrix must be a 2-D sparse COO matrix.")
        if matrix.shape[0] != matrix.shape[1]:
            raise ValueError("The matrix must be a square matrix.")
        if matrix.shape[0] != self.gdof:
            raise ValueError("The matrix size must match the gdof of the space.")
        return matrix

    def check_vector(self, vector):
        if not isinstance(vector, torch.Tensor):
            raise ValueError("The type of vector must be a tensor.")
        if vector.ndim != 1:
            raise ValueError("The vector must be 1-D (batched vectors are not on the accelerated path).")
        if vector.shape[0] != self.gdof:
            raise ValueError("The vector size must match the gdof of the space.")
        return vector

    def apply(self, A, f, uh=None, gd=None, *, check=True):
        f = self.apply_vector(f, A, uh, gd, check=check)
        A = self.apply_matrix(A, check=check)
        return A, f

    def apply_matrix(self, matrix, *, check=True):
        A = self.check_matrix(matrix) if check else matrix
        lib = _lib.load()
        n = A.shape[0]
        dev = A.device
        ws = _lib.workspace(lib.fb2_bc_workspace_bytes(n), dev)
        crow = torch.empty(n + 1, dtype=torch.int64, device=dev)
        nnz = C.c_int64(0)
        _lib.call("fb2_bc_matrix_count", n, _lib.ptr(A.crow), _lib.ptr(A.col), _lib.ptr(self._mask), _lib.ptr(crow), C.byref(nnz),
                  _lib.ptr(ws), _lib.stream())
        col = torch.empty(nnz.value, dtype=torch.int32, device=dev)
        val = torch.empty(nnz.value, dtype=torch.float64, device=dev)
        _lib.call("fb2_bc_matrix_fill", n, _lib.ptr(A.crow), _lib.ptr(A.col), _lib.ptr(A.values), _lib.ptr(self._mask), _lib.ptr(crow),
                  _lib.ptr(col), _lib.ptr(val), _lib.stream())
        return CSRTensor(crow, col, val, A.sparse_shape)

    def apply_vector(self, vector, matrix, uh=None, gd=None, *, check=True):
        A = self.check_matrix(matrix) if check else matrix
        f = self.check_vector(vector) if check else vector
        gd = self.gd if gd is None else gd
        if gd is None:
            raise RuntimeError("The boundary condition is None.")
        if uh is None:
            uh = torch.zeros_like(f)
        uh, _ = self.space.boundary_interpolate(gd=gd, uh=uh, threshold=self.threshold, method=self.method)
        uh = uh.contiguous()
        blk, tile, mr = A.spmv_plan()
        out = torch.empty_like(f)
        n = A.shape[0]
        _lib.call("fb2_cg_residual", n, A.nnz, _lib.ptr(A.crow), _lib.ptr(A.col), _lib.ptr(A.values), _lib.ptr(uh),
                  _lib.ptr(f.contiguous()), _lib.ptr(out), _lib.ptr(blk), tile, mr, _lib.stream())      # out = f - A uh
        _lib.call("fb2_bc_vector", n, _lib.ptr(self._mask), _lib.ptr(uh), _lib.ptr(out), _lib.stream())
        return out


class DirichletBCOperator:
    """Matrix-free constrained operator (fem/dirichlet_bc_operator.py:13-67): `op @ u` applies the form to u with
    the boundary entries masked out and passes the boundary entries of u through; `apply(F, uh)` builds the right-hand
    side.  `form` is a BilinearForm (assembled or not) or any operator with `@` and `.shape`; usable as `A` in cg()."""

    def __init__(self, form, gd=None, *, threshold=None, isDDof=None, left: bool = True):
        self.form, self.gd = form, gd
        space = form._spaces[0] if hasattr(form, "_spaces") else getattr(form, "space", None)
        self.space = space
        if isDDof is None:
            if space is None:
                raise ValueError("isDDof is required when the operator carries no space")
            isDDof = space.is_boundary_dof(threshold=threshold)
        self.is_boundary_dof = isDDof
        self.boundary_dof_index = isDDof.nonzero().reshape(-1)
        self.shape = tuple(form.shape)

    def init_solution(self):
        uh = torch.zeros(self.shape[1], dtype=torch.float64, device=self.is_boundary_dof.device)
        self.space.boundary_interpolate(self.gd, uh, threshold=self.is_boundary_dof)
        return uh

    def apply(self, F, uh):
        F = F - self.form @ uh
        F[self.is_boundary_dof] = uh[self.is_boundary_dof]
        return F

    def __matmul__(self, u):
        bd = self.is_boundary_dof
        v = u.clone()
        val = v[bd]
        v[bd] = 0.0
        v = self.form @ v
        v[bd] = val
        return v
