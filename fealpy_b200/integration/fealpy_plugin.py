"""Binding of the C ABI into a real FEALPy installation (the reference-side stub of INTEGRATION.md).

Works on FEALPy's own objects when its `pytorch` backend is active and the mesh lives on CUDA:

    from fealpy.backend import backend_manager as bm; bm.set_backend('pytorch')
    import fealpy_b200.integration as b200; b200.install()        # registers the 'b200' variants
    A = BilinearForm(space).add_integrator(ScalarDiffusionIntegrator()).assembly()   # a fealpy.sparse.CSRTensor
    x = fealpy.solver.cg(A, b)                                    # routed to fb2_cg

Plug-in points used (all exist in the reference today):
  * `@Integrator.assembly.register('b200')`  (decorator/variantmethod.py:64-68)       -> K1 kernels
  * `BilinearForm.assembly` override keeping its signature (fem/bilinear_form.py:83-105) -> symbolic + fused numeric
  * `fealpy.solver.cg` wrapper (solver/cg.py:14-17)                                      -> fb2_cg
  * `IterativeSolverManager.register_solver / register_pc` (solver/iterative_solver_manger.py:26-45) -> 'b200_cg', 'b200_jacobi'

Routing is by where the data lives: forms whose mesh holds CUDA torch tensors go to the CUDA library, everything else
(numpy backend, torch on the CPU) runs the reference's own code unchanged -- that is dispatch on the container, not a
CPU fallback of this package (fealpy_b200 itself has none).  Reference features outside the accelerated path (region=,
batched coefficients, index= sub-selection, other integrators / variants, two-space forms) raise NotImplementedError
inside the adapter, and the patched `assembly()` then runs the reference implementation.
FEALPy itself is NOT imported at module import time: this package has no dependency on it.
"""
from __future__ import annotations

import torch

from ..fem import BilinearForm as B200BilinearForm
from ..fem import LinearElasticityIntegrator, ScalarDiffusionIntegrator, ScalarMassIntegrator
from ..functionspace import LagrangeFESpace, TensorFunctionSpace
from ..mesh import TetrahedronMesh, TriangleMesh
from ..solver import cg as b200_cg
from ..sparse import CSRTensor as B200CSR

_FULL = slice(None)


def on_cuda(space) -> bool:
    """does the (scalar) space's mesh hold CUDA torch tensors?"""
    scalar = getattr(space, "scalar_space", space)
    mesh = getattr(scalar, "mesh", None)
    try:
        node = mesh.entity("node")
    except Exception:
        return False
    return isinstance(node, torch.Tensor) and node.is_cuda


def adapt_space(space):
    """FEALPy LagrangeFESpace / TensorFunctionSpace (torch-cuda tensors) -> fealpy_b200 space that shares node / cell
    storage and REUSES the reference's own cell_to_dof (so the numbering is the reference's by construction, for any
    mesh it can build).  The adapted space -- and with it the cached symbolic pattern -- is kept on the reference space."""
    if not on_cuda(space):
        raise NotImplementedError("fealpy_b200 plugin: needs bm.set_backend('pytorch') and a mesh created with device='cuda'")
    scalar = getattr(space, "scalar_space", space)
    if type(scalar).__name__ != "LagrangeFESpace" or getattr(scalar, "ctype", "C") != "C":
        raise NotImplementedError(f"{type(scalar).__name__} is not on the accelerated path")
    mesh = scalar.mesh
    node, cell = mesh.entity("node"), mesh.entity("cell")
    TD = cell.shape[1] - 1
    cls = {2: TriangleMesh, 3: TetrahedronMesh}.get(TD)
    if cls is None or node.shape[1] != TD or type(mesh).__name__ not in ("TriangleMesh", "TetrahedronMesh"):
        raise NotImplementedError("only triangle (2-D) and tetrahedron (3-D) meshes are on the accelerated path")
    if not 1 <= scalar.p <= 3:
        raise NotImplementedError("Lagrange degree p = 1..3 is on the accelerated path")
    cached = getattr(scalar, "_b200_adapted", None)
    key = (node.data_ptr(), cell.data_ptr(), tuple(node.shape), tuple(cell.shape), scalar.p)
    if cached is None or cached[0] != key:
        m = cls(node, cell.to(torch.int32))
        s = LagrangeFESpace(m, scalar.p)
        c2d = scalar.cell_to_dof()
        m._c2ip[scalar.p] = c2d.to(torch.int32).contiguous()      # the reference's numbering, not ours
        s.number_of_global_dofs = lambda n=scalar.number_of_global_dofs(): n
        cached = (key, s)
        scalar._b200_adapted = cached
    s = cached[1]
    if scalar is space:
        return s
    tcache = getattr(space, "_b200_adapted", None)
    if tcache is None or tcache[0] is not s:
        shape = (-1,) + tuple(space.dof_shape) if not space.dof_priority else tuple(space.dof_shape) + (-1,)
        tcache = (s, TensorFunctionSpace(s, shape))
        space._b200_adapted = tcache
    return tcache[1]


def adapt_integrator(I):
    name = type(I).__name__
    if name == "GroupIntegrator":
        if getattr(I, "get_region", lambda: None)() is not None:
            raise NotImplementedError("region / sub-domain integration is not on the accelerated path")
        return [adapt_integrator(i) for i in I.ints]
    if getattr(I, "get_region", lambda: None)() is not None:
        raise NotImplementedError("region / sub-domain integration is not on the accelerated path")
    if getattr(I, "batched", False):
        raise NotImplementedError("batched coefficients are not on the accelerated path")
    index = getattr(I, "index", _FULL)
    if not (isinstance(index, slice) and index == _FULL):
        raise NotImplementedError("index= sub-selection is not on the accelerated path")
    method = None
    try:
        method = I.assembly.get_key(I)
    except Exception:
        pass
    if method not in (None, "fast", "b200"):
        raise NotImplementedError(f"variant {method!r} is not on the accelerated path")
    if name == "ScalarDiffusionIntegrator":
        return ScalarDiffusionIntegrator(coef=I.coef, q=I.q, method="fast" if method == "fast" else None)
    if name == "ScalarMassIntegrator":
        return ScalarMassIntegrator(coef=I.coef, q=I.q)
    if name == "LinearElasticityIntegrator":
        return LinearElasticityIntegrator(I.material, q=I.q)
    raise NotImplementedError(f"{name} is not on the accelerated path")


def assemble_with_b200(bform, *, format="csr"):
    """drop-in body for fealpy.fem.BilinearForm.assembly: same groups, same output container type"""
    from fealpy.sparse import CSRTensor as RefCSR
    spaces = getattr(bform, "_spaces", (bform.space,))
    if len(spaces) != 1 or getattr(bform, "batch_size", 0):
        raise NotImplementedError("two-space / batched forms are not on the accelerated path")
    if getattr(bform, "_transposed", False):
        raise NotImplementedError("transposed forms are not on the accelerated path")
    space = adapt_space(spaces[0])
    bf = B200BilinearForm(space)
    for group in bform.integrators.values():
        a = adapt_integrator(group)
        bf.add_integrator(*a) if isinstance(a, list) else bf.add_integrator(a)
    A = bf.assembly()
    out = RefCSR(A.crow, A.col, A.values, spshape=A.sparse_shape)
    return out if format == "csr" else out.tocoo()


def _as_b200_csr(A):
    if isinstance(A, B200CSR):
        return A
    cached = getattr(A, "_b200_view", None)
    if cached is not None and cached[0] == (A.crow.data_ptr(), A.col.data_ptr(), A.values.data_ptr()):
        return cached[1]
    col = A.col if A.col.dtype == torch.int32 else A.col.to(torch.int32)
    crow = A.crow if A.crow.dtype == torch.int64 else A.crow.to(torch.int64)
    shape = A.sparse_shape if hasattr(A, "sparse_shape") else tuple(A.shape[-2:])
    V = B200CSR(crow, col, A.values, shape)
    try:
        A._b200_view = ((A.crow.data_ptr(), A.col.data_ptr(), A.values.data_ptr()), V)      # keeps the SpMV plan with the matrix
    except Exception:
        pass
    return V


def cg_with_b200(A, b, x0=None, M=None, **kw):
    """drop-in body for fealpy.solver.cg when b is a CUDA tensor: CSR matrices (fealpy or fealpy_b200) go to fb2_cg,
    any other operator with `@` to the operator path of fealpy_b200.solver.cg"""
    if hasattr(A, "crow") and hasattr(A, "col") and hasattr(A, "values"):
        A = _as_b200_csr(A)
    if M is not None and not isinstance(M, (B200CSR, torch.Tensor)) and hasattr(M, "crow"):
        M = _as_b200_csr(M)
    return b200_cg(A, b, x0, M, **kw)


class B200CGSolver:
    """`IterativeSolverManager` solver entry 'b200_cg' (same protocol as the reference's CGSolver,
    solver/iterative_solver_manger.py:165-181)"""

    def setup(self, A, pc=None, rtol=1e-8, atol=1e-8, maxit=5000, matrix_type="G"):
        self.A, self.pc, self.rtol, self.atol, self.maxit, self.matrix_type = A, pc, rtol, atol, maxit, matrix_type

    def solve(self, A, b):
        return cg_with_b200(A, b, M=self.pc, rtol=self.rtol, atol=self.atol, maxit=self.maxit)


class B200JacobiPreconditioner:
    """`IterativeSolverManager` preconditioner entry 'b200_jacobi': the diagonal CSRTensor of 1/diag(A)
    (solver/iterative_solver_manger.py:273-280), which fb2_cg fuses into its vector kernels"""

    def apply(self, A):
        d = _as_b200_csr(A).diags()
        return B200CSR(d.crow, d.col, 1.0 / d.values, d.sparse_shape)


def install():
    """register the 'b200' variants on the reference classes (idempotent)"""
    import fealpy.fem as fem
    import fealpy.solver as solver
    from fealpy.fem import BilinearForm

    if getattr(BilinearForm, "_b200_installed", False):
        return
    for cls_name in ("ScalarDiffusionIntegrator", "ScalarMassIntegrator", "LinearElasticityIntegrator"):
        ref_cls = getattr(fem, cls_name)

        def assembly_b200(self, space, indices=None):
            if indices is not None:
                raise NotImplementedError("chunked (indices=...) assembly is not needed on the GPU path")
            return adapt_integrator(self).assembly(adapt_space(space))
        ref_cls.assembly.register("b200")(assembly_b200)

    ref_assembly = BilinearForm.assembly

    def assembly(self, *, format="csr"):
        if on_cuda(self.space):
            try:
                self._M = assemble_with_b200(self, format=format)
                return self._M
            except NotImplementedError:
                pass                                   # a reference feature outside the accelerated path
        return ref_assembly(self, format=format)
    BilinearForm.assembly = assembly
    BilinearForm._b200_ref_assembly = ref_assembly
    BilinearForm._b200_installed = True

    ref_cg = solver.cg

    def cg(A, b, x0=None, M=None, **kw):
        if isinstance(b, torch.Tensor) and b.is_cuda and b.dtype == torch.float64:
            return cg_with_b200(A, b, x0, M, **kw)
        return ref_cg(A, b, x0, M, **kw)
    cg._b200_ref_cg = ref_cg
    solver.cg = cg
    try:
        from fealpy.solver.iterative_solver_manger import IterativeSolverManager
        if "b200_cg" not in IterativeSolverManager._SOLVER_MAPPING:
            IterativeSolverManager.register_solver("b200_cg")(B200CGSolver)
        if "b200_jacobi" not in IterativeSolverManager._PC_MAPPING:
            IterativeSolverManager.register_pc("b200_jacobi")(B200JacobiPreconditioner)
    except ImportError:
        pass
