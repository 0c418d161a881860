"""Binding of the C ABI into a real FEALPy installation (the reference-side stub of INTEGRATION.md).

Works on FEALPy's own objects when its `pytorch` backend is active and the mesh lives on CUDA:

    from fealpy.backend import backend_manager as bm; bm.set_backend('pytorch')
    import fealpy_b200.integration as b200; b200.install()        # registers the 'b200' variants
    A = BilinearForm(space).add_integrator(ScalarDiffusionIntegrator(method='b200')).assembly()   # fealpy.sparse.CSRTensor
    x = fealpy.solver.cg(A, b)                                    # routed to fb2_cg

Plug-in points used (all exist in the reference today):
  * `@Integrator.assembly.register('b200')`  (decorator/variantmethod.py:64-68)   -> K1 kernels
  * `BilinearForm.assembly` override keeping its signature (fem/bilinear_form.py:83-105) -> symbolic + fused numeric
  * `fealpy.solver.cg` wrapper (solver/cg.py:14-17)                                   -> fb2_cg
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


def adapt_space(space):
    """FEALPy LagrangeFESpace / TensorFunctionSpace (torch-cuda tensors) -> fealpy_b200 space that
    shares node / cell storage and REUSES the reference's own cell_to_dof (so the numbering is the
    reference's by construction, for any mesh it can build)."""
    scalar = getattr(space, "scalar_space", space)
    mesh = scalar.mesh
    node, cell = mesh.entity("node"), mesh.entity("cell")
    if not isinstance(node, torch.Tensor):
        raise RuntimeError("fealpy_b200 plugin needs bm.set_backend('pytorch') and a mesh created with device='cuda'")
    TD = cell.shape[1] - 1
    cls = {2: TriangleMesh, 3: TetrahedronMesh}.get(TD)
    if cls is None or node.shape[1] != TD:
        raise NotImplementedError("only triangle (2-D) and tetrahedron (3-D) meshes are on the accelerated path")
    m = cls(node, cell.to(torch.int32))                        # raises RuntimeError for CPU tensors: no fallback
    s = LagrangeFESpace(m, scalar.p)
    c2d = scalar.cell_to_dof()
    m._c2ip[scalar.p] = c2d.to(torch.int32).contiguous()      # the reference's numbering, not ours
    s.number_of_global_dofs = lambda n=scalar.number_of_global_dofs(): n
    if scalar is space:
        return s
    shape = (-1,) + tuple(space.dof_shape) if not space.dof_priority else tuple(space.dof_shape) + (-1,)
    return TensorFunctionSpace(s, shape)


def adapt_integrator(I):
    name = type(I).__name__
    method = None
    try:
        method = I.assembly.get_key(I)
    except Exception:
        pass
    if name == "ScalarDiffusionIntegrator":
        return ScalarDiffusionIntegrator(coef=I.coef, q=I.q, method="fast" if method == "fast" else None)
    if name == "ScalarMassIntegrator":
        return ScalarMassIntegrator(coef=I.coef, q=I.q)
    if name == "LinearElasticityIntegrator":
        return LinearElasticityIntegrator(I.material, q=I.q)
    if name == "GroupIntegrator":
        return [adapt_integrator(i) for i in I.ints]
    raise NotImplementedError(f"{name} is not on the accelerated path")


def assemble_with_b200(bform, *, format="csr"):
    """drop-in body for fealpy.fem.BilinearForm.assembly: same groups, same output container type"""
    from fealpy.sparse import CSRTensor as RefCSR
    space = adapt_space(bform.space)
    bf = B200BilinearForm(space)
    for group in bform.integrators.values():
        a = adapt_integrator(group)
        bf.add_integrator(*a) if isinstance(a, list) else bf.add_integrator(a)
    A = bf.assembly()
    out = RefCSR(A.crow, A.col, A.values, spshape=A.sparse_shape)
    return out if format == "csr" else out.tocoo()


def cg_with_b200(A, b, x0=None, M=None, **kw):
    """drop-in body for fealpy.solver.cg when A is a (fealpy or fealpy_b200) CSR matrix on CUDA"""
    if not isinstance(A, B200CSR):
        A = B200CSR(A.crow, A.col.to(torch.int32), A.values, A.sparse_shape if hasattr(A, "sparse_shape") else A.shape[-2:])
    if M is not None and not isinstance(M, (B200CSR, torch.Tensor)):
        M = B200CSR(M.crow, M.col.to(torch.int32), M.values, M.shape[-2:])
    return b200_cg(A, b, x0, M, **kw)


def install():
    """register the 'b200' variants on the reference classes (idempotent)"""
    import fealpy.fem as fem
    import fealpy.solver as solver
    from fealpy.fem import BilinearForm

    for cls_name, ours in (("ScalarDiffusionIntegrator", ScalarDiffusionIntegrator), ("ScalarMassIntegrator", ScalarMassIntegrator),
                           ("LinearElasticityIntegrator", LinearElasticityIntegrator)):
        ref_cls = getattr(fem, cls_name)

        def make(ours=ours):
            def assembly_b200(self, space, indices=None):
                return adapt_integrator(self).assembly(adapt_space(space))
            return assembly_b200
        ref_cls.assembly.register("b200")(make())

    if not getattr(BilinearForm, "_b200_installed", False):
        ref_assembly = BilinearForm.assembly

        def assembly(self, *, format="csr"):
            try:
                self._M = assemble_with_b200(self, format=format)
                return self._M
            except NotImplementedError:
                return ref_assembly(self, format=format)       # features outside the accelerated path
        BilinearForm.assembly = assembly
        BilinearForm._b200_installed = True
        ref_cg = solver.cg

        def cg(A, b, x0=None, M=None, **kw):
            if isinstance(b, torch.Tensor) and b.is_cuda and hasattr(A, "crow"):
                return cg_with_b200(A, b, x0, M, **kw)
            return ref_cg(A, b, x0, M, **kw)
        solver.cg = cg
