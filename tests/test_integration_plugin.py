"""The reference-side binding (fealpy_b200.integration) against the REAL FEALPy objects.

The unmodified reference is installed under the git-ignored baseline/_ref (see baseline/ref_loader.py) and travels
to the GPU box with the snapshot; /root/reference is never read here.
  * CPU part: install() registers the plug-ins and leaves numpy / torch-CPU FEALPy users on the reference code.
  * GPU part (-m gpu): FEALPy `pytorch` backend + device='cuda' meshes -> BilinearForm.assembly() returns a
    fealpy.sparse.CSRTensor equal to the golden CSR, fealpy.solver.cg returns the golden solution.
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import ref_loader  # noqa: E402

import cases as C  # noqa: E402
import golden_util as G  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="baseline/_ref (the installed reference) is not present")


@pytest.fixture(scope="module")
def fealpy_installed():
    ref_loader.install()
    from fealpy.backend import backend_manager as bm
    import fealpy_b200.integration as b200
    b200.install()
    b200.install()                                       # idempotent
    yield bm
    bm.set_backend("numpy")


def _poisson_form(mesh, p):
    from fealpy.functionspace import LagrangeFESpace
    from fealpy.fem import BilinearForm, ScalarDiffusionIntegrator, ScalarMassIntegrator
    space = LagrangeFESpace(mesh, p)
    bform = BilinearForm(space)
    bform.add_integrator(ScalarDiffusionIntegrator())
    bform.add_integrator(ScalarMassIntegrator())
    return space, bform


def test_install_registers_plugins(fealpy_installed):
    from fealpy.fem import BilinearForm, ScalarDiffusionIntegrator, ScalarMassIntegrator, LinearElasticityIntegrator
    from fealpy.solver.iterative_solver_manger import IterativeSolverManager
    import fealpy.solver as solver
    for cls in (ScalarDiffusionIntegrator, ScalarMassIntegrator, LinearElasticityIntegrator):
        assert "b200" in cls.assembly.virtual_table
    assert getattr(BilinearForm, "_b200_installed", False)
    assert "b200_cg" in IterativeSolverManager._SOLVER_MAPPING and "b200_jacobi" in IterativeSolverManager._PC_MAPPING
    assert hasattr(solver.cg, "_b200_ref_cg")


def test_numpy_backend_still_assembles_after_install(fealpy_installed):
    """install() must not take the reference path away from CPU users: numpy-backend forms run the reference code and
    reproduce the golden matrix; cg on numpy arrays is the reference's"""
    bm = fealpy_installed
    bm.set_backend("numpy")
    from fealpy.mesh import TetrahedronMesh
    from fealpy.solver import cg
    case = C.by_name("tet_p2_3x2x1_diffmass")
    gold = G.load(case["name"])
    mesh = TetrahedronMesh.from_box(C.box_of(case), *case["dims"])
    space, bform = _poisson_form(mesh, case["p"])
    A = bform.assembly()
    assert isinstance(A.values, np.ndarray)
    assert np.array_equal(A.crow, gold["crow"]) and np.array_equal(A.col, gold["col"])
    assert np.max(np.abs(A.values - gold["values"])) <= 1e-14 * np.max(np.abs(gold["values"]))
    x, info = cg(A, gold["b"], returninfo=True)
    assert info["niter"] == gold["info"]["niter"]


def test_torch_cpu_forms_run_the_reference(fealpy_installed):
    """torch tensors on the CPU are not ours either: the adapter refuses (NotImplementedError), assembly() routes to the
    reference implementation"""
    bm = fealpy_installed
    bm.set_backend("pytorch")
    try:
        import fealpy_b200.integration as b200
        from fealpy.mesh import TriangleMesh
        from fealpy.fem import ScalarDiffusionIntegrator
        mesh = TriangleMesh.from_box([0, 1, 0, 1], 3, 3)
        space, bform = _poisson_form(mesh, 2)
        with pytest.raises(NotImplementedError):
            b200.adapt_space(space)
        with pytest.raises(NotImplementedError):
            ScalarDiffusionIntegrator(method="b200").assembly(space)
        A = bform.assembly()
        assert isinstance(A.values, torch.Tensor) and not A.values.is_cuda and A.nnz > 0
    finally:
        bm.set_backend("numpy")


def test_adapter_rejects_features_outside_the_path(fealpy_installed):
    import fealpy_b200.integration as b200
    from fealpy.fem import ScalarDiffusionIntegrator, ScalarMassIntegrator
    I = ScalarDiffusionIntegrator()
    I.set_region(np.arange(3))
    with pytest.raises(NotImplementedError):
        b200.adapt_integrator(I)
    with pytest.raises(NotImplementedError):
        b200.adapt_integrator(ScalarDiffusionIntegrator(batched=True))
    with pytest.raises(NotImplementedError):
        b200.adapt_integrator(ScalarMassIntegrator(index=np.arange(2)))
    assert b200.adapt_integrator(ScalarMassIntegrator(coef=2.0, q=4)).q == 4


# ---------------------------------------------------------------------------------------------------------------
# GPU: real FEALPy objects on CUDA
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture()
def fealpy_cuda(fealpy_installed):
    bm = fealpy_installed
    bm.set_backend("pytorch")
    yield bm
    bm.set_backend("numpy")


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["tet_p2_3x2x1_diffmass", "tri_p2_5x4_diffmass", "tet_p2_8_diffmass", "tet_p3_2_diffmass"])
def test_fealpy_cuda_assembly_and_cg_match_golden(name, fealpy_cuda):
    """north star: `BilinearForm(space).add_integrator(...).assembly()` returns the same fealpy.sparse CSRTensor, and
    fealpy.solver.cg consumes it -- on FEALPy's own mesh / space / form / integrator objects"""
    import gpu_util as U
    from fealpy.mesh import TriangleMesh, TetrahedronMesh
    from fealpy.sparse import CSRTensor as RefCSR
    import fealpy.solver as solver
    case = C.by_name(name)
    gold = G.load(name)
    cls = TriangleMesh if case["mesh"] == "tri" else TetrahedronMesh
    mesh = cls.from_box(C.box_of(case), *case["dims"], device="cuda")
    space, bform = _poisson_form(mesh, case["p"])
    A = bform.assembly()
    assert isinstance(A, RefCSR), type(A)
    assert A.values.is_cuda and A.crow.dtype == torch.int64 and A.col.dtype == torch.int32 and A.values.dtype == torch.float64
    U.assert_csr_matches(A, gold)
    assert bform._M is A
    b = torch.as_tensor(gold["b"], device="cuda")
    x, info = solver.cg(A, b, returninfo=True)
    assert abs(info["niter"] - gold["info"]["niter"]) <= 1
    assert np.linalg.norm(x.cpu().numpy() - gold["x"]) / np.linalg.norm(gold["x"]) <= 1e-10
    # the matrix is a genuine reference object: the reference's own conversions work on it (its torch-cuda matmul is NOT
    # exercised: `bm.csr_spmm` builds a torch.sparse_csr_tensor from int64 crow + int32 col, which torch reads out of bounds
    # -- "illegal memory access" -- on this stack; fealpy.solver.cg on CUDA therefore only works through this plug-in)
    S = A.to_scipy()
    assert S.shape == tuple(A.shape) and S.nnz == A.nnz
    np.testing.assert_allclose(S @ gold["x"], gold["b"], rtol=0, atol=1e-7 * np.max(np.abs(gold["b"])))


@pytest.mark.gpu
def test_fealpy_cuda_elasticity_and_variant(fealpy_cuda):
    import gpu_util as U
    from fealpy.mesh import TetrahedronMesh
    from fealpy.functionspace import LagrangeFESpace, TensorFunctionSpace
    from fealpy.fem import BilinearForm, LinearElasticityIntegrator, ScalarDiffusionIntegrator
    from fealpy.material.elastic_material import LinearElasticMaterial
    case = C.by_name("tet_p1_3_elasticity_prio")
    gold = G.load(case["name"])
    mesh = TetrahedronMesh.from_box(C.box_of(case), *case["dims"], device="cuda")
    sspace = LagrangeFESpace(mesh, 1)
    space = TensorFunctionSpace(sspace, shape=(3, -1))
    spec = case["groups"][0][0][1]
    mat = LinearElasticMaterial("m", elastic_modulus=spec["E"], poisson_ratio=spec["nu"], hypo=spec["hypo"])
    A = BilinearForm(space).add_integrator(LinearElasticityIntegrator(mat, q=spec["q"])).assembly()
    U.assert_csr_matches(A, gold)
    # the registered 'b200' variant of an integrator returns the element matrices of the K1 kernels
    Ke = ScalarDiffusionIntegrator(method="b200").assembly(sspace)
    Ke_ref = ScalarDiffusionIntegrator().assembly(sspace)
    assert float((Ke - Ke_ref).abs().max()) <= 1e-12 * float(Ke_ref.abs().max())


@pytest.mark.gpu
def test_fealpy_cuda_unsupported_feature_falls_back_to_reference(fealpy_cuda):
    """a region-restricted integrator is outside the accelerated path: the patched assembly() must run the reference
    implementation (on torch-cuda) instead of silently integrating over the whole mesh"""
    from fealpy.mesh import TriangleMesh
    from fealpy.functionspace import LagrangeFESpace
    from fealpy.fem import BilinearForm, ScalarDiffusionIntegrator
    mesh = TriangleMesh.from_box([0, 1, 0, 1], 4, 4, device="cuda")
    space = LagrangeFESpace(mesh, 1)
    region = torch.arange(0, 8, device="cuda")
    A_sub = BilinearForm(space).add_integrator(ScalarDiffusionIntegrator(region=region)).assembly()
    A_all = BilinearForm(space).add_integrator(ScalarDiffusionIntegrator()).assembly()
    assert A_sub.nnz < A_all.nnz


@pytest.mark.gpu
def test_iterative_solver_manager_entries(fealpy_cuda):
    from fealpy.mesh import TetrahedronMesh
    from fealpy.solver.iterative_solver_manger import IterativeSolverManager
    mesh = TetrahedronMesh.from_box([0, 1, 0, 1, 0, 1], 4, 4, 4, device="cuda")
    space, bform = _poisson_form(mesh, 2)
    A = bform.assembly()
    from fealpy_b200.integration.fealpy_plugin import _as_b200_csr
    b = _as_b200_csr(A) @ torch.ones(A.shape[0], dtype=torch.float64, device="cuda")     # (not the reference's torch-cuda matmul, see above)
    ism = IterativeSolverManager()
    ism.set_matrix(A, matrix_type="SP")
    ism.set_solver("b200_cg")
    ism.set_pc("b200_jacobi")
    ism.set_tolerances(rtol=1e-10, atol=1e-14, maxit=2000)
    x = ism.solve(b)
    assert float((x - 1.0).abs().max()) <= 1e-8
