"""GPU parity tests: the CUDA path (through the C ABI) against golden outputs of the real
reference (tests/golden/*.npz) and against the oracle on seeded inputs.  Run with -m gpu."""
import numpy as np
import pytest
import torch

import cases as C
import golden_util as G

pytestmark = pytest.mark.gpu

IDS = [c["name"] for c in C.CASES]


@pytest.fixture(scope="module")
def U():
    import gpu_util
    return gpu_util


@pytest.mark.parametrize("case", [c for c in C.CASES if c.get("jitter") is None], ids=lambda c: c["name"])
def test_from_box_bit_exact(case, U):
    gold = G.load(case["name"])
    mesh = U.make_mesh(case, gold, from_box=True)
    assert np.array_equal(mesh.node.cpu().numpy(), gold["node"]), "node coordinates must match numpy.linspace bits"
    assert np.array_equal(mesh.cell.cpu().numpy(), gold["cell"])
    assert mesh.cell.dtype == torch.int32 and mesh.node.dtype == torch.float64


@pytest.mark.parametrize("case", C.CASES, ids=IDS)
def test_numbering_bit_exact(case, U):
    gold = G.load(case["name"])
    mesh = U.make_mesh(case, gold)
    sspace, space = U.make_space(case, mesh)
    assert np.array_equal(mesh.edge.cpu().numpy(), gold["edge"])
    assert np.array_equal(sspace.cell_to_dof().cpu().numpy(), gold["cell2dof_scalar"])
    assert np.array_equal(space.cell_to_dof().cpu().numpy(), gold["cell2dof"])
    assert space.number_of_global_dofs() == gold["info"]["gdof"]


@pytest.mark.parametrize("case", [c for c in C.CASES if c.get("elem")], ids=lambda c: c["name"])
def test_element_matrices(case, U):
    gold = G.load(case["name"])
    mesh, space, bform, groups = U.make_form(case, gold)
    k = 0
    for ints in groups:
        for it in ints:
            Ke = it.assembly(space).cpu().numpy()
            ref = gold[f"Ke_{k}"]
            assert Ke.shape == ref.shape
            assert np.max(np.abs(Ke - ref)) <= 1e-12 * np.max(np.abs(ref)), f"K_e[{k}] of {case['name']}"
            k += 1


@pytest.mark.parametrize("path", ["auto", "gather", "coo"])
@pytest.mark.parametrize("case", C.CASES, ids=IDS)
def test_assembly_csr(case, path, U):
    gold = G.load(case["name"])
    mesh, space, bform, _ = U.make_form(case, gold, path=path)
    A = bform.assembly()
    assert A.shape == (gold["info"]["gdof"],) * 2 and A.nnz == gold["info"]["nnz"]
    U.assert_csr_matches(A, gold)
    # second call (warm pattern cache) gives the same bits as the first
    B = bform.assembly()
    assert torch.equal(A.values, B.values) and torch.equal(A.col, B.col) and torch.equal(A.crow, B.crow)


@pytest.mark.parametrize("path", ["auto", "gather"])
@pytest.mark.parametrize("case", [c for c in C.CASES if not c.get("values_only_checksum")], ids=lambda c: c["name"])
def test_matrix_free_product_matches_reference_run(case, path, U):
    """row f4 pinned to the reference: `bform @ u` on the UNASSEMBLED form (fem/bilinear_form.py:126-158) as the real FEALPy
    computed it (tests/golden, tools/gen_golden.py) -- scalar forms through the fused kernel where it applies ('auto') and
    through the K_e product ('gather'), tensor-space (elasticity) forms through the K_e product"""
    gold = G.load(case["name"])
    mesh, space, bform, _ = U.make_form(case, gold, path=path)
    w = bform @ U.t64(gold["matfree_u"])
    assert bform._M is None
    ref = gold["matfree_w"]
    assert np.max(np.abs(w.cpu().numpy() - ref)) <= 1e-12 * np.max(np.abs(ref))
    if path == "gather":
        assert bform.last_matfree == "ke"


@pytest.mark.parametrize("case", [c for c in C.CASES if c.get("cg")], ids=lambda c: c["name"])
def test_spmv_and_cg(case, U):
    from fealpy_b200.solver import cg
    gold = G.load(case["name"])
    mesh, space, bform, _ = U.make_form(case, gold)
    A = bform.assembly()
    n = A.shape[0]
    ones = torch.ones(n, dtype=torch.float64, device="cuda")
    b = A @ ones
    scale = np.max(np.abs(A.values.cpu().numpy()))
    assert np.max(np.abs(b.cpu().numpy() - gold["b"])) <= 1e-12 * scale
    bref = U.t64(gold["b"])
    x, info = cg(A, bref, returninfo=True)
    xr = gold["x"]
    assert np.linalg.norm(x.cpu().numpy() - xr) / np.linalg.norm(xr) <= 1e-10
    assert abs(info["niter"] - gold["info"]["niter"]) <= 1
    assert info["residual"] < max(1e-12, 1e-8 * np.linalg.norm(gold["b"]))
    # inputs untouched, x0 honoured, maxit honoured
    assert torch.equal(bref, U.t64(gold["b"]))
    x1, info1 = cg(A, bref, x0=x, returninfo=True)
    assert info1["niter"] <= 2
    x2, info2 = cg(A, bref, maxit=3, returninfo=True)
    assert info2["niter"] == 3
    xz = cg(A, torch.zeros_like(bref))
    assert float(xz.abs().max()) == 0.0


def test_jacobi_preconditioned_cg_matches_oracle(U):
    from oracle import fem_oracle as O
    from fealpy_b200.solver import cg
    case = C.by_name("tet_p2_4_diffmass_jit")
    gold = G.load(case["name"])
    mesh, space, bform, _ = U.make_form(case, gold)
    A = bform.assembly()
    from fealpy_b200.sparse import CSRTensor
    d = A.diags()                                                 # a CSRTensor, as in the reference (sparse/csr_tensor.py:467-475)
    assert isinstance(d, CSRTensor) and d.nnz == A.shape[0]
    M = CSRTensor(d.crow, d.col, 1.0 / d.values, A.shape)        # solver/iterative_solver_manger.py:273-280
    x, info = cg(A, U.t64(gold["b"]), M=M, returninfo=True)
    x2 = cg(A, U.t64(gold["b"]), M=1.0 / A.diagonal())            # the diagonal as a vector is accepted too
    assert float((x - x2).abs().max()) == 0.0
    crow, col, val = gold["crow"], gold["col"], gold["values"]
    diag = O.csr_matvec(crow, col, val, np.ones(len(crow) - 1)) * 0
    rows = np.repeat(np.arange(len(crow) - 1), np.diff(crow))
    diag = np.zeros(len(crow) - 1); np.add.at(diag, rows[rows == col], val[rows == col])
    xo, oinfo = O.cg(lambda v: O.csr_matvec(crow, col, val, v), gold["b"], Minv=lambda r: r / diag)
    assert abs(info["niter"] - oinfo["niter"]) <= 1
    assert np.linalg.norm(x.cpu().numpy() - xo) / np.linalg.norm(xo) <= 1e-10


def test_sort_and_scan_primitives(U):
    import ctypes as Ct
    from fealpy_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    for n, bits in [(1, 8), (37, 5), (4096, 13), (100003, 40), (1 << 20, 50)]:
        keys_np = rng.integers(0, 1 << bits, size=n, dtype=np.uint64)
        keys = torch.as_tensor(keys_np.view(np.int64), device="cuda")
        vals = torch.empty(n, dtype=torch.int32, device="cuda")
        ws = _lib.workspace(lib.fb2_sort_workspace_bytes(n), "cuda")
        _lib.call("fb2_sort_pairs", _lib.ptr(keys), _lib.ptr(vals), 1, n, bits, _lib.ptr(ws), _lib.stream())
        order = np.argsort(keys_np, kind="stable")
        assert np.array_equal(keys.cpu().numpy().view(np.uint64), keys_np[order])
        assert np.array_equal(vals.cpu().numpy().view(np.uint32), order.astype(np.uint32)), "sort must be stable"
    for n in [1, 255, 4096, 4097, 1000003]:
        a = rng.integers(0, 7, size=n, dtype=np.int32)
        out = torch.empty(n + 1, dtype=torch.int64, device="cuda")
        ws = _lib.workspace(lib.fb2_scan_workspace_bytes(n), "cuda")
        _lib.call("fb2_exclusive_scan_i32", _lib.ptr(U.t64(a)), _lib.ptr(out), n, _lib.ptr(ws), _lib.stream())
        assert np.array_equal(out.cpu().numpy(), np.concatenate([[0], np.cumsum(a, dtype=np.int64)]))


def test_reference_sparse_vectors(U):
    # test/sparse/test_coo_tensor.py:29-48 and :377-392 of the reference
    from fealpy_b200.sparse import COOTensor
    RT = G.ref_testdata()
    coo = COOTensor(U.t64(RT["coo_indices"]), U.t64(RT["coo_values"]), (3, 3), is_coalesced=False)
    c = coo.coalesce()
    assert c.is_coalesced
    assert np.array_equal(c.indices.cpu().numpy(), RT["coo_expected_indices"])
    assert np.array_equal(c.values.cpu().numpy(), RT["coo_expected_values"])
    D = RT["tocsr_dense"]
    rr, cc = np.nonzero(D)
    A = COOTensor(U.t64(np.stack([rr, cc])), U.t64(D[rr, cc]), D.shape, is_coalesced=True).tocsr()
    assert np.array_equal(A.crow.cpu().numpy(), RT["tocsr_crow"])
    assert np.array_equal(A.col.cpu().numpy(), RT["tocsr_col"])
    assert np.array_equal(A.values.cpu().numpy(), RT["tocsr_values"])
    assert np.array_equal(A.toarray().cpu().numpy(), D)


def test_reference_element_vectors(U):
    # test/fem/test_scalar_diffusion_integrator.py:16-24, test_scalar_mass_integrator.py:16-24
    from fealpy_b200.mesh import TriangleMesh
    from fealpy_b200.functionspace import LagrangeFESpace
    from fealpy_b200.fem import ScalarDiffusionIntegrator, ScalarMassIntegrator
    RT = G.ref_testdata()
    mesh = TriangleMesh.from_box([0, 1, 0, 1], 1, 1)
    space = LagrangeFESpace(mesh, 2)
    np.testing.assert_array_almost_equal(ScalarDiffusionIntegrator(1, 3).assembly(space).cpu().numpy(), RT["diffusion_p2_q3_box1_Ke"])
    np.testing.assert_array_almost_equal(ScalarMassIntegrator(1, 3).assembly(space).cpu().numpy(), RT["mass_p2_q3_box1_Ke"])
    assert np.array_equal(space.cell_to_dof().cpu().numpy(), RT["lfs_box1_p2_cell_to_dof"])
    assert np.array_equal(space.is_boundary_dof().cpu().numpy(), RT["lfs_box1_p2_is_boundary_dof"])


@pytest.mark.parametrize("k", [0, 1])
@pytest.mark.parametrize("p", [1, 2, 3])
def test_reference_bilinear_form_matmul(k, p, U):
    # test/fem/test_bilinear_form.py:19-44: element-wise product == assembled product
    from fealpy_b200.mesh import TriangleMesh
    from fealpy_b200.functionspace import LagrangeFESpace
    from fealpy_b200.fem import BilinearForm, ScalarDiffusionIntegrator
    RT = G.ref_testdata()
    mesh = TriangleMesh(U.t64(RT[f"bform_mesh{k}_node"]), U.t64(RT[f"bform_mesh{k}_cell"].astype(np.int32)))
    space = LagrangeFESpace(mesh, p)
    gdof = space.number_of_global_dofs()
    x = torch.rand(gdof, dtype=torch.float64, device="cuda")
    I = ScalarDiffusionIntegrator()
    Ke, c2d = I.assembly(space), space.cell_to_dof().long()
    y = torch.zeros(gdof, dtype=torch.float64, device="cuda")
    y.index_add_(0, c2d.reshape(-1), torch.einsum("cij,cj->ci", Ke, x[c2d]).reshape(-1))
    bform = BilinearForm(space).add_integrator(I)
    w = bform @ x                      # matrix-free (row f4): no global matrix yet
    assert bform._M is None
    z = bform.assembly() @ x
    assert bform._M is not None
    assert float((y - z).norm()) < 1e-12 and float((w - z).norm()) < 1e-12
    assert float(((bform @ x) - z).norm()) == 0.0


@pytest.mark.parametrize("mesh_kind,p,n", [("tri", 1, 9), ("tri", 2, 7), ("tri", 3, 6), ("tet", 1, 5), ("tet", 2, 4), ("tet", 3, 3)])
def test_summed_element_block_matches_oracle(mesh_kind, p, n, U):
    """the K_e gather path sums several scalar integrators into ONE block (constant terms folded into the first
    quadrature-loop kernel, later kernels accumulate in place; symmetric quadrature kernel for scalar fields): against the
    oracle's element matrices summed in numpy -- variable + per-cell + constant diffusion, matrix-valued diffusion, variable
    and constant mass with different quadrature orders, in one form"""
    from oracle import fem_oracle as O
    from fealpy_b200.mesh import TriangleMesh, TetrahedronMesh
    from fealpy_b200.functionspace import LagrangeFESpace
    from fealpy_b200.fem import BilinearForm, ScalarDiffusionIntegrator, ScalarMassIntegrator
    rng = np.random.default_rng(4321 + 10 * p + n)
    if mesh_kind == "tri":
        node, cell = O.tri_from_box([0, 1, 0, 2], n, n + 1)
    else:
        node, cell = O.tet_from_box([0, 1, 0, 2, -1, 0], n, n + 1, n)
    node = node + rng.uniform(-0.2, 0.2, node.shape) / (n + 1)
    om = O.Mesh(node, cell)
    NC, GD = cell.shape[0], node.shape[1]
    qd, qm = p + 2, p + 3
    NQd, NQm = O.quadrature(om.TD, qd)[0].shape[0], O.quadrature(om.TD, qm)[0].shape[0]
    k_quad = rng.uniform(0.5, 2.0, (NC, NQd))
    k_cell = rng.uniform(0.5, 2.0, NC)
    k_mat = rng.uniform(-0.3, 0.3, (NC, NQd, GD, GD)) + np.eye(GD)
    c_quad = rng.uniform(0.5, 2.0, (NC, NQm))
    ke_ref = (O.diffusion_element(om, p, q=qd, coef=k_quad) + O.diffusion_element(om, p, q=qd, coef=k_cell)
              + O.diffusion_element(om, p, q=qd + 1, coef=0.75) + O.diffusion_element(om, p, q=qd, coef=k_mat)
              + O.mass_element(om, p, q=qm, coef=c_quad) + O.mass_element(om, p, q=qm, coef=2.5))
    cls = TriangleMesh if mesh_kind == "tri" else TetrahedronMesh
    mesh = cls(U.t64(node), U.t64(cell.astype(np.int32)))
    space = LagrangeFESpace(mesh, p)
    bf = BilinearForm(space, assembly_path="gather")
    bf.add_integrator(ScalarDiffusionIntegrator(coef=U.t64(k_quad), q=qd), ScalarDiffusionIntegrator(coef=U.t64(k_cell), q=qd))
    bf.add_integrator(ScalarDiffusionIntegrator(coef=0.75, q=qd + 1))
    bf.add_integrator(ScalarDiffusionIntegrator(coef=U.t64(k_mat), q=qd))
    bf.add_integrator(ScalarMassIntegrator(coef=U.t64(c_quad), q=qm), ScalarMassIntegrator(coef=2.5, q=qm))
    ke = bf._summed_ke()
    scale = np.abs(ke_ref).max()
    assert np.abs(ke.cpu().numpy() - ke_ref).max() <= 1e-12 * scale
    # the per-integrator kernels (symmetric variant for scalar fields, full variant for matrix coefficients) on their own
    for I, ref in ((ScalarDiffusionIntegrator(coef=U.t64(k_quad), q=qd), O.diffusion_element(om, p, q=qd, coef=k_quad)),
                   (ScalarDiffusionIntegrator(coef=U.t64(k_mat), q=qd), O.diffusion_element(om, p, q=qd, coef=k_mat)),
                   (ScalarMassIntegrator(coef=U.t64(c_quad), q=qm), O.mass_element(om, p, q=qm, coef=c_quad))):
        k1 = I.assembly(space).cpu().numpy()
        assert np.abs(k1 - ref).max() <= 1e-12 * np.abs(ref).max()
    sym = ScalarDiffusionIntegrator(coef=U.t64(k_quad), q=qd).assembly(space)
    assert torch.equal(sym, sym.transpose(1, 2)), "scalar-field blocks are exactly symmetric"
    # and the assembled matrix through the gather path equals the COO pipeline on the same form
    A = bf.assembly()
    bf2 = BilinearForm(space, assembly_path="coo")
    for it in bf.integrators.values():
        bf2.add_integrator(it)
    B = bf2.assembly()
    assert torch.equal(A.crow, B.crow) and torch.equal(A.col, B.col)
    assert float((A.values - B.values).abs().max()) <= 1e-12 * float(B.values.abs().max())


@pytest.mark.parametrize("mesh_kind,p,n", [("tri", 1, 9), ("tri", 2, 7), ("tri", 3, 6), ("tet", 1, 5), ("tet", 2, 4), ("tet", 3, 3)])
def test_fused_matfree_matches_oracle(mesh_kind, p, n, U):
    """row f4 without K_e: `form @ u` before assembly on constant / per-cell coefficient forms runs the fused
    cell kernel + row-owner gather (csrc/assemble.cu matfree_cell_kernel); compared with the oracle's restatement of
    fem/bilinear_form.py:126-158 (einsum + index_add on the oracle's element matrices), with the K_e-based product and
    with the assembled product.  Jittered geometry, per-cell diffusion coefficient, scalar mass coefficient."""
    from oracle import fem_oracle as O
    from fealpy_b200.mesh import TriangleMesh, TetrahedronMesh
    from fealpy_b200.functionspace import LagrangeFESpace
    from fealpy_b200.fem import BilinearForm, ScalarDiffusionIntegrator, ScalarMassIntegrator
    rng = np.random.default_rng(1234 + 10 * p + n)
    if mesh_kind == "tri":
        node, cell = O.tri_from_box([0, 1, 0, 2], n, n + 1)
    else:
        node, cell = O.tet_from_box([0, 1, 0, 2, -1, 0], n, n + 1, n)
    node = node + rng.uniform(-0.2, 0.2, node.shape) / (n + 1)
    om = O.Mesh(node, cell)
    NC = cell.shape[0]
    kappa = rng.uniform(0.5, 2.0, NC)
    c2d = om.cell_to_ipoint(p)
    gdof = int(c2d.max()) + 1
    groups = [(O.diffusion_element(om, p, coef=kappa), c2d), (O.mass_element(om, p, coef=2.5), c2d)]
    u = rng.standard_normal(gdof)
    v_ref = O.matfree_apply(groups, gdof, u)

    cls = TriangleMesh if mesh_kind == "tri" else TetrahedronMesh
    mesh = cls(U.t64(node), U.t64(cell.astype(np.int32)))
    space = LagrangeFESpace(mesh, p)
    assert np.array_equal(space.cell_to_dof().cpu().numpy(), c2d)

    def form(path):
        bf = BilinearForm(space, assembly_path=path)
        bf.add_integrator(ScalarDiffusionIntegrator(coef=U.t64(kappa)))
        bf.add_integrator(ScalarMassIntegrator(coef=2.5))
        return bf
    ud = U.t64(u)
    f1 = form("auto")
    v1 = f1 @ ud
    assert f1.last_matfree == "fused" and f1._M is None and not hasattr(space, "_b200_symbolic")
    f2 = form("gather")
    v2 = f2 @ ud
    assert f2.last_matfree == "ke"
    v3 = form("auto").assembly() @ ud
    scale = np.abs(v_ref).max()
    for v in (v1, v2, v3):
        assert np.abs(v.cpu().numpy() - v_ref).max() <= 1e-12 * scale
    assert torch.equal(v1, f1 @ ud), "bit-reproducible (fixed summation order, no atomics)"
    # only a mass term / only a diffusion term
    for ints, grp in (([ScalarMassIntegrator(coef=U.t64(kappa))], [(O.mass_element(om, p, coef=kappa), c2d)]),
                      ([ScalarDiffusionIntegrator(coef=3.0)], [(O.diffusion_element(om, p, coef=3.0), c2d)])):
        bf = BilinearForm(space)
        bf.add_integrator(*ints)
        w = (bf @ ud).cpu().numpy()
        w_ref = O.matfree_apply(grp, gdof, u)
        assert bf.last_matfree == "fused" and np.abs(w - w_ref).max() <= 1e-12 * np.abs(w_ref).max()


def test_fused_matfree_at_size_and_in_cg(U):
    """tet P2 40^3: the fused matrix-free product equals the assembled SpMV to rounding, and cg() on the unassembled form
    (operator-valued A, solver/cg.py:10-14) reaches the assembled solve's solution"""
    from fealpy_b200.mesh import TetrahedronMesh
    from fealpy_b200.functionspace import LagrangeFESpace
    from fealpy_b200.fem import BilinearForm, ScalarDiffusionIntegrator, ScalarMassIntegrator
    from fealpy_b200.solver import cg
    mesh = TetrahedronMesh.from_box([0, 1, 0, 1, 0, 1], 40, 40, 40)
    space = LagrangeFESpace(mesh, 2)
    form = BilinearForm(space).add_integrator(ScalarDiffusionIntegrator()).add_integrator(ScalarMassIntegrator())
    g = torch.Generator(device="cuda").manual_seed(7)
    u = torch.rand(space.number_of_global_dofs(), dtype=torch.float64, device="cuda", generator=g)
    w = form @ u
    assert form.last_matfree == "fused" and form._M is None
    xs = cg(form, w, atol=1e-14, rtol=1e-12, maxit=2000)        # matrix-free solve: A x = A u
    form2 = BilinearForm(space).add_integrator(ScalarDiffusionIntegrator()).add_integrator(ScalarMassIntegrator())
    A = form2.assembly()
    z = A @ u
    assert float((w - z).abs().max()) <= 1e-13 * float(z.abs().max())
    assert float((xs - u).norm() / u.norm()) <= 1e-9


@pytest.mark.parametrize("mesh_kind,p,n", [("tri", 1, 96), ("tri", 3, 40), ("tet", 1, 20), ("tet", 2, 14), ("tet", 3, 6)])
def test_paths_agree_and_properties_at_size(mesh_kind, p, n, U):
    """sizes beyond the golden ladder: the three device paths must give the same pattern
    bit-for-bit and values to 1e-12; diffusion rows sum to 0, the mass matrix sums to |domain|,
    both symmetric (size-independent properties)."""
    from fealpy_b200.mesh import TriangleMesh, TetrahedronMesh
    from fealpy_b200.functionspace import LagrangeFESpace
    from fealpy_b200.fem import BilinearForm, ScalarDiffusionIntegrator, ScalarMassIntegrator
    mesh = TriangleMesh.from_box([0, 1, 0, 1], n, n) if mesh_kind == "tri" else TetrahedronMesh.from_box([0, 1, 0, 1, 0, 1], n, n, n)
    space = LagrangeFESpace(mesh, p)
    mats = {}
    for path in ("auto", "gather", "coo"):
        bf = BilinearForm(space, assembly_path=path)
        bf.add_integrator(ScalarDiffusionIntegrator())
        bf.add_integrator(ScalarMassIntegrator())
        mats[path] = bf.assembly()
    A = mats["auto"]
    scale = float(A.values.abs().max())
    for path in ("gather", "coo"):
        B = mats[path]
        assert torch.equal(A.crow, B.crow) and torch.equal(A.col, B.col), path
        assert float((A.values - B.values).abs().max()) <= 1e-12 * scale, path
    one = torch.ones(A.shape[0], dtype=torch.float64, device="cuda")
    K = BilinearForm(space).add_integrator(ScalarDiffusionIntegrator()).assembly()
    Mm = BilinearForm(space).add_integrator(ScalarMassIntegrator()).assembly()
    assert float((K @ one).abs().max()) <= 1e-11 * float(K.values.abs().max())
    assert abs(float((Mm @ one).sum()) - 1.0) <= 1e-12
    S = A.to_scipy()
    assert abs(S - S.T).max() <= 1e-12 * scale
    assert S.has_sorted_indices, "columns must be ascending within every row"


def test_unsupported_features_raise(U):
    from fealpy_b200.mesh import TriangleMesh
    from fealpy_b200.functionspace import LagrangeFESpace
    from fealpy_b200.fem import BilinearForm, ScalarDiffusionIntegrator
    from fealpy_b200.solver import cg
    mesh = TriangleMesh.from_box([0, 1, 0, 1], 2, 2)
    with pytest.raises(NotImplementedError):
        LagrangeFESpace(mesh, 4)
    space = LagrangeFESpace(mesh, 1)
    with pytest.raises(NotImplementedError):
        BilinearForm(space, batch_size=2)
    with pytest.raises(NotImplementedError):
        ScalarDiffusionIntegrator(method="isopara")
    with pytest.raises(ValueError):
        BilinearForm(space).add_integrator(ScalarDiffusionIntegrator()).assembly(format="dense")
    A = BilinearForm(space).add_integrator(ScalarDiffusionIntegrator()).assembly()
    with pytest.raises(ValueError):
        cg(A, torch.zeros(3, dtype=torch.float64, device="cuda"))
    with pytest.raises(TypeError):          # not an operator: nothing with `@` (solver/cg.py:10-14 accepts any SupportsMatmul)
        cg(object(), torch.zeros(A.shape[0], dtype=torch.float64, device="cuda"))


def test_fused_vs_sorted_coo_at_64(U):
    """config-2 family at n=64 (1.57M cells, 61M nnz, 315M COO entries through the radix sort):
    the fused row-owner path and the literal COO sort path give the same CSR."""
    from fealpy_b200.mesh import TetrahedronMesh
    from fealpy_b200.functionspace import LagrangeFESpace
    from fealpy_b200.fem import BilinearForm, ScalarDiffusionIntegrator, ScalarMassIntegrator
    mesh = TetrahedronMesh.from_box([0, 1, 0, 1, 0, 1], 64, 64, 64)
    space = LagrangeFESpace(mesh, 2)
    out = {}
    for path in ("auto", "coo"):
        bf = BilinearForm(space, assembly_path=path)
        bf.add_integrator(ScalarDiffusionIntegrator())
        bf.add_integrator(ScalarMassIntegrator())
        out[path] = bf.assembly()
    A, B = out["auto"], out["coo"]
    assert A.nnz == 60_859_905 == B.nnz and A.shape[0] == 2_146_689       # closed forms of SURVEY.md section 8
    assert torch.equal(A.crow, B.crow) and torch.equal(A.col, B.col)
    assert float((A.values - B.values).abs().max()) <= 1e-12 * float(A.values.abs().max())


def test_config2_full_size_properties(U):
    """BASELINE config 2 (tet P2, 128^3): sizes, pattern invariants, K 1 = 0, 1^T M 1 = 1, CG recovers x = 1."""
    from fealpy_b200.mesh import TetrahedronMesh
    from fealpy_b200.functionspace import LagrangeFESpace
    from fealpy_b200.fem import BilinearForm, ScalarDiffusionIntegrator, ScalarMassIntegrator
    from fealpy_b200.solver import cg
    mesh = TetrahedronMesh.from_box([0, 1, 0, 1, 0, 1], 128, 128, 128)
    space = LagrangeFESpace(mesh, 2)
    assert mesh.number_of_cells() == 12_582_912 and space.number_of_global_dofs() == 16_974_593
    K = BilinearForm(space).add_integrator(ScalarDiffusionIntegrator()).assembly()
    assert K.nnz == 484_609_025
    one = torch.ones(K.shape[0], dtype=torch.float64, device="cuda")
    assert float((K @ one).abs().max()) <= 1e-10 * float(K.values.abs().max())
    crow, col = K.crow, K.col
    assert int(crow[0]) == 0 and int(crow[-1]) == K.nnz and bool((crow[1:] > crow[:-1]).all())
    inner = torch.ones(K.nnz, dtype=torch.bool, device="cuda")
    inner[crow[1:-1]] = False                                   # first entry of every row but row 0
    assert bool((col[1:] > col[:-1])[inner[1:]].all()), "columns strictly ascending within rows"
    del inner
    vK = K.values.clone()
    del K
    Mm = BilinearForm(space).add_integrator(ScalarMassIntegrator()).assembly()
    assert abs(float((Mm @ one).sum()) - 1.0) <= 1e-11
    bf = BilinearForm(space)
    bf.add_integrator(ScalarDiffusionIntegrator())
    bf.add_integrator(ScalarMassIntegrator())
    A = bf.assembly()
    assert float((A.values - (vK + Mm.values)).abs().max()) <= 1e-12 * float(A.values.abs().max())   # linearity
    b = A @ one
    x, info = cg(A, b, returninfo=True)
    assert info["niter"] < 2000 and float((x - 1.0).abs().max()) < 1e-6


@pytest.mark.parametrize("p", [1, 2])
@pytest.mark.parametrize("world", [1, 2, 3])
def test_slab_partition_assembly_matches_global(p, world, U):
    """multi-GPU row partition, exercised rank by rank on one GPU: every rank's window matrix,
    restricted to its owned rows and mapped back to global ids, is bit-identical (pattern AND
    values) to the corresponding rows of the single-GPU matrix."""
    from fealpy_b200.mesh import TetrahedronMesh
    from fealpy_b200.functionspace import LagrangeFESpace
    from fealpy_b200.fem import BilinearForm, ScalarDiffusionIntegrator, ScalarMassIntegrator
    from fealpy_b200.parallel import SlabProblem
    box, dims = [0, 1, 0, 2, 0, 1], (7, 5, 4)
    mesh = TetrahedronMesh.from_box(box, *dims)
    space = LagrangeFESpace(mesh, p)

    def form(sp):
        bf = BilinearForm(sp)
        bf.add_integrator(ScalarDiffusionIntegrator())
        bf.add_integrator(ScalarMassIntegrator())
        return bf.assembly()
    G = form(space)
    gcrow, gcol, gval = G.crow.cpu().numpy(), G.col.cpu().numpy(), G.values.cpu().numpy()
    seen = np.zeros(G.shape[0], dtype=int)
    for r in range(world):
        sp = SlabProblem(box, *dims, p, world, r)
        part = sp.part
        if world == 1:      # closed-form numbering == sort-based numbering of the generic path
            assert torch.equal(sp.space.cell_to_dof(), space.cell_to_dof())
            assert torch.equal(sp.mesh.node, mesh.node) and torch.equal(sp.mesh.cell, mesh.cell)
        A = form(sp.space)
        crow, col, val = A.crow.cpu().numpy(), A.col.cpu().numpy(), A.values.cpu().numpy()
        l2g = part.local_to_global(np.arange(part.n_local))
        for lo, hi in (part.own_nodes, part.own_edges):
            if hi <= lo:
                continue
            g0, g1 = l2g[lo], l2g[hi - 1] + 1
            seen[g0:g1] += 1
            a, b = crow[lo], crow[hi]
            ga, gb = gcrow[g0], gcrow[g1]
            assert np.array_equal(crow[lo:hi + 1] - a, gcrow[g0:g1 + 1] - ga)
            assert np.array_equal(l2g[col[a:b]], gcol[ga:gb])
            assert np.array_equal(val[a:b], gval[ga:gb]), "owned rows must be bit-identical to the single-GPU rows"
    assert np.all(seen == 1)


def test_distributed_driver_single_rank_matches_cg(U):
    """the distributed CG driver with the CUDA ops on one rank == fb2_cg (same kernels, same order)"""
    from fealpy_b200.fem import BilinearForm, ScalarDiffusionIntegrator, ScalarMassIntegrator
    from fealpy_b200.parallel import SlabProblem, CudaCgOps, dist_cg
    from fealpy_b200.solver import cg
    sp = SlabProblem([0, 1, 0, 1, 0, 1], 8, 8, 8, 2, 1, 0)
    bf = BilinearForm(sp.space)
    bf.add_integrator(ScalarDiffusionIntegrator())
    bf.add_integrator(ScalarMassIntegrator())
    A = bf.assembly()
    b = A @ torch.ones(A.shape[0], dtype=torch.float64, device="cuda")
    x1, i1 = cg(A, b, returninfo=True)
    x2, i2 = dist_cg(CudaCgOps(A, sp.part.own_ranges), b, torch.zeros_like(b), sp.part.exchanges, check_every=4)
    assert i1["niter"] == i2["niter"]
    assert float((x1 - x2).abs().max()) <= 1e-13
    G = U  # keep fixture used


@pytest.mark.parametrize("case", C.BC_CASES, ids=lambda c: c["name"])
def test_dirichlet_operator_matches_reference_run(case, U):
    """DirichletBCOperator on the UNASSEMBLED form (fem/dirichlet_bc_operator.py:13-67) against the reference's own
    operator outputs (golden): Dirichlet set, init_solution(), apply(F, uh) and `op @ u` through the fused matrix-free
    product; then cg(op, ...) reproduces the golden solution of the constrained system (whole-boundary cases)"""
    from fealpy_b200.fem import (BilinearForm, LinearForm, ScalarDiffusionIntegrator, ScalarMassIntegrator, ScalarSourceIntegrator,
                                 DirichletBCOperator)
    from fealpy_b200.functionspace import LagrangeFESpace
    from fealpy_b200.solver import cg
    gold = G.load(case["name"])
    mesh = U.make_mesh(case, gold)
    space = LagrangeFESpace(mesh, case["p"])
    form = BilinearForm(space).add_integrator(ScalarDiffusionIntegrator())
    if case.get("reaction"):
        form.add_integrator(ScalarMassIntegrator())
    thr = U.torch_coef_func(C.THRESHOLDS[case["threshold"]]) if case.get("threshold") else None
    op = DirichletBCOperator(form, gd=U.torch_coef_func(C.kappa_cart), threshold=thr)
    assert np.array_equal(op.is_boundary_dof.cpu().numpy(), gold["op_isbd"])
    uh = op.init_solution()
    np.testing.assert_allclose(uh.cpu().numpy(), gold["op_uh"], atol=1e-15)
    F = LinearForm(space).add_integrator(ScalarSourceIntegrator(U.torch_coef_func(C.source_cart))).assembly()
    Fop = op.apply(F, uh)
    assert np.max(np.abs(Fop.cpu().numpy() - gold["op_F"])) <= 1e-12 * np.max(np.abs(gold["op_F"]))
    w = op @ U.t64(gold["op_u"])
    assert form._M is None and form.last_matfree == "fused"
    assert np.max(np.abs(w.cpu().numpy() - gold["op_w"])) <= 1e-12 * np.max(np.abs(gold["op_w"]))
    if not case.get("threshold"):      # whole boundary: the same constrained system as the golden DirichletBC solve of this case
        x = cg(op, Fop, atol=1e-14, rtol=1e-11)
        assert np.linalg.norm(x.cpu().numpy() - gold["x"]) / np.linalg.norm(gold["x"]) <= 1e-9


@pytest.mark.parametrize("case", C.BC_CASES, ids=lambda c: c["name"])
def test_poisson_source_dirichlet_cg(case, U):
    """rows f1/f2: LinearForm + ScalarSourceIntegrator, DirichletBC.apply, then cg -- against the
    reference run (golden) of the same Poisson problem"""
    from fealpy_b200.fem import BilinearForm, LinearForm, ScalarDiffusionIntegrator, ScalarSourceIntegrator, DirichletBC
    from fealpy_b200.functionspace import LagrangeFESpace
    from fealpy_b200.solver import cg
    gold = G.load(case["name"])
    mesh = U.make_mesh(case, gold)
    space = LagrangeFESpace(mesh, case["p"])
    from fealpy_b200.fem import ScalarMassIntegrator
    bform = BilinearForm(space).add_integrator(ScalarDiffusionIntegrator())
    if case.get("reaction"):
        bform.add_integrator(ScalarMassIntegrator())
    A = bform.assembly()
    U.assert_csr_matches(A, gold)
    F = LinearForm(space).add_integrator(ScalarSourceIntegrator(U.torch_coef_func(C.source_cart))).assembly()
    assert np.max(np.abs(F.cpu().numpy() - gold["F"])) <= 1e-13 * np.max(np.abs(gold["F"]))
    thr = U.torch_coef_func(C.THRESHOLDS[case["threshold"]]) if case.get("threshold") else None
    method = case.get("method")
    assert np.array_equal(space.is_boundary_dof(threshold=thr, method=method).cpu().numpy(), gold["isbd"])
    if thr is not None:       # the face-centroid and the interpolation-point selections differ on these cases
        assert np.array_equal(space.is_boundary_dof(threshold=thr, method="interp").cpu().numpy(), gold["isbd_val"])
        with pytest.raises(ValueError):
            space.is_boundary_dof(threshold=thr, method="nonsense")
    np.testing.assert_allclose(space.interpolation_points().cpu().numpy(), gold["ipoints"], atol=1e-14)
    bc = DirichletBC(space, gd=U.torch_coef_func(C.kappa_cart), threshold=thr, method=method)
    A2, F2 = bc.apply(A, F)
    assert np.max(np.abs(F2.cpu().numpy() - gold["F_bc"])) <= 1e-12 * np.max(np.abs(gold["F_bc"]))
    S = A2.to_scipy()
    assert S.has_sorted_indices
    # expected canonical pattern: interior-interior entries of the unconstrained pattern + unit diagonal
    # on boundary rows (the reference drops the same entries in DirichletBC._mul, dirichlet_bc.py:213-229)
    crow, col, isbd = gold["crow"], gold["col"], gold["isbd"]
    rows = np.repeat(np.arange(len(crow) - 1), np.diff(crow))
    keep = ~(isbd[rows] | isbd[col])
    exp_rows = np.concatenate([rows[keep], np.nonzero(isbd)[0]])
    exp_cols = np.concatenate([col[keep], np.nonzero(isbd)[0]])
    order = np.lexsort((exp_cols, exp_rows))
    exp_indptr = np.concatenate([[0], np.cumsum(np.bincount(exp_rows, minlength=len(crow) - 1))])
    assert np.array_equal(S.indptr, exp_indptr) and np.array_equal(S.indices, exp_cols[order])
    # values against the reference's constrained matrix (stored with its rounding-noise zeros eliminated)
    from scipy.sparse import csr_matrix
    R = csr_matrix((gold["Abc_data"], gold["Abc_indices"], gold["Abc_indptr"]), shape=S.shape)
    D = (S - R)
    assert (abs(D).max() if D.nnz else 0.0) <= 1e-12 * np.max(np.abs(gold["Abc_data"]))
    x, info = cg(A2, F2, returninfo=True, atol=1e-14, rtol=1e-11)
    assert np.linalg.norm(x.cpu().numpy() - gold["x"]) / np.linalg.norm(gold["x"]) <= 1e-10
    assert abs(info["niter"] - gold["info"]["niter"]) <= 2
    # inputs untouched
    U.assert_csr_matches(A, gold)


def test_config1_tri_p1_1024_full_size(U):
    """BASELINE config 1: tri P1 1024^2, diffusion q=3 (+ mass for CG): closed-form sizes, K 1 = 0, CG."""
    from fealpy_b200.mesh import TriangleMesh
    from fealpy_b200.functionspace import LagrangeFESpace
    from fealpy_b200.fem import BilinearForm, ScalarDiffusionIntegrator, ScalarMassIntegrator
    from fealpy_b200.solver import cg
    mesh = TriangleMesh.from_box([0, 1, 0, 1], 1024, 1024)
    space = LagrangeFESpace(mesh, 1)
    K = BilinearForm(space).add_integrator(ScalarDiffusionIntegrator(q=3)).assembly()
    assert (mesh.number_of_cells(), K.shape[0], K.nnz) == (2_097_152, 1_050_625, 7_346_177)
    one = torch.ones(K.shape[0], dtype=torch.float64, device="cuda")
    assert float((K @ one).abs().max()) <= 1e-11 * float(K.values.abs().max())
    bf = BilinearForm(space)
    bf.add_integrator(ScalarDiffusionIntegrator(q=3))
    bf.add_integrator(ScalarMassIntegrator(q=3))
    A = bf.assembly()
    x, info = cg(A, A @ one, returninfo=True)
    assert info["niter"] < 10000 and float((x - 1.0).abs().max()) < 1e-5


def test_config3_tri_p3_varcoef_full_size(U):
    """BASELINE config 3: tri P3 1024^2 with a variable coefficient (quadrature-loop kernel + gather path):
    a coefficient callable returning a constant must reproduce the constant-coefficient (fused) matrix."""
    from fealpy_b200.mesh import TriangleMesh
    from fealpy_b200.functionspace import LagrangeFESpace
    from fealpy_b200.fem import BilinearForm, ScalarDiffusionIntegrator
    from fealpy_b200.decorator import cartesian
    mesh = TriangleMesh.from_box([0, 1, 0, 1], 1024, 1024)
    space = LagrangeFESpace(mesh, 3)

    @cartesian
    def kappa(p):
        return 1.0 + 0.5 * torch.sin(2 * torch.pi * p[..., 0]) * torch.cos(2 * torch.pi * p[..., 1])

    bf = BilinearForm(space).add_integrator(ScalarDiffusionIntegrator(coef=kappa, q=6))
    A = bf.assembly()
    assert bf.last_path == "gather"
    assert (A.shape[0], A.nnz) == (9_443_329, 160_462_849)
    one = torch.ones(A.shape[0], dtype=torch.float64, device="cuda")
    assert float((A @ one).abs().max()) <= 1e-10 * float(A.values.abs().max())
    S = A.to_scipy()
    assert abs(S - S.T).max() <= 1e-12 * float(A.values.abs().max())

    @cartesian
    def two(p):
        return torch.full(p.shape[:-1], 2.0, dtype=torch.float64, device=p.device)
    B = BilinearForm(space).add_integrator(ScalarDiffusionIntegrator(coef=two, q=6)).assembly()
    Cc = BilinearForm(space).add_integrator(ScalarDiffusionIntegrator(coef=2.0, q=6)).assembly()
    assert torch.equal(B.col, Cc.col) and float((B.values - Cc.values).abs().max()) <= 1e-12 * float(Cc.values.abs().max())


def test_config4_tet_p1_elasticity_full_size(U):
    """BASELINE config 4: tet P1 x3 elasticity 128^3 (interleaved dofs): sizes, rigid translations in the
    kernel (K t = 0), symmetry sample."""
    from fealpy_b200.mesh import TetrahedronMesh
    from fealpy_b200.functionspace import LagrangeFESpace, TensorFunctionSpace
    from fealpy_b200.fem import BilinearForm, LinearElasticityIntegrator
    from fealpy_b200.material import LinearElasticMaterial
    mesh = TetrahedronMesh.from_box([0, 1, 0, 1, 0, 1], 128, 128, 128)
    space = TensorFunctionSpace(LagrangeFESpace(mesh, 1), shape=(-1, 3))
    mat = LinearElasticMaterial("m", elastic_modulus=1.0, poisson_ratio=0.3, hypo="3D")
    A = BilinearForm(space).add_integrator(LinearElasticityIntegrator(mat, q=4)).assembly()
    assert (A.shape[0], A.nnz) == (6_440_067, 286_222_473)
    scale = float(A.values.abs().max())
    for comp in range(3):
        t = torch.zeros(A.shape[0], dtype=torch.float64, device="cuda")
        t[comp::3] = 1.0
        assert float((A @ t).abs().max()) <= 1e-10 * scale
    x = torch.rand(A.shape[0], dtype=torch.float64, device="cuda")
    y = torch.rand(A.shape[0], dtype=torch.float64, device="cuda")
    assert abs(float(y @ (A @ x)) - float(x @ (A @ y))) <= 1e-9 * abs(float(y @ (A @ x)))


@pytest.mark.parametrize("batch_first", [False, True])
@pytest.mark.parametrize("case", [c for c in C.CASES if c.get("cg") and "bcg_B" in G.load(c["name"])], ids=lambda c: c["name"])
def test_batched_rhs_cg_matches_reference_run(case, batch_first, U):
    """cg with three right-hand sides at once against the real reference's batched solve (golden bcg_B / bcg_x / bcg_niter)"""
    from fealpy_b200.solver import cg
    gold = G.load(case["name"])
    mesh, space, bform, _ = U.make_form(case, gold)
    A = bform.assembly()
    Bm = gold["bcg_B"]
    x, info = cg(A, U.t64(Bm.T.copy() if batch_first else Bm), batch_first=batch_first, returninfo=True, atol=1e-14, rtol=1e-12)
    xg = x.cpu().numpy().T if batch_first else x.cpu().numpy()
    assert abs(info["niter"] - gold["info"]["bcg_niter"]) <= 1
    assert np.linalg.norm(xg - gold["bcg_x"]) / np.linalg.norm(gold["bcg_x"]) <= 1e-10


@pytest.mark.parametrize("batch_first", [False, True])
def test_batched_rhs_cg_matches_oracle(batch_first, U):
    """b of shape (dof, batch): per-column alpha/beta, joint stopping test (solver/cg.py:58-121)"""
    from oracle import fem_oracle as O
    from fealpy_b200.solver import cg
    case = C.by_name("tet_p2_3x2x1_diffmass")
    gold = G.load(case["name"])
    mesh, space, bform, _ = U.make_form(case, gold)
    A = bform.assembly()
    n = A.shape[0]
    rng = np.random.default_rng(3)
    Bm = rng.standard_normal((n, 3))
    crow, col, val = gold["crow"], gold["col"], gold["values"]
    xo, oinfo = O.cg(lambda v: np.stack([O.csr_matvec(crow, col, val, v[:, k]) for k in range(v.shape[1])], axis=1), Bm,
                     atol=1e-14, rtol=1e-12)
    bt = U.t64(Bm.T.copy() if batch_first else Bm)
    x, info = cg(A, bt, batch_first=batch_first, returninfo=True, atol=1e-14, rtol=1e-12)
    xg = x.cpu().numpy().T if batch_first else x.cpu().numpy()
    assert abs(info["niter"] - oinfo["niter"]) <= 1
    assert np.linalg.norm(xg - xo) / np.linalg.norm(xo) <= 1e-10
    y = (A @ U.t64(Bm)).cpu().numpy()                                     # SpMM against the oracle SpMV
    ref = np.stack([O.csr_matvec(crow, col, val, Bm[:, k]) for k in range(3)], axis=1)
    assert np.max(np.abs(y - ref)) <= 1e-12 * np.max(np.abs(ref))
    # wider batches (8 columns per pass of A, then the remainder) and a solve cut by maxit between two host polls: the
    # device-side stopping test must freeze x at the stopping iteration (the host reads the state every 8 iterations)
    for nb, maxit in ((8, 5), (11, 13), (1, 3)):
        Bw = rng.standard_normal((n, nb))
        mv = lambda v: np.stack([O.csr_matvec(crow, col, val, v[:, k]) for k in range(v.shape[1])], axis=1)
        yw = (A @ U.t64(Bw)).cpu().numpy()
        assert np.max(np.abs(yw - mv(Bw))) <= 1e-12 * np.max(np.abs(yw))
        xo, oinfo = O.cg(mv, Bw, atol=0.0, rtol=0.0, maxit=maxit)
        bt = U.t64(Bw.T.copy() if batch_first else Bw)
        x, info = cg(A, bt, batch_first=batch_first, returninfo=True, atol=0.0, rtol=0.0, maxit=maxit)
        xg = x.cpu().numpy().T if batch_first else x.cpu().numpy()
        assert info["niter"] == oinfo["niter"] == maxit
        assert np.linalg.norm(xg - xo) / np.linalg.norm(xo) <= 1e-10
        assert abs(info["residual"] - oinfo["residual"]) <= 1e-9 * abs(oinfo["residual"])


def _relabelled_mesh(kind, dims, seed):
    """from_box topology with a random node relabelling, shuffled cells and jittered geometry: nothing of the
    closed-form numbering survives, the generic topology path must agree with the oracle"""
    from oracle import fem_oracle as O
    rng = np.random.default_rng(seed)
    if kind == "tri":
        node, cell = O.tri_from_box([0, 1, 0, 1], *dims)
    else:
        node, cell = O.tet_from_box([0, 1, 0, 1, 0, 1], *dims)
    node = C.perturb(node, dims, seed)
    perm = rng.permutation(len(node))                # new id of old node k
    node2 = np.empty_like(node)
    node2[perm] = node
    cell2 = perm[cell][rng.permutation(len(cell))].astype(np.int32)
    return node2, cell2


@pytest.mark.parametrize("kind,dims,p", [("tri", (5, 4), 1), ("tri", (4, 5), 2), ("tri", (4, 3), 3),
                                         ("tet", (3, 2, 2), 1), ("tet", (2, 3, 2), 2), ("tet", (2, 2, 2), 3),
                                         ("tri", (1, 1), 3), ("tet", (1, 1, 1), 3)])
def test_relabelled_mesh_matches_oracle(kind, dims, p, U):
    from oracle import fem_oracle as O
    from fealpy_b200.mesh import TriangleMesh, TetrahedronMesh
    from fealpy_b200.functionspace import LagrangeFESpace
    from fealpy_b200.fem import BilinearForm, ScalarDiffusionIntegrator, ScalarMassIntegrator
    from fealpy_b200.solver import cg
    node, cell = _relabelled_mesh(kind, dims, seed=17 + p)
    om = O.Mesh(node, cell)
    mesh = (TriangleMesh if kind == "tri" else TetrahedronMesh)(U.t64(node), U.t64(cell))
    space = LagrangeFESpace(mesh, p)
    assert np.array_equal(mesh.edge.cpu().numpy(), om.edge)
    c2d = om.cell_to_ipoint(p)
    assert np.array_equal(space.cell_to_dof().cpu().numpy(), c2d)
    gdof = om.number_of_global_ipoints(p)
    assert space.number_of_global_dofs() == gdof
    crow, col, val = O.assemble([(O.diffusion_element(om, p), c2d), (O.mass_element(om, p), c2d)], gdof)
    for path in ("auto", "gather", "coo"):
        bf = BilinearForm(space, assembly_path=path)
        bf.add_integrator(ScalarDiffusionIntegrator())
        bf.add_integrator(ScalarMassIntegrator())
        A = bf.assembly()
        assert np.array_equal(A.crow.cpu().numpy(), crow) and np.array_equal(A.col.cpu().numpy(), col), path
        assert np.max(np.abs(A.values.cpu().numpy() - val)) <= 1e-12 * np.max(np.abs(val)), path
    b = O.csr_matvec(crow, col, val, np.ones(gdof))
    xo, oinfo = O.cg(lambda v: O.csr_matvec(crow, col, val, v), b, atol=1e-14, rtol=1e-12)
    x, info = cg(A, U.t64(b), atol=1e-14, rtol=1e-12, returninfo=True)
    assert abs(info["niter"] - oinfo["niter"]) <= 2
    assert np.linalg.norm(x.cpu().numpy() - xo) / np.linalg.norm(xo) <= 1e-10
    assert np.array_equal(space.is_boundary_dof().cpu().numpy(), O.boundary_dof_flag(om, p))


def test_dirichlet_problem_at_config2_size(U):
    """row f1 at BASELINE config 2's size (tet P2 128^3: 2 146 689 nodes -- face keys need the two-leg sort; 17 M dofs): the
    boundary dof count is the closed form (2n+1)^3 - (2n-1)^3, boundary rows of the constrained matrix are unit rows, the
    right-hand side carries g on the boundary, interior rows keep their interior entries, and CG reduces the residual"""
    from fealpy_b200.mesh import TetrahedronMesh
    from fealpy_b200.functionspace import LagrangeFESpace
    from fealpy_b200.fem import BilinearForm, ScalarDiffusionIntegrator, DirichletBC, LinearForm, ScalarSourceIntegrator
    from fealpy_b200.solver import cg
    n = 128
    mesh = TetrahedronMesh.from_box([0, 1, 0, 1, 0, 1], n, n, n)
    space = LagrangeFESpace(mesh, 2)
    A = BilinearForm(space).add_integrator(ScalarDiffusionIntegrator()).assembly()

    def g(p):
        return torch.sin(p[..., 0]) * torch.cos(p[..., 1]) + p[..., 2]
    g.coordtype = "cartesian"
    F = LinearForm(space).add_integrator(ScalarSourceIntegrator(1.0)).assembly()
    assert abs(float(F.sum()) - 1.0) <= 1e-12                      # integral of 1 over the unit cube
    bc = DirichletBC(space, gd=g)
    bd = bc.is_boundary_dof
    assert int(bd.sum()) == (2 * n + 1) ** 3 - (2 * n - 1) ** 3
    ip = space.interpolation_points(index=bd)
    on_boundary = ((ip[bd] == 0.0) | (ip[bd] == 1.0)).any(dim=1)
    assert bool(on_boundary.all())
    A2, F2 = bc.apply(A, F)
    assert torch.equal(F2[bd], g(ip[bd]))
    rows = bd.nonzero().reshape(-1)
    assert bool(((A2.crow[rows + 1] - A2.crow[rows]) == 1).all())
    assert torch.equal(A2.col[A2.crow[rows]].long(), rows) and bool((A2.values[A2.crow[rows]] == 1.0).all())
    assert not bool(bd[A2.col.long()][~bd.repeat_interleave(A2.crow[1:] - A2.crow[:-1])].any()), "no boundary column in an interior row"
    x, info = cg(A2, F2, maxit=60, returninfo=True)
    assert info["niter"] == 60 and float((F2 - A2 @ x).norm()) < float(F2.norm())      # 60 CG steps reduce the residual


@pytest.mark.gpu
@pytest.mark.parametrize("dims,p", [((5, 4, 3), 3), ((6, 5, 4), 2)])
def test_wide_face_keys_match_oracle(dims, p, U, monkeypatch):
    """tetrahedral meshes with more than 2^21 nodes (from_box 128^3) cannot pack a sorted vertex triple into one 64-bit
    key: faces are then sorted in two stable legs (csrc/topo.cu).  FB2_TOPO_FORCE_WIDE=1 takes that path on a relabelled
    small mesh: face table, cell2face, boundary flags and the P3 numbering (which uses the faces) against the oracle, and
    bit-identical to the single-key path."""
    from oracle import fem_oracle as O
    from fealpy_b200.mesh import TetrahedronMesh
    from fealpy_b200.functionspace import LagrangeFESpace
    node, cell = _relabelled_mesh("tet", dims, seed=91 + p)
    om = O.Mesh(node, cell)
    ref = TetrahedronMesh(U.t64(node), U.t64(cell))
    f0, c2f0 = ref.face.clone(), ref.cell2face.clone()
    monkeypatch.setenv("FB2_TOPO_FORCE_WIDE", "1")
    mesh = TetrahedronMesh(U.t64(node), U.t64(cell))
    assert torch.equal(mesh.face, f0) and torch.equal(mesh.cell2face, c2f0)
    assert np.array_equal(mesh.face.cpu().numpy(), om.face) and np.array_equal(mesh.cell2face.cpu().numpy(), om.cell2face)
    space = LagrangeFESpace(mesh, p)
    assert np.array_equal(space.cell_to_dof().cpu().numpy(), om.cell_to_ipoint(p))
    assert np.array_equal(space.is_boundary_dof().cpu().numpy(), O.boundary_dof_flag(om, p))


@pytest.mark.gpu
@pytest.mark.parametrize("kind,dims,p,tile", [("tet", (6, 5, 4), 2, 2560), ("tet", (5, 4, 3), 3, 2560), ("tri", (31, 17), 1, 512),
                                              ("tri", (9, 8), 3, 640), ("tet", (7, 6, 5), 1, 256)])
def test_asm4_schedule_invariants(kind, dims, p, tile, U, monkeypatch):
    """the batch schedule of the v4 numeric kernel (csrc/assemble.cu asm4_schedule_kernel), checked entry by entry:
    every (row, cell, i) pair of the mesh exactly once; a batch holds one local index and 32 DIFFERENT rows;
    every row meets its cells in (i, cell) order; the number of batches per (tile, i) is the minimum
    max(ceil(n/32), longest run in one row); and the first-touch masks mark exactly the first write of every value."""
    from fealpy_b200.mesh import TetrahedronMesh, TriangleMesh
    from fealpy_b200.functionspace import LagrangeFESpace
    from fealpy_b200.fem import bilinear_form as bfm
    monkeypatch.setattr(bfm, "ASM4_TILE", tile)
    mesh = (TetrahedronMesh.from_box([0, 1, 0, 1, 0, 1], *dims) if kind == "tet" else TriangleMesh.from_box([0, 1, 0, 1], *dims))
    space = LagrangeFESpace(mesh, p)
    sym = bfm.symbolic_pattern(space)
    pl = bfm.asm4_plan(space)
    c2d = space.cell_to_dof().cpu().numpy()
    NC, L = c2d.shape
    crow = sym["crow"].cpu().numpy()
    blk_row = pl["blk_row"].cpu().numpy()[:pl["ntile"] + 1]
    batch_ptr = pl["batch_ptr"].cpu().numpy()
    batch_i = pl["batch_i"].cpu().numpy()
    # the schedule is one packed block per batch (include/fealpy_b200.h); ent_* are strided views of it
    ent_cell = pl["ent_cell"].contiguous().cpu().numpy().reshape(-1, 32)
    ent_word = pl["ent_base"].contiguous().cpu().numpy().astype(np.uint32).reshape(-1, 32)
    assert np.array_equal(pl["header_i"].cpu().numpy(), batch_i), "the block header repeats the batch's local index"
    ent_base, ent_first = ent_word & 0xfff, ent_word >> 12
    assert batch_ptr[0] == 0 and batch_ptr[-1] == ent_cell.shape[0] == batch_i.shape[0]
    sb = sym["slot_bytes"]
    raw = pl["ent_slots"].contiguous().cpu().numpy().view(np.uint8 if sb == 1 else np.uint16)
    ent_slot = raw.reshape(ent_cell.shape[0], 32, -1)[:, :, :L].astype(np.int64)     # (batch, lane, column) -> position in the row
    seen = np.zeros((NC, L), dtype=int)
    for t in range(pl["ntile"]):
        r0, r1 = blk_row[t], blk_row[t + 1]
        v0 = crow[r0]
        base_to_row = {int(crow[r] - v0): r for r in range(r0, r1) if crow[r + 1] > crow[r]}
        last = {}                                   # (row, i) -> last cell seen, in batch order
        per_i = {}
        touched = np.zeros(int(crow[r1] - v0), dtype=bool)
        for b in range(batch_ptr[t], batch_ptr[t + 1]):
            i = int(batch_i[b])
            live = ent_cell[b] >= 0
            assert live.any(), "empty batch"
            cells, bases = ent_cell[b][live], ent_base[b][live]
            rows = np.array([base_to_row[int(x)] for x in bases])
            assert len(set(rows.tolist())) == rows.size, "a row twice in one batch"
            assert np.array_equal(c2d[cells, i], rows), "entry does not belong to its row / local index"
            seen[cells, i] += 1
            pos = bases.astype(np.int64)[:, None] + ent_slot[b][live]                  # tile-local value index of every column
            first = ((ent_first[b][live][:, None] >> np.arange(L)) & 1).astype(bool)
            assert np.array_equal(first, ~touched[pos]), "first-touch mask != (value not yet written in execution order)"
            assert np.unique(pos).size == pos.size
            touched[pos] = True
            for r, c in zip(rows.tolist(), cells.tolist()):
                assert last.get((r, i), -1) < c, "cells of a row out of order"
                last[(r, i)] = c
            per_i.setdefault(i, []).append(rows)
        assert touched.all(), "a value of the tile is never written (the kernel does not zero-fill)"
        assert sorted(per_i) == sorted(set(per_i)), "batches of one local index are contiguous"
        for i, lst in per_i.items():
            allr = np.concatenate(lst)
            mult = np.unique(allr, return_counts=True)[1].max()
            assert len(lst) == max(-(-allr.size // 32), mult), "not the minimum number of batches"
    assert np.all(seen == 1), "every (cell, local index) pair exactly once"


def test_sparse_edge_cases(U):
    """ragged / degenerate COO inputs through K2 (sort + segmented reduce) against the oracle's coalesce/tocsr:
    no entries at all, one entry repeated, empty leading / trailing / interior rows, a single dense row."""
    from oracle import fem_oracle as O
    from fealpy_b200.sparse import COOTensor
    rng = np.random.default_rng(5)
    cases = {
        "empty": (np.zeros(0, int), np.zeros(0, int), np.zeros(0), (4, 4)),
        "one_entry_x100": (np.full(100, 2), np.full(100, 1), rng.standard_normal(100), (3, 5)),
        "ragged": (np.array([5, 5, 2, 2, 2, 7, 5]), np.array([0, 9, 3, 3, 1, 8, 0]), rng.standard_normal(7), (10, 10)),
        "dense_row": (np.full(4097, 1), rng.permutation(4097), rng.standard_normal(4097), (3, 4097)),
        "1x1": (np.zeros(3, int), np.zeros(3, int), np.array([1.0, 2.0, 4.0]), (1, 1)),
    }
    for name, (r, c, v, shape) in cases.items():
        idx = torch.tensor(np.stack([r, c]), dtype=torch.int32, device="cuda")       # col keeps the index dtype (coo_tensor.py:137-157)
        coo = COOTensor(idx, U.t64(v), shape, is_coalesced=False)
        if len(r) > len(set(zip(r.tolist(), c.tolist()))):
            with pytest.raises(ValueError):          # duplicates: the reference requires coalesce() before tocsr()
                coo.tocsr()
        A = coo.coalesce().tocsr()
        crow, col, val = O.tocsr(*O.coalesce(r, c, v), shape[0])
        assert A.shape == shape and A.nnz == len(col), name
        assert np.array_equal(A.crow.cpu().numpy(), crow), name
        assert np.array_equal(A.col.cpu().numpy(), col), name
        assert np.array_equal(A.values.cpu().numpy(), val), name            # same left-to-right order as np.add.at: bit-exact
        x = rng.standard_normal(shape[1])
        y = (A @ U.t64(x)).cpu().numpy()
        yref = np.zeros(shape[0])
        np.add.at(yref, np.repeat(np.arange(shape[0]), np.diff(crow)), val * x[col])
        assert np.allclose(y, yref, rtol=0, atol=1e-13 * (1 + np.abs(val).sum())), name


@pytest.mark.parametrize("kind,n,prio,hypo", [("tet", 12, False, "3D"), ("tet", 9, True, "3D"), ("tri", 48, False, "plane_strain"),
                                              ("tri", 40, True, "plane_stress")])
def test_elasticity_p1_fused_agrees_with_ke_paths(kind, n, prio, hypo, U):
    """LinearElasticityIntegrator on a P1 tensor space: the fused path (entries from per-cell grad-lambda records, no K_e in
    HBM) against the K_e gather path and the literal COO pipeline at a size the golden cases do not reach; plus
    the rigid-body null space (translations) and symmetry as size-independent properties."""
    from fealpy_b200.mesh import TriangleMesh, TetrahedronMesh
    from fealpy_b200.functionspace import LagrangeFESpace, TensorFunctionSpace
    from fealpy_b200.fem import BilinearForm, LinearElasticityIntegrator
    from fealpy_b200.material import LinearElasticMaterial
    if kind == "tet":
        mesh = TetrahedronMesh.from_box([0, 1, 0, 2, 0, 1], n, n + 1, n - 1)
    else:
        mesh = TriangleMesh.from_box([0, 2, 0, 1], n, n - 3)
    GD = mesh.geo_dimension()
    space = TensorFunctionSpace(LagrangeFESpace(mesh, 1), shape=(GD, -1) if prio else (-1, GD))
    assert bool(space.dof_priority) == prio
    mat = LinearElasticMaterial("m", elastic_modulus=2.0, poisson_ratio=0.27, hypo=hypo)
    out = {}
    for path in ("auto", "gather", "coo"):
        bf = BilinearForm(space, assembly_path=path)
        bf.add_integrator(LinearElasticityIntegrator(mat, q=3))
        out[path] = (bf.assembly(), bf.last_path)
    A, used = out["auto"]
    assert used == "fused-elasticity-p1" and out["gather"][1] == "gather"
    scale = float(A.values.abs().max())
    for path in ("gather", "coo"):
        B = out[path][0]
        assert torch.equal(A.crow, B.crow) and torch.equal(A.col, B.col), path
        assert float((A.values - B.values).abs().max()) <= 1e-12 * scale, path
    gdof = space.scalar_space.number_of_global_dofs()
    for a in range(GD):                          # a rigid translation produces no strain
        t = torch.zeros(GD * gdof, dtype=torch.float64, device="cuda")
        if prio:
            t[a * gdof:(a + 1) * gdof] = 1.0
        else:
            t[a::GD] = 1.0
        assert float((A @ t).abs().max()) <= 1e-11 * scale
    x = torch.rand(GD * gdof, dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    y = torch.rand(GD * gdof, dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(4))
    assert abs(float(y @ (A @ x) - x @ (A @ y))) <= 1e-10 * scale * GD * gdof


@pytest.mark.parametrize("kind,dims,p", [("tet", (5, 4, 3), 2), ("tet", (3, 3, 2), 3), ("tri", (12, 9), 3)])
def test_symbolic_without_stash_matches(kind, dims, p, U, monkeypatch):
    """the symbolic phase ranks every row's candidate columns once and replays the ranks from a scratch buffer; when
    that buffer cannot be allocated the fill pass ranks again -- both routes must give the same pattern and slot map"""
    from fealpy_b200.mesh import TriangleMesh, TetrahedronMesh
    from fealpy_b200.functionspace import LagrangeFESpace
    from fealpy_b200.fem import bilinear_form as bfm
    Mesh = TetrahedronMesh if kind == "tet" else TriangleMesh
    box = [0, 1, 0, 1, 0, 1] if kind == "tet" else [0, 1, 0, 1]
    a = bfm.symbolic_pattern(LagrangeFESpace(Mesh.from_box(box, *dims), p))
    monkeypatch.setenv("FB2_SYM_NO_STASH", "1")
    b = bfm.symbolic_pattern(LagrangeFESpace(Mesh.from_box(box, *dims), p))
    for k in ("crow", "col", "slots", "adj_ptr", "adj_pair"):
        assert torch.equal(a[k], b[k]), k
    assert a["nnz"] == b["nnz"] and a["max_row"] == b["max_row"] and a["slot_bytes"] == b["slot_bytes"]


# ---------------------------------------------------------------------------------------------------------------
# round 2: public geometry / basis arrays of the CUDA code against the reference's OWN test vectors
# (test/mesh/tetrahedron_mesh_data.py:1819,1994,2176, test/mesh/triangle_mesh_data.py:60,68,104,
#  test/backend/backend_data.py:12,198, test/functionspace/lagrange_fe_space_data.py)
# ---------------------------------------------------------------------------------------------------------------
def test_reference_geometry_vectors_on_device(U):
    from fealpy_b200.mesh import TriangleMesh, TetrahedronMesh
    from fealpy_b200.functionspace import LagrangeFESpace
    from fealpy_b200.quadrature import simplex_quadrature
    RT = G.ref_testdata()
    # backend primitives
    m = TriangleMesh(U.t64(RT["bk_tri2d_node"]), U.t64(RT["bk_tri2d_cell"].astype(np.int32)))
    p = int(RT["bk_tri2d_p"])
    np.testing.assert_allclose(m.entity_measure("cell").cpu().numpy(), RT["bk_tri2d_simple_measure"], atol=1e-14)
    np.testing.assert_allclose(m.cell_area().cpu().numpy(), RT["bk_tri2d_simple_measure"], atol=1e-14)
    np.testing.assert_allclose(m.grad_lambda().cpu().numpy(), RT["bk_tri2d_triangle_grad_lambda_2d"], atol=1e-14)
    np.testing.assert_allclose(m.shape_function(RT["bk_tri2d_bcs"], p).cpu().numpy(), RT["bk_tri2d_simple_shape_function"], atol=1e-14)
    np.testing.assert_allclose(m.grad_shape_function(RT["bk_tri2d_bcs"], p, variables="u").cpu().numpy(),
                               RT["bk_tri2d_simple_grad_shape_function"], atol=1e-14)
    t = TetrahedronMesh(U.t64(RT["bk_tet_node"]), U.t64(RT["bk_tet_cell"].astype(np.int32)))
    np.testing.assert_allclose(t.grad_lambda().cpu().numpy(), RT["bk_tet_tetrahedron_grad_lambda_3d"], atol=1e-14)
    # tetrahedron mesh vectors on from_box(3, 2, 1)
    t = TetrahedronMesh.from_box([0, 1, 0, 1, 0, 1], 3, 2, 1)
    np.testing.assert_allclose(t.cell_volume().cpu().numpy(), RT["tet_321_cm"], atol=1e-14)
    np.testing.assert_allclose(t.entity_measure("cell").cpu().numpy(), RT["tet_321_cm"], atol=1e-14)
    np.testing.assert_allclose(t.grad_lambda().cpu().numpy(), RT["tet_321_glambda"], atol=1e-14)
    t1 = TetrahedronMesh.from_box([0, 1, 0, 1, 0, 1], 1, 1, 1)
    bcs, _ = simplex_quadrature(3, 3).get_quadrature_points_and_weights()
    np.testing.assert_allclose(t1.grad_shape_function(bcs, 2).cpu().numpy(), RT["tet_111_gsf_u_p2_q3"], atol=1e-14)   # default 'u'
    gx = t1.grad_shape_function(bcs, 2, variables="x").cpu().numpy()
    np.testing.assert_allclose(np.transpose(gx, (1, 0, 2, 3)), RT["tet_111_gsf_x_p2_q3"], atol=1e-13)    # reference layout (NQ, NC, l, GD)
    # triangle mesh vectors
    m = TriangleMesh.from_box([0, 1, 0, 1], 2, 2)
    bcs, _ = simplex_quadrature(2, 3).get_quadrature_points_and_weights()
    np.testing.assert_allclose(m.grad_shape_function(bcs, 2).cpu().numpy(), RT["tri_22_gphi_p2_q3"], atol=1e-13)      # default 'x'
    m2 = TriangleMesh.from_box([-1, 1, -1, 1], 2, 2)
    np.testing.assert_allclose(m2.grad_lambda().cpu().numpy(), RT["tri_glambda_m11_22"], atol=1e-14)
    # LagrangeFESpace.basis / grad_basis (test/functionspace/lagrange_fe_space_data.py)
    m = TriangleMesh.from_box([0, 1, 0, 1], 1, 1)
    space = LagrangeFESpace(m, 2)
    bcs = RT["lfs_box1_p2_bcs"]
    np.testing.assert_allclose(space.basis(bcs).cpu().numpy(), RT["lfs_box1_p2_basis"], atol=1e-14)
    np.testing.assert_allclose(space.grad_basis(bcs).cpu().numpy(), RT["lfs_box1_p2_grad_basis"], atol=1e-13)
    assert np.array_equal(space.is_boundary_dof().cpu().numpy(), RT["lfs_box1_p2_is_boundary_dof"])
    # index= sub-selection and the oracle on a jittered mesh
    from oracle import fem_oracle as O
    case = C.by_name("tet_p2_4_diffmass_jit")
    gold = G.load(case["name"])
    mesh = U.make_mesh(case, gold)
    om = O.Mesh(gold["node"], gold["cell"])
    np.testing.assert_allclose(mesh.entity_measure("cell").cpu().numpy(), om.cell_measure(), rtol=1e-13)
    np.testing.assert_allclose(mesh.grad_lambda().cpu().numpy(), om.grad_lambda(), rtol=1e-12, atol=1e-13)
    idx = torch.tensor([5, 0, 17], device="cuda")
    np.testing.assert_allclose(mesh.grad_lambda(index=idx).cpu().numpy(), om.grad_lambda()[[5, 0, 17]], rtol=1e-12, atol=1e-13)
    bcs, _ = O.quadrature(3, 4)
    np.testing.assert_allclose(LagrangeFESpace(mesh, 2).grad_basis(bcs).cpu().numpy(), O.grad_basis(om, 2, bcs), rtol=1e-11, atol=1e-12)


def test_user_integrator_written_against_basis_api(U):
    """plug-in point (i) of SURVEY 8(b): an integrator that only uses space.grad_basis / mesh.entity_measure /
    quadrature_formula and returns (NC, l, l) -- assembled through the gather path, equal to the built-in one"""
    from fealpy_b200.fem import BilinearForm, ScalarDiffusionIntegrator
    from fealpy_b200.fem.integrators import Integrator

    class MyDiffusion(Integrator):
        def assembly(self, space, indices=None):
            mesh = space.mesh
            bcs, ws = mesh.quadrature_formula(space.p + 3).get_quadrature_points_and_weights()
            gphi = space.grad_basis(bcs, variable="x")                       # (NC, NQ, l, GD)
            cm = mesh.entity_measure("cell")
            w = torch.as_tensor(ws, device=gphi.device)
            return torch.einsum("q,c,cqid,cqjd->cij", w, cm, gphi, gphi)

    case = C.by_name("tet_p2_4_diffmass_jit")
    gold = G.load(case["name"])
    mesh, space, _, _ = U.make_form(case, gold)
    A = BilinearForm(space).add_integrator(MyDiffusion()).assembly()
    B = BilinearForm(space).add_integrator(ScalarDiffusionIntegrator()).assembly()
    assert torch.equal(A.col, B.col) and torch.equal(A.crow, B.crow)
    assert float((A.values - B.values).abs().max()) <= 1e-12 * float(B.values.abs().max())


@pytest.mark.parametrize("n", [16, 24])
def test_tet_p2_against_oracle_full_values(n, U):
    """every CSR value of a jittered tet P2 diffusion + mass matrix at 16^3 / 24^3 against the oracle (itself pinned to the
    reference on the smaller golden cases): pattern bit-exact, values <= 1e-12 max|A|; CG solution <= 1e-10"""
    from oracle import fem_oracle as O
    from fealpy_b200.mesh import TetrahedronMesh
    from fealpy_b200.functionspace import LagrangeFESpace
    from fealpy_b200.fem import BilinearForm, ScalarDiffusionIntegrator, ScalarMassIntegrator
    from fealpy_b200.solver import cg
    node, cell = O.tet_from_box([0, 1, 0, 1, 0, 1], n, n, n)
    node = C.perturb(node, (n, n, n), seed=100 + n)
    om = O.Mesh(node, cell)
    c2d = om.cell_to_ipoint(2)
    gdof = om.number_of_global_ipoints(2)
    crow, col, val = O.assemble([(O.diffusion_element(om, 2), c2d), (O.mass_element(om, 2), c2d)], gdof)
    mesh = TetrahedronMesh(U.t64(node), U.t64(cell.astype(np.int32)))
    space = LagrangeFESpace(mesh, 2)
    assert np.array_equal(space.cell_to_dof().cpu().numpy(), c2d)
    for path in ("auto", "gather"):
        bf = BilinearForm(space, assembly_path=path)
        bf.add_integrator(ScalarDiffusionIntegrator())
        bf.add_integrator(ScalarMassIntegrator())
        A = bf.assembly()
        assert np.array_equal(A.crow.cpu().numpy(), crow) and np.array_equal(A.col.cpu().numpy(), col)
        got = A.values.cpu().numpy()
        assert np.max(np.abs(got - val)) <= 1e-12 * np.max(np.abs(val)), path
    if n == 16:
        b = O.csr_matvec(crow, col, val, np.ones(gdof))
        xo, oinfo = O.cg(lambda v: O.csr_matvec(crow, col, val, v), b)
        x, info = cg(A, U.t64(b), returninfo=True)
        assert abs(info["niter"] - oinfo["niter"]) <= 1
        assert np.linalg.norm(x.cpu().numpy() - xo) / np.linalg.norm(xo) <= 1e-10


def test_share_pattern_and_out_buffer(U):
    from fealpy_b200.fem import BilinearForm, ScalarDiffusionIntegrator, ScalarMassIntegrator
    case = C.by_name("tet_p2_3x2x1_diffmass")
    gold = G.load(case["name"])
    mesh, space, bform, _ = U.make_form(case, gold)
    A1 = bform.assembly()
    A2 = BilinearForm(space).add_integrator(ScalarMassIntegrator()).assembly()
    assert A1.col.data_ptr() == A2.col.data_ptr()                   # default: the cached pattern is shared (documented)
    bf = BilinearForm(space, share_pattern=False)
    bf.add_integrator(ScalarDiffusionIntegrator())
    bf.add_integrator(ScalarMassIntegrator())
    A3 = bf.assembly()
    assert A3.col.data_ptr() != A1.col.data_ptr() and A3.crow.data_ptr() != A1.crow.data_ptr()
    A3.col.zero_()                                                  # a private copy: editing it must not touch the cache
    U.assert_csr_matches(bform.assembly(), gold)
    out = torch.empty(A1.nnz, dtype=torch.float64, device="cuda")
    A4 = bform.assembly(out=out)
    assert A4.values.data_ptr() == out.data_ptr()
    U.assert_csr_matches(A4, gold)
    with pytest.raises(ValueError):
        bform.assembly(out=torch.empty(3, dtype=torch.float64, device="cuda"))


def test_operator_cg_and_dirichlet_operator(U):
    """cg with operator-valued A (solver/cg.py:10-14): the unassembled BilinearForm (matrix-free product), a user operator,
    an operator-valued M; DirichletBCOperator (fem/dirichlet_bc_operator.py:13-67) against DirichletBC.apply"""
    from fealpy_b200.fem import (BilinearForm, LinearForm, ScalarDiffusionIntegrator, ScalarMassIntegrator, ScalarSourceIntegrator,
                                 DirichletBC, DirichletBCOperator)
    from fealpy_b200.functionspace import LagrangeFESpace
    from fealpy_b200.solver import cg
    case = C.by_name("tet_p2_3x2x1_diffmass")
    gold = G.load(case["name"])
    mesh, space, bform, _ = U.make_form(case, gold)
    b = U.t64(gold["b"])
    x_free, i_free = cg(bform, b, returninfo=True)                   # matrix-free: bform._M is None
    assert bform._M is None
    assert abs(i_free["niter"] - gold["info"]["niter"]) <= 1
    assert np.linalg.norm(x_free.cpu().numpy() - gold["x"]) / np.linalg.norm(gold["x"]) <= 1e-10
    A = bform.assembly()

    class Op:
        shape = A.shape

        def __matmul__(self, v):
            return A @ v

    class JacobiOp:
        def __init__(self, d): self.d = d
        def __matmul__(self, v): return v / self.d

    x_op, i_op = cg(Op(), b, returninfo=True)
    x_csr, i_csr = cg(A, b, returninfo=True)
    assert i_op["niter"] == i_csr["niter"] and float((x_op - x_csr).abs().max()) <= 1e-12
    x_m, i_m = cg(A, b, M=JacobiOp(A.diagonal()), returninfo=True)
    x_d, i_d = cg(A, b, M=1.0 / A.diagonal(), returninfo=True)
    assert abs(i_m["niter"] - i_d["niter"]) <= 1 and float((x_m - x_d).abs().max()) <= 1e-10
    with pytest.raises(TypeError):
        cg(object(), b)
    # DirichletBCOperator on a Poisson problem == DirichletBC.apply + cg
    bcase = C.by_name("bc_tet_p2_3")
    bgold = G.load(bcase["name"])
    bmesh = U.make_mesh(bcase, bgold)
    bspace = LagrangeFESpace(bmesh, bcase["p"])
    form = BilinearForm(bspace).add_integrator(ScalarDiffusionIntegrator())
    F = LinearForm(bspace).add_integrator(ScalarSourceIntegrator(U.torch_coef_func(C.source_cart))).assembly()
    op = DirichletBCOperator(form, gd=U.torch_coef_func(C.kappa_cart))
    uh = op.init_solution()
    Fb = op.apply(F, uh)
    assert np.max(np.abs(Fb.cpu().numpy() - bgold["F_bc"])) <= 1e-12 * np.max(np.abs(bgold["F_bc"]))
    x, info = cg(op, Fb, returninfo=True, atol=1e-14, rtol=1e-11)
    assert np.linalg.norm(x.cpu().numpy() - bgold["x"]) / np.linalg.norm(bgold["x"]) <= 1e-10


@pytest.mark.parametrize("case", C.TENSOR_BC_CASES, ids=lambda c: c["name"])
def test_elasticity_dirichlet_jacobi_cg(case, U):
    """TensorFunctionSpace.is_boundary_dof / boundary_interpolate + DirichletBC + Jacobi-preconditioned cg against the
    reference run (golden)"""
    from fealpy_b200.mesh import TriangleMesh, TetrahedronMesh
    from fealpy_b200.functionspace import LagrangeFESpace, TensorFunctionSpace
    from fealpy_b200.fem import BilinearForm, LinearElasticityIntegrator, DirichletBC
    from fealpy_b200.material import LinearElasticMaterial
    from fealpy_b200.sparse import CSRTensor
    from fealpy_b200.solver import cg
    from scipy.sparse import csr_matrix
    gold = G.load(case["name"])
    mesh = U.make_mesh(case, gold)
    GD = mesh.geo_dimension()
    space = TensorFunctionSpace(LagrangeFESpace(mesh, case["p"]), shape=(GD, -1) if case["dof_priority"] else (-1, GD))
    mat = LinearElasticMaterial("m", elastic_modulus=case["E"], poisson_ratio=case["nu"], hypo=case["hypo"])
    A = BilinearForm(space).add_integrator(LinearElasticityIntegrator(mat, q=case["q"])).assembly()
    U.assert_csr_matches(A, gold)
    thr = U.torch_coef_func(C.THRESHOLDS[case["threshold"]])
    bc = DirichletBC(space, gd=U.torch_coef_func(C.gd_vector), threshold=thr)
    assert np.array_equal(bc.is_boundary_dof.cpu().numpy(), gold["isbd"])
    uh, _ = space.boundary_interpolate(U.torch_coef_func(C.gd_vector), threshold=thr)
    np.testing.assert_allclose(uh.cpu().numpy(), gold["uh"], atol=1e-15)
    A2, F2 = bc.apply(A, U.t64(gold["F"]))
    assert np.max(np.abs(F2.cpu().numpy() - gold["F_bc"])) <= 1e-12 * np.max(np.abs(gold["F_bc"]))
    S = A2.to_scipy()
    R = csr_matrix((gold["Abc_data"], gold["Abc_indices"], gold["Abc_indptr"]), shape=S.shape)
    D = S - R
    assert (abs(D).max() if D.nnz else 0.0) <= 1e-12 * np.max(np.abs(gold["Abc_data"]))
    d = A2.diags()
    x, info = cg(A2, F2, M=CSRTensor(d.crow, d.col, 1.0 / d.values, A2.shape), returninfo=True, atol=1e-14, rtol=1e-11)
    assert np.linalg.norm(x.cpu().numpy() - gold["x"]) / np.linalg.norm(gold["x"]) <= 1e-10
    assert abs(info["niter"] - gold["info"]["niter"]) <= 2
    # vector load: LinearForm + VectorSourceIntegrator on the tensor space against the reference run
    from fealpy_b200.fem import LinearForm, VectorSourceIntegrator
    from fealpy_b200.basis import host_tables
    Fv = LinearForm(space).add_integrator(VectorSourceIntegrator(source=U.torch_coef_func(C.gd_vector), q=case["q"])).assembly()
    assert np.max(np.abs(Fv.cpu().numpy() - gold["F_vsrc"])) <= 1e-13 * np.max(np.abs(gold["F_vsrc"]))
    bcs = host_tables(mesh.TD, case["p"], case["p"] + 3 if case["q"] is None else case["q"])["bcs"]
    vals = U.t64(C.gd_vector(mesh.bc_to_point(bcs).cpu().numpy()))            # the same field as an (NC, NQ, GD) tensor
    Fv2 = LinearForm(space).add_integrator(VectorSourceIntegrator(source=vals, q=case["q"])).assembly()
    assert torch.equal(Fv, Fv2), "tensor-valued and callable sources give the same load"


@pytest.mark.parametrize("kind,dims,p,world", [("tet", (6, 5, 4), 2, 3), ("tet", (5, 4, 4), 1, 2), ("tri", (14, 11), 3, 5), ("tet", (4, 3, 3), 3, 2)])
def test_morton_partition_owned_rows_bit_identical_on_gpu(kind, dims, p, world, U):
    """general-mesh partition (parallel/mesh_partition.py) on a relabelled / shuffled / jittered mesh: one GPU plays every
    rank in turn; each rank's owned rows of ITS local assembly are bit-identical to the single-GPU matrix, every dof is owned
    exactly once (the 2-GPU NCCL solve of the same partition is tests/test_multi_gpu.py)"""
    from fealpy_b200.mesh import TriangleMesh, TetrahedronMesh
    from fealpy_b200.functionspace import LagrangeFESpace
    from fealpy_b200.parallel import verify_partition
    node, cell = _relabelled_mesh(kind, dims, seed=41 + p)
    mesh = (TriangleMesh if kind == "tri" else TetrahedronMesh)(U.t64(node), U.t64(cell))
    space = LagrangeFESpace(mesh, p)
    owned = 0
    for r in range(world):
        v = verify_partition(mesh, space, world, r, torch.device("cuda"), solve=False)
        assert v["owned_rows_bit_identical"], v
        owned += v["n_owned"]
    assert owned == space.number_of_global_dofs()


@pytest.mark.parametrize("dims", [(5, 4, 3), (3, 7, 2), (1, 1, 1), (6, 6, 6)])
def test_from_box_closed_form_numbering(dims, U, monkeypatch):
    """P2 numbering of a from_box tetrahedral mesh in closed form (fb2_tet_box_slab over the whole box: no edge sort on the
    assembly path) == the generic sorted-unique construction, which is pinned bit-exact against the reference's golden files"""
    from fealpy_b200.mesh import TetrahedronMesh
    m1 = TetrahedronMesh.from_box([0, 1, 0, 2, -1, 1], *dims)
    c_closed = m1.cell_to_ipoint(2)
    assert m1._edge is None, "no edge construction on the closed-form path"
    n_closed = m1.number_of_global_ipoints(2)
    monkeypatch.setenv("FB2_BOX_CLOSED_FORM", "0")
    m2 = TetrahedronMesh.from_box([0, 1, 0, 2, -1, 1], *dims)
    assert torch.equal(c_closed, m2.cell_to_ipoint(2)) and m2._edge is not None
    assert n_closed == m2.number_of_global_ipoints(2) and m1.number_of_edges() == m2.edge.shape[0]
