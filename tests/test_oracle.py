"""CPU tests: the oracle (oracle/fem_oracle.py) against (1) the reference's own golden
vectors (tests/golden/ref_testdata.npz) and (2) outputs of the real reference on the case
ladder (tests/golden/<case>.npz, made by tools/gen_golden.py)."""
import numpy as np
import pytest

import cases as C
import golden_util as G
from oracle import fem_oracle as O

RT = G.ref_testdata()


# ---------------------------------------------------------------- reference's own vectors
def test_ref_multi_index():
    assert np.array_equal(O.multi_index_matrix(2, 2), RT["multi_index_p2_d2"])
    assert np.array_equal(O.multi_index_matrix(2, 1), RT["multi_index_p2_d1"])


def test_ref_backend_primitives():
    m = O.Mesh(RT["bk_tri2d_node"], RT["bk_tri2d_cell"])
    p = int(RT["bk_tri2d_p"])
    np.testing.assert_allclose(m.cell_measure(), RT["bk_tri2d_simple_measure"], atol=1e-14)
    np.testing.assert_allclose(m.grad_lambda(), RT["bk_tri2d_triangle_grad_lambda_2d"], atol=1e-14)
    np.testing.assert_allclose(O.shape_function(RT["bk_tri2d_bcs"], p), RT["bk_tri2d_simple_shape_function"], atol=1e-14)
    np.testing.assert_allclose(O.grad_shape_function(RT["bk_tri2d_bcs"], p), RT["bk_tri2d_simple_grad_shape_function"], atol=1e-14)
    np.testing.assert_allclose(m.bc_to_point(RT["bk_tri2d_bcs"]), RT["bk_tri2d_bc_to_points"], atol=1e-14)
    t = O.Mesh(RT["bk_tet_node"], RT["bk_tet_cell"])
    np.testing.assert_allclose(t.grad_lambda(), RT["bk_tet_tetrahedron_grad_lambda_3d"], atol=1e-14)


def test_ref_element_matrices_p2_q3():
    # test/fem/test_scalar_diffusion_integrator.py:16-24, test_scalar_mass_integrator.py:16-24
    node, cell = O.tri_from_box([0, 1, 0, 1], 1, 1)
    m = O.Mesh(node, cell)
    np.testing.assert_array_almost_equal(O.diffusion_element(m, 2, q=3, coef=1), RT["diffusion_p2_q3_box1_Ke"])
    np.testing.assert_array_almost_equal(O.mass_element(m, 2, q=3, coef=1), RT["mass_p2_q3_box1_Ke"])


def test_ref_tet_mesh_vectors():
    # test/mesh/test_tetrahedron_mesh.py:62-162
    node, cell = O.tet_from_box([0, 1, 0, 1, 0, 1], 3, 2, 1)
    m = O.Mesh(node, cell)
    np.testing.assert_allclose(m.cell_measure(), RT["tet_321_cm"], atol=1e-14)
    np.testing.assert_allclose(m.grad_lambda(), RT["tet_321_glambda"], atol=1e-14)
    assert np.array_equal(m.cell_to_ipoint(4), RT["tet_321_cell2ipoint_p4"])
    np.testing.assert_allclose(m.interpolation_points(4), RT["tet_321_ipoints_p4"], atol=1e-14)
    # thresholded from_box: same node/cell construction, then cells with x,y,z<0.5 barycentre removed
    bc = node[cell].mean(axis=1)
    keep = ~((bc[:, 0] < 0.5) & (bc[:, 1] < 0.5) & (bc[:, 2] < 0.5))
    cell2 = cell[keep]
    valid = np.zeros(len(node), bool); valid[cell2] = True
    remap = np.zeros(len(node), np.int32); remap[valid] = np.arange(valid.sum(), dtype=np.int32)
    m2 = O.Mesh(node[valid], remap[cell2])
    assert np.array_equal(m2.cell, RT["tet_from_box_321_thr_cell"])
    assert np.array_equal(m2.edge, RT["tet_from_box_321_thr_edge"])
    assert np.array_equal(m2.face, RT["tet_from_box_321_thr_face"])
    np.testing.assert_allclose(m2.node, RT["tet_from_box_321_thr_node"], atol=1e-14)
    # grad_shape_function p=2 q=3 on from_box(1,1,1)
    node1, cell1 = O.tet_from_box([0, 1, 0, 1, 0, 1], 1, 1, 1)
    m1 = O.Mesh(node1, cell1)
    bcs, _ = O.quadrature(3, 3)
    np.testing.assert_allclose(O.grad_shape_function(bcs, 2), RT["tet_111_gsf_u_p2_q3"], atol=1e-14)
    got = np.transpose(O.grad_basis(m1, 2, bcs), (1, 0, 2, 3))      # reference layout is (NQ, NC, ldof, GD)
    np.testing.assert_allclose(got, RT["tet_111_gsf_x_p2_q3"], atol=1e-13)


def test_ref_tri_mesh_vectors():
    node, cell = O.tri_from_box([0, 1, 0, 1], 2, 2)
    assert np.array_equal(cell, RT["tri_22_cell"])
    m = O.Mesh(node, cell)
    bcs, _ = O.quadrature(2, 3)
    np.testing.assert_allclose(O.grad_basis(m, 2, bcs), RT["tri_22_gphi_p2_q3"], atol=1e-13)
    assert np.array_equal(m.cell_to_ipoint(4), RT["tri_22_cip_p4"])
    np.testing.assert_allclose(m.interpolation_points(4), RT["tri_22_ips_p4"], atol=1e-14)
    node2, cell2 = O.tri_from_box([-1, 1, -1, 1], 2, 2)
    np.testing.assert_allclose(O.Mesh(node2, cell2).grad_lambda(), RT["tri_glambda_m11_22"], atol=1e-14)


def test_ref_lagrange_space_vectors():
    node, cell = O.tri_from_box([0, 1, 0, 1], 1, 1)
    m = O.Mesh(node, cell)
    assert np.array_equal(m.cell_to_ipoint(2), RT["lfs_box1_p2_cell_to_dof"])
    assert np.array_equal(O.boundary_dof_flag(m, 2), RT["lfs_box1_p2_is_boundary_dof"])
    bcs = RT["lfs_box1_p2_bcs"]
    np.testing.assert_allclose(O.shape_function(bcs, 2)[None], RT["lfs_box1_p2_basis"], atol=1e-14)
    np.testing.assert_allclose(O.grad_basis(m, 2, bcs), RT["lfs_box1_p2_grad_basis"], atol=1e-13)
    np.testing.assert_allclose(m.interpolation_points(2), RT["lfs_box1_p2_ipoints"], atol=1e-14)


def test_ref_coalesce_and_tocsr():
    # test/sparse/test_coo_tensor.py:29-48
    r, c, v = O.coalesce(RT["coo_indices"][0], RT["coo_indices"][1], RT["coo_values"])
    assert np.array_equal(np.stack([r, c]), RT["coo_expected_indices"])
    assert np.array_equal(v, RT["coo_expected_values"])
    # test/sparse/test_coo_tensor.py:377-392
    D = RT["tocsr_dense"]
    rr, cc = np.nonzero(D)
    crow, col, val = O.tocsr(rr, cc, D[rr, cc], D.shape[0])
    assert np.array_equal(crow, RT["tocsr_crow"]) and np.array_equal(col, RT["tocsr_col"])
    assert np.array_equal(val, RT["tocsr_values"])


@pytest.mark.parametrize("k", [0, 1])
@pytest.mark.parametrize("p", [1, 2, 3, 4])
def test_ref_bilinear_form_matmul(k, p):
    # test/fem/test_bilinear_form.py:19-44: matrix-free (element-wise) product == assembled product
    m = O.Mesh(RT[f"bform_mesh{k}_node"], RT[f"bform_mesh{k}_cell"])
    Ke = O.diffusion_element(m, p)
    c2d = m.cell_to_ipoint(p)
    gdof = m.number_of_global_ipoints(p)
    x = np.random.default_rng(p).random(gdof)
    y = O.matfree_apply([(Ke, c2d)], gdof, x)
    crow, col, val = O.assemble([(Ke, c2d)], gdof)
    assert np.linalg.norm(y - O.csr_matvec(crow, col, val, x)) < 1e-12


# ---------------------------------------------------------------- outputs of the reference run
def build_oracle_case(case, gold):
    m = O.Mesh(gold["node"], gold["cell"])
    p = case["p"]
    c2d = m.cell_to_ipoint(p)
    sg = m.number_of_global_ipoints(p)
    tensor = case.get("tensor")
    if tensor is not None:
        c2d_use = O.tensor_cell_to_dof(c2d, sg, m.GD, tensor["dof_priority"])
        gdof = sg * m.GD
    else:
        c2d_use, gdof = c2d, sg
    groups, Kes, k = [], [], 0
    for grp in case["groups"]:
        acc = None
        for kind, spec in grp:
            q = spec.get("q")
            if kind == "elasticity":
                lam, mu = O.lame(spec["E"], spec["nu"])
                D = O.elastic_matrix(lam, mu, spec["hypo"], spec["E"], spec["nu"])
                Ke = O.elasticity_element(m, p, D, q=q, dof_priority=tensor["dof_priority"])
            else:
                NQ = len(O.quadrature(m.TD, p + 3 if q is None else q)[1])
                coef = G.case_coef(case, spec, k, gold, m.NC, NQ, m.GD)
                if kind == "diffusion":
                    Ke = O.diffusion_element(m, p, q=q, coef=coef, method=spec.get("method"))
                else:
                    Ke = O.mass_element(m, p, q=q, coef=coef)
            Kes.append(Ke)
            acc = Ke if acc is None else acc + Ke
            k += 1
        groups.append((acc, c2d_use))
    return m, c2d, c2d_use, gdof, groups, Kes


@pytest.mark.parametrize("case", [c for c in C.CASES if not c.get("values_only_checksum")], ids=lambda c: c["name"])
def test_oracle_matrix_free_matches_reference_run(case):
    """BilinearForm.__matmul__ before assembly as the real reference computed it (fem/bilinear_form.py:126-158)"""
    gold = G.load(case["name"])
    m, c2d, c2d_use, gdof, groups, Kes = build_oracle_case(case, gold)
    w = O.matfree_apply(groups, gdof, gold["matfree_u"])
    assert G.rel_err(w, gold["matfree_w"]) < 1e-13


@pytest.mark.parametrize("case", C.CASES, ids=[c["name"] for c in C.CASES])
def test_oracle_matches_reference_run(case):
    gold = G.load(case["name"])
    m, c2d, c2d_use, gdof, groups, Kes = build_oracle_case(case, gold)
    assert np.array_equal(c2d, gold["cell2dof_scalar"])
    assert np.array_equal(c2d_use, gold["cell2dof"])
    assert np.array_equal(m.edge, gold["edge"])
    assert gdof == gold["info"]["gdof"]
    if case.get("elem"):
        for k, Ke in enumerate(Kes):
            assert G.rel_err(Ke, gold[f"Ke_{k}"]) < 1e-13
    crow, col, val = O.assemble(groups, gdof)
    assert crow.dtype == np.int64 and col.dtype == gold["col"].dtype
    assert np.array_equal(crow, gold["crow"]) and np.array_equal(col, gold["col"])
    assert len(val) == gold["info"]["nnz"]
    if "values" in gold:
        assert G.rel_err(val, gold["values"]) < 1e-13
    else:
        assert G.rel_err(val[::97], gold["values_sample"]) < 1e-13
        assert abs(val.sum() - gold["values_sum"][0]) < 1e-10 * gold["values_sum"][1]
    if case.get("cg"):
        mv = lambda v: O.csr_matvec(crow, col, val, v)
        b = mv(np.ones(gdof))
        # b = A.1 cancels the O(1) diffusion entries: scale the error by max|A|
        assert np.max(np.abs(b - gold["b"])) < 1e-13 * np.max(np.abs(val))
        x, info = O.cg(mv, b)
        assert abs(info["niter"] - gold["info"]["niter"]) <= 1
        assert np.linalg.norm(x - gold["x"]) / np.linalg.norm(gold["x"]) < 1e-10


@pytest.mark.parametrize("case", C.BC_CASES, ids=[c["name"] for c in C.BC_CASES])
def test_oracle_bc_matches_reference_run(case):
    gold = G.load(case["name"])
    m = O.Mesh(gold["node"], gold["cell"])
    p = case["p"]
    F = O.source_vector(m, p, C.source_cart)
    assert G.rel_err(F, gold["F"]) < 1e-13
    thr = C.THRESHOLDS[case["threshold"]] if case.get("threshold") else None
    isbd = O.boundary_dof_flag(m, p, thr, case.get("method"))
    assert np.array_equal(isbd, gold["isbd"])
    # the values DirichletBC writes are selected with method='interp' whatever the flag method (lagrange_fe_space.py:131)
    isbd_val = O.boundary_dof_flag(m, p, thr, "interp")
    if thr is not None:
        assert np.array_equal(isbd_val, gold["isbd_val"])
        assert np.array_equal(isbd_val, isbd) == (case.get("method") == "interp")      # centroid and interp selections differ here
    ip = m.interpolation_points(p)
    np.testing.assert_allclose(ip, gold["ipoints"], atol=1e-14)
    uh = np.zeros(len(F)); uh[isbd_val] = C.kappa_cart(ip[isbd_val])
    A, F2 = O.dirichlet_apply(gold["crow"], gold["col"], gold["values"], F, uh, isbd)
    assert G.rel_err(F2, gold["F_bc"]) < 1e-13
    A.eliminate_zeros(); A.sort_indices()
    assert np.array_equal(A.indptr, gold["Abc_indptr"]) and np.array_equal(A.indices, gold["Abc_indices"])
    assert G.rel_err(A.data, gold["Abc_data"]) < 1e-13
    x, info = O.cg(lambda v: A @ v, F2, atol=1e-14, rtol=1e-11)
    assert abs(info["niter"] - gold["info"]["niter"]) <= 2
    assert np.linalg.norm(x - gold["x"]) / np.linalg.norm(gold["x"]) < 1e-10


BCG_CASES = [c for c in C.CASES if c.get("cg") and "bcg_B" in G.load(c["name"])]


@pytest.mark.parametrize("case", BCG_CASES, ids=lambda c: c["name"])
def test_oracle_batched_cg_matches_reference_run(case):
    """cg with a (dof, batch) right-hand side as the real reference solved it (solver/cg.py:58-121)"""
    gold = G.load(case["name"])
    crow, col, val = gold["crow"], gold["col"], gold["values"]
    mv = lambda v: np.stack([O.csr_matvec(crow, col, val, v[:, k]) for k in range(v.shape[1])], axis=1)
    x, info = O.cg(mv, gold["bcg_B"], atol=1e-14, rtol=1e-12)
    assert abs(info["niter"] - gold["info"]["bcg_niter"]) <= 1
    assert np.linalg.norm(x - gold["bcg_x"]) / np.linalg.norm(gold["bcg_x"]) < 1e-10


@pytest.mark.parametrize("case", C.BC_CASES, ids=[c["name"] for c in C.BC_CASES])
def test_oracle_dirichlet_operator_matches_reference_run(case):
    """the matrix-free constrained operator (fem/dirichlet_bc_operator.py:13-67) restated on the oracle's matrix-free
    product, against the reference's own DirichletBCOperator outputs on the unassembled form"""
    gold = G.load(case["name"])
    m = O.Mesh(gold["node"], gold["cell"])
    p = case["p"]
    c2d = m.cell_to_ipoint(p)
    gdof = m.number_of_global_ipoints(p)
    ke = O.diffusion_element(m, p)
    if case.get("reaction"):
        ke = ke + O.mass_element(m, p)
    groups = [(ke, c2d)]
    thr = C.THRESHOLDS[case["threshold"]] if case.get("threshold") else None
    bd = O.boundary_dof_flag(m, p, thr, None)                 # the operator's constructor has no `method`
    assert np.array_equal(bd, gold["op_isbd"])
    ip = m.interpolation_points(p)
    uh = np.zeros(gdof)
    uh[bd] = C.kappa_cart(ip[bd])                              # init_solution(): boundary_interpolate with the flag as threshold
    np.testing.assert_allclose(uh, gold["op_uh"], atol=1e-15)
    F = O.source_vector(m, p, C.source_cart)
    Fop = F - O.matfree_apply(groups, gdof, uh)
    Fop[bd] = uh[bd]
    assert G.rel_err(Fop, gold["op_F"]) < 1e-13
    v = gold["op_u"].copy()
    val = v[bd].copy()
    v[bd] = 0.0
    w = O.matfree_apply(groups, gdof, v)
    w[bd] = val
    assert G.rel_err(w, gold["op_w"]) < 1e-13


@pytest.mark.parametrize("case", C.TENSOR_BC_CASES, ids=[c["name"] for c in C.TENSOR_BC_CASES])
def test_oracle_tensor_bc_matches_reference_run(case):
    """elasticity + TensorFunctionSpace Dirichlet data + Jacobi CG against the reference run"""
    gold = G.load(case["name"])
    m = O.Mesh(gold["node"], gold["cell"])
    p, GD, prio = case["p"], m.GD, case["dof_priority"]
    sg = m.number_of_global_ipoints(p)
    gdof = GD * sg
    lam, mu = O.lame(case["E"], case["nu"])
    D = O.elastic_matrix(lam, mu, case["hypo"], case["E"], case["nu"])
    c2d = O.tensor_cell_to_dof(m.cell_to_ipoint(p), sg, GD, prio)
    crow, col, val = O.assemble([(O.elasticity_element(m, p, D, q=case["q"], dof_priority=prio), c2d)], gdof)
    assert np.array_equal(crow, gold["crow"]) and np.array_equal(col, gold["col"]) and G.rel_err(val, gold["values"]) < 1e-13
    thr = C.THRESHOLDS[case["threshold"]]
    isbd = O.tensor_boundary_dof_flag(m, p, GD, prio, thr, None)
    assert np.array_equal(isbd, gold["isbd"])
    sflag = O.boundary_dof_flag(m, p, thr, None)
    ip = m.interpolation_points(p)
    uh = np.zeros(gdof)
    v = C.gd_vector(ip[sflag])
    if prio:
        uh.reshape(GD, sg)[:, sflag] = v.T
    else:
        uh.reshape(sg, GD)[sflag, :] = v
    np.testing.assert_allclose(uh, gold["uh"], atol=1e-15)
    A, F2 = O.dirichlet_apply(crow, col, val, gold["F"], uh, isbd)
    assert G.rel_err(F2, gold["F_bc"]) < 1e-13
    diag = A.diagonal()
    x, info = O.cg(lambda w: A @ w, F2, Minv=lambda r: r / diag, atol=1e-14, rtol=1e-11)
    assert abs(info["niter"] - gold["info"]["niter"]) <= 2
    assert np.linalg.norm(x - gold["x"]) / np.linalg.norm(gold["x"]) < 1e-10
    # vector load of the reference's LinearForm + VectorSourceIntegrator on the tensor space
    Fv = O.vector_source_vector(m, p, C.gd_vector, GD, prio, q=case["q"])
    assert G.rel_err(Fv, gold["F_vsrc"]) < 1e-13
