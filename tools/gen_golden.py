"""Generate tests/golden/*.npz by running the REAL reference (numpy backend) on the case
ladder of tests/cases.py, and at the same time pin the oracle (oracle/fem_oracle.py)
against it.  Build-container only (needs /root/reference); the fixtures are committed.

    python tools/gen_golden.py            # all cases
"""
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, ROOT)
import ref_import  # noqa: E402
ref_import.install()

from fealpy.backend import backend_manager as bm  # noqa: E402
from fealpy.mesh import TriangleMesh, TetrahedronMesh  # noqa: E402
from fealpy.functionspace import LagrangeFESpace, TensorFunctionSpace  # noqa: E402
from fealpy.fem import (BilinearForm, LinearForm, ScalarDiffusionIntegrator, ScalarMassIntegrator,  # noqa: E402
                        LinearElasticityIntegrator, ScalarSourceIntegrator, VectorSourceIntegrator, DirichletBC)
from fealpy.material.elastic_material import LinearElasticMaterial  # noqa: E402
from fealpy.solver import cg  # noqa: E402
from fealpy.decorator import cartesian  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases as C  # noqa: E402
from oracle import fem_oracle as O  # noqa: E402

bm.set_backend("numpy")
GOLD = os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)


def ref_mesh(case):
    box = C.box_of(case)
    if case["mesh"] == "tri":
        m = TriangleMesh.from_box(box, *case["dims"])
        cls = TriangleMesh
    else:
        m = TetrahedronMesh.from_box(box, *case["dims"])
        cls = TetrahedronMesh
    if case.get("jitter") is not None:
        node = C.perturb(np.asarray(m.node), case["dims"], case["jitter"])
        m = cls(node, np.asarray(m.cell))
    return m


def rel_err(a, b):
    scale = np.max(np.abs(b)) if b.size else 1.0
    return float(np.max(np.abs(a - b)) / scale) if b.size else 0.0


def resolve_coef(spec, mesh_o, q, p):
    """-> (reference-side coef, oracle-side coef, array-to-store or None)"""
    coef = spec.get("coef")
    if coef is None or isinstance(coef, (int, float)):
        return coef, coef, None
    if coef in C.COEF_FUNCS:
        f = C.COEF_FUNCS[coef]
        return cartesian(lambda pts, _f=f: _f(pts)), f, None
    TD = mesh_o.TD
    qq = p + 3 if q is None else q
    NQ = len(O.quadrature(TD, qq)[1])
    arr = C.coef_array(coef, mesh_o.NC, NQ, mesh_o.GD, seed=zlib.crc32(coef.encode()) % 1000 + mesh_o.NC)      # (str hash() is salted per process: not reproducible)
    return arr, arr, arr


def run_case(case):
    name = case["name"]
    mesh = ref_mesh(case)
    p = case["p"]
    sspace = LagrangeFESpace(mesh, p)
    mesh_o = O.Mesh(np.asarray(mesh.node), np.asarray(mesh.cell))
    GD = mesh_o.GD
    out = dict(node=np.asarray(mesh.node), cell=np.asarray(mesh.cell))

    # ---- numbering
    c2d_ref = np.asarray(sspace.cell_to_dof())
    c2d_o = mesh_o.cell_to_ipoint(p)
    assert np.array_equal(c2d_ref, c2d_o), f"{name}: oracle cell_to_dof differs"
    assert c2d_ref.dtype == c2d_o.dtype, (c2d_ref.dtype, c2d_o.dtype)
    assert np.array_equal(np.asarray(mesh.edge), mesh_o.edge), f"{name}: edge differs"
    sgdof = sspace.number_of_global_dofs()
    assert sgdof == mesh_o.number_of_global_ipoints(p)
    out["cell2dof_scalar"] = c2d_ref
    out["edge"] = np.asarray(mesh.edge)

    space = sspace
    tensor = case.get("tensor")
    if tensor is not None:
        shape = (GD, -1) if tensor["dof_priority"] else (-1, GD)
        space = TensorFunctionSpace(sspace, shape=shape)
        c2d_t = np.asarray(space.cell_to_dof())
        c2d_to = O.tensor_cell_to_dof(c2d_o, sgdof, GD, tensor["dof_priority"])
        assert np.array_equal(c2d_t, c2d_to), f"{name}: tensor cell_to_dof differs"
        out["cell2dof"] = c2d_t
        c2d_use = c2d_to
    else:
        out["cell2dof"] = c2d_ref
        c2d_use = c2d_o
    gdof = space.number_of_global_dofs()

    bform = BilinearForm(space)
    groups_o = []
    k = 0
    for grp in case["groups"]:
        ints, Ke_o_sum = [], None
        for kind, spec in grp:
            q = spec.get("q")
            if kind == "elasticity":
                mat = LinearElasticMaterial("m", elastic_modulus=spec["E"], poisson_ratio=spec["nu"], hypo=spec["hypo"])
                I = LinearElasticityIntegrator(mat, q=q)
                lam, mu = O.lame(spec["E"], spec["nu"])
                D = O.elastic_matrix(lam, mu, spec["hypo"], spec["E"], spec["nu"])
                assert np.array_equal(D, np.asarray(mat.elastic_matrix())[0, 0]), f"{name}: D differs"
                Ke_o = O.elasticity_element(mesh_o, p, D, q=q, dof_priority=tensor["dof_priority"])
            else:
                cref, corc, carr = resolve_coef(spec, mesh_o, q, p)
                if carr is not None:
                    out[f"coef_{k}"] = carr
                if kind == "diffusion":
                    I = ScalarDiffusionIntegrator(coef=cref, q=q, method=spec.get("method"))
                    Ke_o = O.diffusion_element(mesh_o, p, q=q, coef=corc, method=spec.get("method"))
                else:
                    I = ScalarMassIntegrator(coef=cref, q=q)
                    Ke_o = O.mass_element(mesh_o, p, q=q, coef=corc)
            Ke_ref = np.asarray(I.assembly(space))
            e = rel_err(Ke_o, Ke_ref)
            assert e < 1e-13, f"{name}: oracle K_e[{k}] differs rel {e}"
            if case.get("elem"):
                out[f"Ke_{k}"] = Ke_ref
            ints.append(I)
            Ke_o_sum = Ke_o if Ke_o_sum is None else Ke_o_sum + Ke_o
            k += 1
        bform.add_integrator(*ints)
        groups_o.append((Ke_o_sum, c2d_use))

    # matrix-free product of the UNASSEMBLED form (BilinearForm.__matmul__, fem/bilinear_form.py:126-158): row f4's pin
    if not case.get("values_only_checksum"):
        um = np.random.default_rng(1000 + gdof).standard_normal(gdof)
        assert bform._M is None
        wm = np.asarray(bform @ um)
        wo = O.matfree_apply(groups_o, gdof, um)
        assert rel_err(wo, wm) < 1e-13, f"{name}: oracle matrix-free product differs"
        out.update(matfree_u=um, matfree_w=wm)
    A = bform.assembly()
    crow, col, val = np.asarray(A.crow), np.asarray(A.col), np.asarray(A.values)
    ocrow, ocol, oval = O.assemble(groups_o, gdof)
    assert crow.dtype == ocrow.dtype == np.int64 and col.dtype == ocol.dtype, (crow.dtype, col.dtype, ocol.dtype)
    assert np.array_equal(crow, ocrow) and np.array_equal(col, ocol), f"{name}: oracle CSR pattern differs"
    e = rel_err(oval, val)
    assert e < 1e-13, f"{name}: oracle CSR values differ rel {e}"
    out.update(crow=crow, col=col)
    if case.get("values_only_checksum"):
        # large case: keep pattern + a strided sample of values + checksums
        out["values_sample"] = val[::97].copy()
        out["values_sum"] = np.array([val.sum(), np.abs(val).sum()])
    else:
        out["values"] = val
    info = dict(gdof=int(gdof), nnz=int(A.nnz))

    if case.get("cg"):
        b = A @ np.ones(gdof)
        x, cinfo = cg(A, b, returninfo=True)
        xo, oinfo = O.cg(lambda v: O.csr_matvec(ocrow, ocol, oval, v), O.csr_matvec(ocrow, ocol, oval, np.ones(gdof)))
        assert abs(oinfo["niter"] - cinfo["niter"]) <= 1, (name, oinfo, cinfo)
        ex = np.linalg.norm(xo - x) / np.linalg.norm(x)
        assert ex < 1e-10, f"{name}: oracle CG solution differs {ex}"
        out.update(b=b, x=np.asarray(x))
        info.update(niter=int(cinfo["niter"]), residual=float(cinfo["residual"]))
        if gdof <= 1000:
            # batched right-hand sides (solver/cg.py:58-121: per-column alpha / beta, joint stopping test) through the reference
            Bm = np.random.default_rng(3000 + gdof).standard_normal((gdof, 3))
            xb, binfo = cg(A, Bm, returninfo=True, atol=1e-14, rtol=1e-12)
            xbo, boinfo = O.cg(lambda v: np.stack([O.csr_matvec(ocrow, ocol, oval, v[:, k]) for k in range(v.shape[1])], axis=1), Bm,
                               atol=1e-14, rtol=1e-12)
            assert abs(boinfo["niter"] - binfo["niter"]) <= 1 and np.linalg.norm(xbo - xb) / np.linalg.norm(xb) < 1e-10, name
            out.update(bcg_B=Bm, bcg_x=np.asarray(xb))
            info.update(bcg_niter=int(binfo["niter"]), bcg_residual=float(binfo["residual"]))
    out["info"] = np.array(json.dumps(info))
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(f"{name:36s} gdof {gdof:7d} nnz {A.nnz:8d} " + (f"cg {info.get('niter')}" if case.get('cg') else ""))


def run_bc_case(case):
    """Poisson with source + Dirichlet data (rows f1/f2): reference LinearForm / DirichletBC / cg.  `threshold` / `method`
    restrict the Dirichlet part (functionspace/dofs.py:23-55); `reaction` adds a mass term so that the problem with a
    partly natural boundary stays well conditioned."""
    name = case["name"]
    mesh = ref_mesh(case)
    p = case["p"]
    space = LagrangeFESpace(mesh, p)
    mesh_o = O.Mesh(np.asarray(mesh.node), np.asarray(mesh.cell))
    gdof = space.number_of_global_dofs()
    f = cartesian(lambda pts: C.source_cart(pts))
    g = cartesian(lambda pts: C.kappa_cart(pts))
    thr = C.THRESHOLDS[case["threshold"]] if case.get("threshold") else None
    method = case.get("method")
    bform = BilinearForm(space)
    bform.add_integrator(ScalarDiffusionIntegrator())
    if case.get("reaction"):
        bform.add_integrator(ScalarMassIntegrator())
    A = bform.assembly()
    lform = LinearForm(space)
    lform.add_integrator(ScalarSourceIntegrator(f))
    F = np.asarray(lform.assembly())
    Fo = O.source_vector(mesh_o, p, C.source_cart)
    assert rel_err(Fo, F) < 1e-13, f"{name}: oracle source vector differs"
    isbd = np.asarray(space.is_boundary_dof(threshold=thr, method=method))
    isbd_o = O.boundary_dof_flag(mesh_o, p, thr, method)
    assert np.array_equal(isbd, isbd_o), f"{name}: boundary flags differ"
    if thr is not None:      # the two methods must really differ on this case, or it pins nothing
        other = O.boundary_dof_flag(mesh_o, p, thr, "interp" if method in (None, "centroid") else None)
        assert not np.array_equal(other, isbd_o), f"{name}: centroid and interp flags coincide"
    # the values DirichletBC writes: boundary_interpolate always selects with method='interp' (lagrange_fe_space.py:131)
    isbd_val = O.boundary_dof_flag(mesh_o, p, thr, "interp")
    ip = np.asarray(space.interpolation_points())
    ipo = mesh_o.interpolation_points(p)
    assert np.max(np.abs(ip - ipo)) < 1e-14, f"{name}: interpolation points differ"
    A2, F2 = DirichletBC(space, gd=g, threshold=thr, method=method).apply(A, F)
    uh = np.zeros(gdof)
    uh[isbd_val] = C.kappa_cart(ipo[isbd_val])
    Ao, F2o = O.dirichlet_apply(np.asarray(A.crow), np.asarray(A.col), np.asarray(A.values), Fo, uh, isbd_o)
    assert rel_err(F2o, np.asarray(F2)) < 1e-13, f"{name}: BC rhs differs"
    x, cinfo = cg(A2, F2, returninfo=True, atol=1e-14, rtol=1e-11)     # tight: compare solutions, not stopping luck
    A2s = A2.to_scipy().tocsr().copy()          # to_scipy shares buffers with A2
    A2s.sum_duplicates(); A2s.sort_indices()
    d = (A2s - Ao)
    assert (abs(d).max() if d.nnz else 0.0) < 1e-13, f"{name}: BC matrix differs"
    A2s.eliminate_zeros()
    # the matrix-free constrained operator on the UNASSEMBLED form (fem/dirichlet_bc_operator.py:13-67); its constructor has
    # no `method`: the Dirichlet set is is_boundary_dof(threshold) with the default (face-centroid) selection
    from fealpy.fem.dirichlet_bc_operator import DirichletBCOperator
    bform_free = BilinearForm(space)
    bform_free.add_integrator(ScalarDiffusionIntegrator())
    if case.get("reaction"):
        bform_free.add_integrator(ScalarMassIntegrator())
    op = DirichletBCOperator(bform_free, gd=g, threshold=thr)
    op_isbd = np.asarray(op.is_boundary_dof)
    assert np.array_equal(op_isbd, O.boundary_dof_flag(mesh_o, p, thr, None)), f"{name}: operator flags differ"
    op_uh = np.asarray(op.init_solution())
    op_F = np.asarray(op.apply(F, op_uh))
    op_u = np.random.default_rng(2000 + gdof).standard_normal(gdof)
    op_w = np.asarray(op @ op_u)
    assert bform_free._M is None
    out = dict(node=np.asarray(mesh.node), cell=np.asarray(mesh.cell), cell2dof=np.asarray(space.cell_to_dof()),
               crow=np.asarray(A.crow), col=np.asarray(A.col), values=np.asarray(A.values),
               op_isbd=op_isbd, op_uh=op_uh, op_F=op_F, op_u=op_u, op_w=op_w,
               F=F, isbd=isbd, isbd_val=isbd_val, ipoints=ip, F_bc=np.asarray(F2),
               Abc_indptr=A2s.indptr.astype(np.int64), Abc_indices=A2s.indices.astype(np.int32), Abc_data=A2s.data,
               x=np.asarray(x),
               info=np.array(json.dumps(dict(gdof=int(gdof), niter=int(cinfo["niter"]), residual=float(cinfo["residual"])))))
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(f"{name:36s} gdof {gdof:7d} cg {cinfo['niter']}  bd {int(isbd.sum())}")


def run_tensor_bc_case(case):
    """linear elasticity on a TensorFunctionSpace with a displaced part of the boundary: reference
    TensorFunctionSpace.is_boundary_dof / boundary_interpolate (functionspace/tensor_space.py:159-260), DirichletBC, and
    cg with the reference's Jacobi preconditioner CSRTensor(diags, 1/diag) (solver/iterative_solver_manger.py:273-280)"""
    from fealpy.sparse import CSRTensor
    name = case["name"]
    mesh = ref_mesh(case)
    p = case["p"]
    sspace = LagrangeFESpace(mesh, p)
    mesh_o = O.Mesh(np.asarray(mesh.node), np.asarray(mesh.cell))
    GD = mesh_o.GD
    prio = case["dof_priority"]
    space = TensorFunctionSpace(sspace, shape=(GD, -1) if prio else (-1, GD))
    gdof = space.number_of_global_dofs()
    sg = sspace.number_of_global_dofs()
    thr = C.THRESHOLDS[case["threshold"]]
    mat = LinearElasticMaterial("m", elastic_modulus=case["E"], poisson_ratio=case["nu"], hypo=case["hypo"])
    A = BilinearForm(space).add_integrator(LinearElasticityIntegrator(mat, q=case["q"])).assembly()
    lam, mu = O.lame(case["E"], case["nu"])
    D = O.elastic_matrix(lam, mu, case["hypo"], case["E"], case["nu"])
    c2d_t = O.tensor_cell_to_dof(mesh_o.cell_to_ipoint(p), sg, GD, prio)
    ocrow, ocol, oval = O.assemble([(O.elasticity_element(mesh_o, p, D, q=case["q"], dof_priority=prio), c2d_t)], gdof)
    assert np.array_equal(np.asarray(A.crow), ocrow) and np.array_equal(np.asarray(A.col), ocol)
    assert rel_err(oval, np.asarray(A.values)) < 1e-13
    F = A @ np.ones(gdof)
    gd = cartesian(lambda pts: C.gd_vector(pts))
    bc = DirichletBC(space, gd=gd, threshold=thr)                      # method=None -> face-centroid selection
    isbd = np.asarray(bc.is_boundary_dof)
    isbd_o = O.tensor_boundary_dof_flag(mesh_o, p, GD, prio, thr, None)
    assert np.array_equal(isbd, isbd_o), f"{name}: tensor boundary flags differ"
    A2, F2 = bc.apply(A, F)
    sflag = O.boundary_dof_flag(mesh_o, p, thr, None)
    ipo = mesh_o.interpolation_points(p)
    uh = np.zeros(gdof)
    val = C.gd_vector(ipo[sflag])
    if prio:
        uh.reshape(GD, sg)[:, sflag] = val.T
    else:
        uh.reshape(sg, GD)[sflag, :] = val
    Ao, F2o = O.dirichlet_apply(ocrow, ocol, np.asarray(A.values), np.asarray(F), uh, isbd_o)
    assert rel_err(F2o, np.asarray(F2)) < 1e-13, f"{name}: tensor BC rhs differs"
    dg = A2.diags()
    M = CSRTensor(dg.crow, dg.col, 1.0 / dg.values, A2.shape)
    x, cinfo = cg(A2, F2, M=M, returninfo=True, atol=1e-14, rtol=1e-11)
    A2s = A2.to_scipy().tocsr().copy()
    A2s.sum_duplicates(); A2s.sort_indices()
    dd = (A2s - Ao)
    assert (abs(dd).max() if dd.nnz else 0.0) < 1e-13, f"{name}: tensor BC matrix differs"
    A2s.eliminate_zeros()
    # vector load: reference LinearForm + VectorSourceIntegrator on the tensor space (fem/vector_source_integrator.py)
    fsrc = cartesian(lambda pts: C.gd_vector(pts))
    Fv = np.asarray(LinearForm(space).add_integrator(VectorSourceIntegrator(source=fsrc, q=case["q"])).assembly())
    Fvo = O.vector_source_vector(mesh_o, p, C.gd_vector, GD, prio, q=case["q"])
    assert rel_err(Fvo, Fv) < 1e-13, f"{name}: oracle vector source differs"
    out = dict(node=np.asarray(mesh.node), cell=np.asarray(mesh.cell), crow=np.asarray(A.crow), col=np.asarray(A.col),
               values=np.asarray(A.values), F=np.asarray(F), isbd=isbd, F_bc=np.asarray(F2), uh=uh, F_vsrc=Fv,
               Abc_indptr=A2s.indptr.astype(np.int64), Abc_indices=A2s.indices.astype(np.int32), Abc_data=A2s.data,
               x=np.asarray(x),
               info=np.array(json.dumps(dict(gdof=int(gdof), niter=int(cinfo["niter"]), residual=float(cinfo["residual"])))))
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(f"{name:36s} gdof {gdof:7d} cg {cinfo['niter']}  bd {int(isbd.sum())}")


if __name__ == "__main__":
    only = sys.argv[1:]
    for case in C.CASES:
        if not only or case["name"] in only:
            run_case(case)
    for case in C.BC_CASES:
        if not only or case["name"] in only:
            run_bc_case(case)
    for case in C.TENSOR_BC_CASES:
        if not only or case["name"] in only:
            run_tensor_bc_case(case)
    total = sum(os.path.getsize(os.path.join(GOLD, f)) for f in os.listdir(GOLD))
    print("golden dir bytes:", total)
