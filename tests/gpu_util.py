"""builders shared by the GPU parity tests: product objects for a case of tests/cases.py"""
import numpy as np
import torch

import cases as C
import golden_util as G


def t64(a, dev="cuda"):
    return torch.as_tensor(np.ascontiguousarray(a), device=dev)


def make_mesh(case, gold, from_box=False):
    from fealpy_b200.mesh import TriangleMesh, TetrahedronMesh
    cls = TriangleMesh if case["mesh"] == "tri" else TetrahedronMesh
    if from_box:
        return cls.from_box(C.box_of(case), *case["dims"])
    return cls(t64(gold["node"]), t64(gold["cell"].astype(np.int32)))


def torch_coef_func(f):
    """wrap a numpy coefficient function of tests/cases.py for torch points (cartesian)"""
    def g(p):
        return torch.as_tensor(f(p.cpu().numpy()), device=p.device)
    g.coordtype = "cartesian"
    return g


def make_space(case, mesh):
    from fealpy_b200.functionspace import LagrangeFESpace, TensorFunctionSpace
    sspace = LagrangeFESpace(mesh, case["p"])
    tensor = case.get("tensor")
    if tensor is None:
        return sspace, sspace
    GD = mesh.geo_dimension()
    shape = (GD, -1) if tensor["dof_priority"] else (-1, GD)
    return sspace, TensorFunctionSpace(sspace, shape=shape)


def make_integrators(case, gold, mesh):
    """-> list of groups, each a list of integrator objects (same order as the golden Ke_k)"""
    from fealpy_b200.fem import ScalarDiffusionIntegrator, ScalarMassIntegrator, LinearElasticityIntegrator
    from fealpy_b200.material import LinearElasticMaterial
    groups, k = [], 0
    for grp in case["groups"]:
        ints = []
        for kind, spec in grp:
            q = spec.get("q")
            if kind == "elasticity":
                mat = LinearElasticMaterial("m", elastic_modulus=spec["E"], poisson_ratio=spec["nu"], hypo=spec["hypo"])
                ints.append(LinearElasticityIntegrator(mat, q=q))
            else:
                coef = spec.get("coef")
                if isinstance(coef, str):
                    coef = torch_coef_func(C.COEF_FUNCS[coef]) if coef in C.COEF_FUNCS else t64(gold[f"coef_{k}"])
                if kind == "diffusion":
                    ints.append(ScalarDiffusionIntegrator(coef=coef, q=q, method=spec.get("method")))
                else:
                    ints.append(ScalarMassIntegrator(coef=coef, q=q))
            k += 1
        groups.append(ints)
    return groups


def make_form(case, gold, path="auto", mesh=None):
    from fealpy_b200.fem import BilinearForm
    mesh = make_mesh(case, gold) if mesh is None else mesh
    sspace, space = make_space(case, mesh)
    bform = BilinearForm(space, assembly_path=path)
    groups = make_integrators(case, gold, mesh)
    for ints in groups:
        bform.add_integrator(*ints)
    return mesh, space, bform, groups


def assert_csr_matches(A, gold, tol=1e-12):
    crow, col, val = A.crow.cpu().numpy(), A.col.cpu().numpy(), A.values.cpu().numpy()
    assert crow.dtype == np.int64 and col.dtype == np.int32 and val.dtype == np.float64
    assert np.array_equal(crow, gold["crow"]), "indptr differs"
    assert np.array_equal(col, gold["col"]), "indices differ"
    if "values" in gold:
        ref = gold["values"]
        scale = np.max(np.abs(ref))
        assert np.max(np.abs(val - ref)) <= tol * scale, f"values differ: {np.max(np.abs(val - ref)) / scale:.3e}"
        big = np.abs(ref) > 1e-6 * scale           # entry-wise relative check away from cancellation
        assert np.max(np.abs(val[big] - ref[big]) / np.abs(ref[big])) <= 1e-10
    else:
        ref = gold["values_sample"]
        scale = np.max(np.abs(ref))
        assert np.max(np.abs(val[::97] - ref)) <= tol * scale
        assert abs(val.sum() - gold["values_sum"][0]) <= 1e-10 * gold["values_sum"][1]
