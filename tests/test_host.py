"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol (no
compute calls), the host tables agree with the oracle, and the product refuses to run without CUDA."""
import os
import re
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "fealpy_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fb2_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from fealpy_b200 import _lib
    lib = _lib.load()                        # builds must exist in-tree; raises otherwise
    names = header_symbols()
    assert len(names) >= 35
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (fb2_\w+)", out))
    assert set(names) <= exported, sorted(set(names) - exported)
    assert set(names) == set(_lib.SIGNATURES), (set(names) ^ set(_lib.SIGNATURES))
    for n in names:
        assert hasattr(lib, n)
    assert lib.fb2_version() >= 100
    assert lib.fb2_last_error() is not None


def test_size_queries_run_without_gpu():
    from fealpy_b200 import _lib
    lib = _lib.load()
    assert lib.fb2_sort_workspace_bytes(1000) > 12000
    assert lib.fb2_coo_workspace_bytes(1000) > lib.fb2_sort_workspace_bytes(1000)
    assert lib.fb2_cg_workspace_bytes(1000, 9000) >= 3 * 8000 and lib.fb2_spmv_plan_blocks(9000, 2048) == 5
    assert lib.fb2_partial_workspace_bytes() >= 4096 * 8
    assert lib.fb2_entity_workspace_bytes(100, 6) > 0 and lib.fb2_sym_workspace_bytes(100, 10, 500) > 0


def test_no_cpu_fallback():
    from fealpy_b200.mesh import TriangleMesh
    from fealpy_b200.solver import cg
    from fealpy_b200.sparse import CSRTensor
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError):
        TriangleMesh.from_box([0, 1, 0, 1], 2, 2)
    with pytest.raises(RuntimeError):
        TriangleMesh(torch.zeros((3, 2), dtype=torch.float64), torch.tensor([[0, 1, 2]], dtype=torch.int32))
    A = CSRTensor(torch.tensor([0, 1, 2]), torch.tensor([0, 1], dtype=torch.int32), torch.ones(2, dtype=torch.float64), (2, 2))
    with pytest.raises(RuntimeError):
        cg(A, torch.ones(2, dtype=torch.float64))


def test_product_does_not_import_oracle_or_reference():
    pkg = os.path.join(ROOT, "fealpy_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                s = open(os.path.join(dp, f)).read()
                assert "import oracle" not in s and "from oracle" not in s, f
                assert "/root/reference" not in s and "import fealpy\n" not in s and "from fealpy " not in s, f


@pytest.mark.parametrize("TD,p,q", [(2, 1, 3), (2, 2, 5), (2, 3, 6), (3, 1, 4), (3, 2, 5), (3, 3, 6)])
def test_host_tables_reproduce_oracle_element_matrices(TD, p, q):
    """the pre-contracted tables the kernels consume, contracted on the host exactly like
    elem.cu / assemble.cu do, reproduce the oracle's element matrices"""
    from fealpy_b200 import basis as B
    from oracle import fem_oracle as O
    bcs, ws = O.quadrature(TD, q)
    t = B.host_tables(TD, p, q)
    assert np.array_equal(t["ws"], ws) and np.array_equal(t["bcs"], bcs)
    assert np.array_equal(B.multi_index_matrix(p, TD), O.multi_index_matrix(p, TD))
    assert np.array_equal(t["phi"], O.shape_function(bcs, p)) and np.array_equal(t["R"], O.grad_shape_function(bcs, p))
    if TD == 2:
        node, cell = O.tri_from_box([0, 1, 0, 1], 3, 2)
    else:
        node, cell = O.tet_from_box([0, 1, 0, 1, 0, 1], 2, 2, 1)
    rng = np.random.default_rng(1)
    node = node + 0.05 * rng.uniform(-1, 1, node.shape)
    m = O.Mesh(node, cell)
    D, cm = m.grad_lambda(), m.cell_measure()
    NV = TD + 1
    G = np.stack([np.einsum("cm,cm->c", D[:, k], D[:, l]) * cm for k in range(NV) for l in range(k, NV)], axis=1)
    K = np.einsum("ijt,ct->cij", t["Ms"], G)
    ref = O.diffusion_element(m, p, q=q)
    assert np.max(np.abs(K - ref)) <= 1e-13 * np.max(np.abs(ref))
    Km = cm[:, None, None] * t["Mm"][None]
    refm = O.mass_element(m, p, q=q)
    assert np.max(np.abs(Km - refm)) <= 1e-13 * np.max(np.abs(refm))
    # elasticity blocks from M4 (interleaved layout)
    lam, mu = O.lame(1.0, 0.3)
    Dm = O.elastic_matrix(lam, mu, "3D" if TD == 3 else "plane_strain")
    A = np.einsum("ijkl,cka,clb,c->cijab", t["M4"], D, D, cm)
    ref_e = O.elasticity_element(m, p, Dm, q=q, dof_priority=False)
    L = t["phi"].shape[1]
    KK = np.zeros_like(ref_e)
    dd = (2 * mu + lam)
    for a in range(TD):
        for b in range(TD):
            if a == b:
                blk = dd * A[..., a, a] + mu * sum(A[..., z, z] for z in range(TD) if z != a)
            else:
                blk = lam * A[..., a, b] + mu * A[..., b, a]
            KK[:, a::TD, b::TD] = blk
    assert np.max(np.abs(KK - ref_e)) <= 1e-12 * np.max(np.abs(ref_e))


def test_material_matches_oracle():
    from fealpy_b200.material import LinearElasticMaterial
    from oracle import fem_oracle as O
    for hypo in ("3D", "plane_strain", "plane_stress"):
        m = LinearElasticMaterial("m", elastic_modulus=2.0, poisson_ratio=0.25, hypo=hypo)
        lam, mu = O.lame(2.0, 0.25)
        assert (m.lam, m.mu) == (lam, mu)
        assert np.array_equal(m.elastic_matrix()[0, 0].numpy(), O.elastic_matrix(lam, mu, hypo, 2.0, 0.25))
    with pytest.raises(ValueError):
        LinearElasticMaterial("m", elastic_modulus=1.0)
