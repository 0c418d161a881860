"""The reference-side binding (fealpy_b200.integration) against the REAL FEALPy, when it is present
(build container only; skipped on the GPU box where /root/reference does not exist)."""
import os
import sys

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")


@pytest.fixture(scope="module")
def fealpy_torch():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import ref_import
    ref_import.install()
    from fealpy.backend import backend_manager as bm
    bm.set_backend("pytorch")
    yield bm
    bm.set_backend("numpy")


def test_install_registers_variants_and_refuses_cpu(fealpy_torch):
    import fealpy_b200.integration as b200
    from fealpy.mesh import TriangleMesh
    from fealpy.functionspace import LagrangeFESpace
    from fealpy.fem import BilinearForm, ScalarDiffusionIntegrator
    import fealpy.solver as solver
    b200.install()
    b200.install()                                       # idempotent
    assert "b200" in ScalarDiffusionIntegrator.assembly.virtual_table
    assert getattr(BilinearForm, "_b200_installed", False)
    mesh = TriangleMesh.from_box([0, 1, 0, 1], 2, 2)     # torch CPU tensors
    space = LagrangeFESpace(mesh, 2)
    # the adapter reads the reference objects (node, cell, p, cell_to_dof) and then insists on CUDA
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            b200.adapt_space(space)
        I = ScalarDiffusionIntegrator(method="b200")
        with pytest.raises(RuntimeError):
            I.assembly(space)
    # features outside the accelerated path fall back to the reference implementation, CPU inputs do not
    bform = BilinearForm(space).add_integrator(ScalarDiffusionIntegrator())
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            bform.assembly()
    assert callable(solver.cg)
