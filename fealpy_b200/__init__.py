"""fealpy_b200 -- B200-native (sm_100a) Lagrange-FEM global assembly + CG behind the FEALPy API.

    from fealpy_b200.mesh import TetrahedronMesh
    from fealpy_b200.functionspace import LagrangeFESpace
    from fealpy_b200.fem import BilinearForm, ScalarDiffusionIntegrator, ScalarMassIntegrator
    from fealpy_b200.solver import cg

Hand-written CUDA kernels behind a C ABI (include/fealpy_b200.h); PyTorch tensors are the only
device containers.  No CPU fallback: without the built library or a CUDA device, calls raise.
"""
__version__ = "0.1.0"
