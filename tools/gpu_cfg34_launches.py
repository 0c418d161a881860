import sys, os, json
sys.path.insert(0, os.getcwd())
import torch
from fealpy_b200.mesh import TetrahedronMesh, TriangleMesh
from fealpy_b200.functionspace import LagrangeFESpace, TensorFunctionSpace
from fealpy_b200.fem import BilinearForm, LinearElasticityIntegrator, ScalarDiffusionIntegrator
from fealpy_b200.material import LinearElasticMaterial
from fealpy_b200.decorator import cartesian
with torch.cuda.stream(torch.cuda.Stream()):
    mesh = TetrahedronMesh.from_box([0, 1, 0, 1, 0, 1], 128, 128, 128)
    space = TensorFunctionSpace(LagrangeFESpace(mesh, 1), shape=(-1, 3))
    mat = LinearElasticMaterial("m", elastic_modulus=1.0, poisson_ratio=0.3, hypo="3D")
    bf = BilinearForm(space); bf.add_integrator(LinearElasticityIntegrator(mat, q=4))
    for _ in range(3): A = bf.assembly()
    torch.cuda.synchronize()
    del A, bf, space, mesh
    mesh = TriangleMesh.from_box([0, 1, 0, 1], 1024, 1024)
    space = LagrangeFESpace(mesh, 3)
    @cartesian
    def kappa(p):
        return 1.0 + 0.5 * torch.sin(2 * torch.pi * p[..., 0]) * torch.cos(2 * torch.pi * p[..., 1])
    bf = BilinearForm(space); bf.add_integrator(ScalarDiffusionIntegrator(coef=kappa, q=6))
    for _ in range(3): A = bf.assembly()
    torch.cuda.synchronize()
