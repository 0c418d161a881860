"""One-off timings of the non-headline BASELINE configs (1, 3, 4) on one GPU: assembly (warm) and CG.
Not a bench line -- recorded in profiles/ for DESIGN.md."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fealpy_b200.mesh import TriangleMesh, TetrahedronMesh
from fealpy_b200.functionspace import LagrangeFESpace, TensorFunctionSpace
from fealpy_b200.fem import BilinearForm, ScalarDiffusionIntegrator, ScalarMassIntegrator, LinearElasticityIntegrator
from fealpy_b200.material import LinearElasticMaterial
from fealpy_b200.decorator import cartesian
from fealpy_b200.solver import cg


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


rows = []
with torch.cuda.stream(torch.cuda.Stream()):
    # config 1
    mesh = TriangleMesh.from_box([0, 1, 0, 1], 1024, 1024)
    space = LagrangeFESpace(mesh, 1)
    bf = BilinearForm(space); bf.add_integrator(ScalarDiffusionIntegrator(q=3)); bf.add_integrator(ScalarMassIntegrator(q=3))
    t, A = timeit(bf.assembly)
    b = A @ torch.ones(A.shape[0], dtype=torch.float64, device="cuda")
    tc, (x, info) = timeit(lambda: cg(A, b, returninfo=True), reps=2, warm=1)
    rows.append(dict(config="1 tri P1 1024^2 diff+mass q=3", nnz=A.nnz, asm_ms=t, nnz_per_s=A.nnz / t * 1e3, path=bf.last_path,
                     cg_niter=info["niter"], cg_ms=tc, cg_it_per_s=info["niter"] / tc * 1e3))
    del A, b, bf, space, mesh
    # config 3
    mesh = TriangleMesh.from_box([0, 1, 0, 1], 1024, 1024)
    space = LagrangeFESpace(mesh, 3)

    @cartesian
    def kappa(p):
        return 1.0 + 0.5 * torch.sin(2 * torch.pi * p[..., 0]) * torch.cos(2 * torch.pi * p[..., 1])
    bf = BilinearForm(space); bf.add_integrator(ScalarDiffusionIntegrator(coef=kappa, q=6))
    t, A = timeit(bf.assembly, reps=3, warm=1)
    rows.append(dict(config="3 tri P3 1024^2 var-coef diffusion q=6", nnz=A.nnz, asm_ms=t, nnz_per_s=A.nnz / t * 1e3, path=bf.last_path))
    del A, bf, space, mesh
    # config 4
    mesh = TetrahedronMesh.from_box([0, 1, 0, 1, 0, 1], 128, 128, 128)
    space = TensorFunctionSpace(LagrangeFESpace(mesh, 1), shape=(-1, 3))
    mat = LinearElasticMaterial("m", elastic_modulus=1.0, poisson_ratio=0.3, hypo="3D")
    bf = BilinearForm(space); bf.add_integrator(LinearElasticityIntegrator(mat, q=4))
    t, A = timeit(bf.assembly, reps=3, warm=1)
    rows.append(dict(config="4 tet P1x3 elasticity 128^3 q=4", nnz=A.nnz, asm_ms=t, nnz_per_s=A.nnz / t * 1e3, path=bf.last_path))
for r in rows:
    print(json.dumps(r))
