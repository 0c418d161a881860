"""Set-up cost of a Dirichlet problem at config-2 size: boundary flags, boundary interpolation of a callable g, apply().
   python tools/gpu_time_bc.py [n]"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from fealpy_b200.mesh import TetrahedronMesh
from fealpy_b200.functionspace import LagrangeFESpace
from fealpy_b200.fem import BilinearForm, ScalarDiffusionIntegrator, DirichletBC, LinearForm, ScalarSourceIntegrator

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
mesh = TetrahedronMesh.from_box([0, 1, 0, 1, 0, 1], n, n, n)
space = LagrangeFESpace(mesh, 2)
A = BilinearForm(space).add_integrator(ScalarDiffusionIntegrator()).assembly()
def g(p): return torch.sin(p[..., 0]) * torch.cos(p[..., 1]) + p[..., 2]
g.coordtype = "cartesian"
def f(p): return 1.0 + p[..., 0] * p[..., 1]
f.coordtype = "cartesian"
for rep in range(2):
    log = []
    def phase(name, fn):
        torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize()
        log.append(f"{name} {1e3 * (time.perf_counter() - t0):.1f}"); return out
    F = phase("LinearForm(source).assembly", lambda: LinearForm(space).add_integrator(ScalarSourceIntegrator(f)).assembly())
    bc = phase("DirichletBC()", lambda: DirichletBC(space, gd=g))
    A2, F2 = phase("bc.apply", lambda: bc.apply(A, F))
    print(f"n {n} pass {rep}: " + " | ".join(log) + f" | boundary dofs {int(bc.is_boundary_dof.sum())} of {A.shape[0]}", flush=True)
