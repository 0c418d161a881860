"""CG on the UNASSEMBLED form (matrix-free product inside the operator CG) against CG on the assembled matrix, config 2.
   python tools/gpu_time_matfree_cg.py [n] [iters]"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from fealpy_b200.mesh import TetrahedronMesh
from fealpy_b200.functionspace import LagrangeFESpace
from fealpy_b200.fem import BilinearForm, ScalarDiffusionIntegrator, ScalarMassIntegrator
from fealpy_b200.solver import cg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 104
torch.cuda.synchronize()
t0 = time.perf_counter()
mesh = TetrahedronMesh.from_box([0, 1, 0, 1, 0, 1], n, n, n)
space = LagrangeFESpace(mesh, 2)
form = BilinearForm(space).add_integrator(ScalarDiffusionIntegrator()).add_integrator(ScalarMassIntegrator())
gdof = space.number_of_global_dofs()
b = form @ torch.ones(gdof, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
t_setup = time.perf_counter() - t0
ts = []
for k in range(4):
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    x, info = cg(form, b, atol=0.0, rtol=0.0, maxit=iters, returninfo=True)
    s1.record()
    torch.cuda.synchronize()
    ts.append(s0.elapsed_time(s1) / info["niter"])
print(f"n {n} matrix-free CG: {min(ts[1:]):.4f} ms/it = {1e3 / min(ts[1:]):.1f} it/s (niter {info['niter']}), set-up (mesh + adjacency + rhs) {1e3 * t_setup:.1f} ms, "
      f"memory {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB", flush=True)
torch.cuda.reset_peak_memory_stats()
t0 = time.perf_counter()
form2 = BilinearForm(space).add_integrator(ScalarDiffusionIntegrator()).add_integrator(ScalarMassIntegrator())
A = form2.assembly()
torch.cuda.synchronize()
t_asm = time.perf_counter() - t0
ts = []
for k in range(4):
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    y, info2 = cg(A, b, atol=0.0, rtol=0.0, maxit=iters, returninfo=True)
    s1.record()
    torch.cuda.synchronize()
    ts.append(s0.elapsed_time(s1) / info2["niter"])
print(f"n {n} assembled CG:   {min(ts[1:]):.4f} ms/it = {1e3 / min(ts[1:]):.1f} it/s (niter {info2['niter']}), symbolic + assembly {1e3 * t_asm:.1f} ms, "
      f"memory {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB;  |x_free - x_asm| / |x| = {float((x - y).norm() / y.norm()):.2e}", flush=True)
