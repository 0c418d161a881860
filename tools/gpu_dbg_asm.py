import sys, os, time
sys.path.insert(0, os.getcwd())
import torch
from fealpy_b200.mesh import TetrahedronMesh
from fealpy_b200.functionspace import LagrangeFESpace
from fealpy_b200.fem import BilinearForm, ScalarDiffusionIntegrator, ScalarMassIntegrator
from fealpy_b200.solver import cg
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    mesh = TetrahedronMesh.from_box([0, 1, 0, 1, 0, 1], 128, 128, 128)
    space = LagrangeFESpace(mesh, 2)
    bf = BilinearForm(space); bf.add_integrator(ScalarDiffusionIntegrator()); bf.add_integrator(ScalarMassIntegrator())
    A = bf.assembly()
    b = A @ torch.ones(A.shape[0], dtype=torch.float64, device="cuda")
    ev = lambda: torch.cuda.Event(enable_timing=True)
    for k in range(8):
        ms0 = torch.cuda.memory_stats()
        s0, s1, s2 = ev(), ev(), ev()
        t0 = time.perf_counter()
        s0.record()
        A_ = bf.assembly()
        t1 = time.perf_counter()
        s1.record()
        x, info = cg(A_, b, atol=0.0, rtol=0.0, maxit=iters, returninfo=True)
        s2.record()
        torch.cuda.synchronize()
        ms1 = torch.cuda.memory_stats()
        print(f"step {k}: asm {s0.elapsed_time(s1):7.3f} ms (host {1e3*(t1-t0):6.2f} ms)  cg {s1.elapsed_time(s2):8.2f} ms  "
              f"cudaMalloc +{ms1['num_device_alloc']-ms0['num_device_alloc']} cudaFree +{ms1['num_device_free']-ms0['num_device_free']} "
              f"retries +{ms1['num_alloc_retries']-ms0['num_alloc_retries']} reserved {ms1['reserved_bytes.all.current']/2**30:.1f} GiB")
