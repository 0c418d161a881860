"""Cost of the general-mesh (Morton) partition at size: one GPU plays rank 0 of `world`.
   python tools/gpu_time_partition.py [n] [world]"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from fealpy_b200.mesh import TetrahedronMesh
from fealpy_b200.functionspace import LagrangeFESpace
from fealpy_b200.fem import BilinearForm, ScalarDiffusionIntegrator, ScalarMassIntegrator
from fealpy_b200.parallel.mesh_partition import PartitionedProblem

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
mesh = TetrahedronMesh.from_box([0, 1, 0, 1, 0, 1], n, n, n)
space = LagrangeFESpace(mesh, 2)
c2d = space.cell_to_dof()
torch.cuda.synchronize()
for rank in (0, world // 2):
    t0 = time.perf_counter()
    pp = PartitionedProblem(mesh, space, world, rank)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    bf = BilinearForm(pp.space).add_integrator(ScalarDiffusionIntegrator()).add_integrator(ScalarMassIntegrator())
    A = bf.assembly()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    A = bf.assembly()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    print(f"n {n} world {world} rank {rank}: partition {1e3 * (t1 - t0):.1f} ms, cells {pp.part.cells.numel()} of {mesh.number_of_cells()}, "
          f"owned dofs {pp.part.n_owned}, local {pp.part.n_local}, neighbours {[e.peer for e in pp.part.exchanges]}, "
          f"cold assembly {1e3 * (t2 - t1):.1f} ms, warm {1e3 * (t3 - t2):.2f} ms, peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
