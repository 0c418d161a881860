"""2-GPU test (NCCL): slab-partitioned assembly + distributed CG against the single-GPU solve.
Skipped unless two CUDA devices are visible (run with `gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, p, dims, out):
    import torch.distributed as dist
    from fealpy_b200.fem import BilinearForm, ScalarDiffusionIntegrator, ScalarMassIntegrator
    from fealpy_b200.parallel import SlabProblem, CudaCgOps, dist_cg
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        sp = SlabProblem([0, 1, 0, 1, 0, 1], *dims, p, world, rank, device=f"cuda:{rank}")
        bf = BilinearForm(sp.space)
        bf.add_integrator(ScalarDiffusionIntegrator())
        bf.add_integrator(ScalarMassIntegrator())
        A = bf.assembly()
        part = sp.part
        # b = A_global @ 1: owned rows of the window matrix are complete rows
        b = A @ torch.ones(A.shape[0], dtype=torch.float64, device=A.device)
        x, info = dist_cg(CudaCgOps(A, part.own_ranges), b, torch.zeros_like(b), part.exchanges)
        own = torch.zeros(part.n_local, dtype=torch.bool, device=A.device)
        own[part.own_nodes[0]:part.own_nodes[1]] = True
        own[part.own_edges[0]:part.own_edges[1]] = True
        out[rank] = (float((x[own] - 1.0).abs().max()), info["niter"], float(info["residual"]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("p", [1, 2])
def test_two_gpu_cg_matches_single_gpu(p):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from fealpy_b200.mesh import TetrahedronMesh
    from fealpy_b200.functionspace import LagrangeFESpace
    from fealpy_b200.fem import BilinearForm, ScalarDiffusionIntegrator, ScalarMassIntegrator
    from fealpy_b200.solver import cg
    dims = (24, 12, 12)
    mesh = TetrahedronMesh.from_box([0, 1, 0, 1, 0, 1], *dims)
    bf = BilinearForm(LagrangeFESpace(mesh, p))
    bf.add_integrator(ScalarDiffusionIntegrator())
    bf.add_integrator(ScalarMassIntegrator())
    A = bf.assembly()
    b = A @ torch.ones(A.shape[0], dtype=torch.float64, device="cuda")
    x, info = cg(A, b, returninfo=True)
    ref_err, ref_it = float((x - 1.0).abs().max()), info["niter"]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), p, dims, out), nprocs=2, join=True)
    for r in range(2):
        err, niter, resid = out[r]
        assert abs(niter - ref_it) <= 1, (niter, ref_it)
        assert err <= max(10 * ref_err, 1e-8)
