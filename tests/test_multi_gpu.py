"""2-GPU tests (NCCL / NVLink peer memory): slab-partitioned assembly + distributed CG against the single-GPU path, through
fealpy_b200.parallel.verify_slab -- the same check `bench.py --gpus N` runs before timing.  Skipped unless two CUDA
devices are visible (run with `gpurun --gpus 2`)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, p, mode, out):
    import torch.distributed as dist
    from fealpy_b200.parallel import verify_slab, make_dist_solver
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        out[rank] = verify_slab(world, rank, dev, p=p, dims_per_rank=(12, 8, 8),
                                solver_factory=lambda A, part, group=None: make_dist_solver(A, part, mode=mode))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["nccl", "peer"])
@pytest.mark.parametrize("p", [1, 2])
def test_two_gpu_slab_matches_single_gpu(p, mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    if mode == "peer":
        pytest.importorskip("fealpy_b200.parallel.peer_cg")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), p, mode, out), nprocs=2, join=True)
    for r in range(2):
        v = out[r]
        assert v["owned_rows_bit_identical"], v
        assert v["niter_ok"] and v["x_rel"] <= 1e-10, v
        assert v["mode"] == mode


def _worker_morton(rank, world, port, kind, dims, p, out):
    import sys
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_gpu_parity import _relabelled_mesh
    from fealpy_b200.mesh import TriangleMesh, TetrahedronMesh
    from fealpy_b200.functionspace import LagrangeFESpace
    from fealpy_b200.parallel import verify_partition
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        node, cell = _relabelled_mesh(kind, dims, seed=23)
        mesh = (TriangleMesh if kind == "tri" else TetrahedronMesh)(torch.as_tensor(node, device=dev), torch.as_tensor(cell, device=dev))
        out[rank] = verify_partition(mesh, LagrangeFESpace(mesh, p), world, rank, dev)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind,dims,p", [("tet", (8, 6, 5), 2), ("tri", (24, 19), 3)])
def test_two_gpu_morton_partition_matches_single_gpu(kind, dims, p):
    """a relabelled / shuffled mesh split in Morton order over 2 GPUs: owned rows bit-identical to the single-GPU matrix,
    NCCL halo exchange with packed sends, distributed CG equal to fb2_cg"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_morton, args=(2, _free_port(), kind, dims, p, out), nprocs=2, join=True)
    for r in range(2):
        v = out[r]
        assert v["owned_rows_bit_identical"], v
        assert v["niter_ok"] and v["x_rel"] <= 1e-10, v


@pytest.mark.parametrize("mode", ["nccl", "peer"])
def test_four_gpu_slab_matches_single_gpu(mode):
    """four x-slabs: the interior ranks have two neighbours (both halo directions live)"""
    if torch.cuda.device_count() < 4:
        pytest.skip("needs 4 GPUs")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(4, _free_port(), 2, mode, out), nprocs=4, join=True)
    for r in range(4):
        v = out[r]
        assert v["owned_rows_bit_identical"], v
        assert v["niter_ok"] and v["x_rel"] <= 1e-10, v
        assert v["mode"] == mode


@pytest.mark.parametrize("kind,dims,p", [("tet", (8, 7, 6), 2), ("tri", (30, 23), 3)])
def test_four_gpu_morton_partition_matches_single_gpu(kind, dims, p):
    """Morton partition over 4 GPUs: ranks with up to three neighbours, packed sends to each (NCCL)"""
    if torch.cuda.device_count() < 4:
        pytest.skip("needs 4 GPUs")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_morton, args=(4, _free_port(), kind, dims, p, out), nprocs=4, join=True)
    for r in range(4):
        v = out[r]
        assert v["owned_rows_bit_identical"], v
        assert v["niter_ok"] and v["x_rel"] <= 1e-10, v
