"""world_size-2 (gloo, CPU) tests of the multi-GPU host logic: slab ownership / window numbering
/ halo slices (fealpy_b200.parallel.box_partition) and the distributed CG driver
(fealpy_b200.parallel.dist_cg) with a numpy backend standing in for the CUDA kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import fem_oracle as O
from fealpy_b200.parallel.box_partition import BoxSlab, box_edges_before, box_number_of_edges


@pytest.mark.parametrize("dims", [(3, 2, 4), (2, 5, 3), (4, 4, 4)])
def test_closed_form_edge_numbering(dims):
    nx, ny, nz = dims
    node, cell = O.tet_from_box([0, 1, 0, 1, 0, 1], *dims)
    m = O.Mesh(node, cell)
    assert m.NE == box_number_of_edges(*dims)
    es = np.sort(m.edge, axis=1)
    for i in range(nx + 1):
        for j in range(ny + 1):
            for k in range(nz + 1):
                nid = i * (ny + 1) * (nz + 1) + j * (nz + 1) + k
                assert np.searchsorted(es[:, 0], nid) == box_edges_before(nx, ny, nz, i, j, k)


@pytest.mark.parametrize("p", [1, 2])
@pytest.mark.parametrize("world", [1, 2, 3])
def test_slab_partition_covers_and_matches_global_matrix(p, world):
    import dist_util as D
    nx, ny, nz = 6, 3, 2
    parts = [BoxSlab(nx, ny, nz, p, world, r) for r in range(world)]
    assert sum(s.n_owned for s in parts) == parts[0].gdof
    gm, gc2d, _, _ = D.oracle_slab_problem(parts[0], p)
    gcrow, gcol, gval = O.assemble([(O.diffusion_element(gm, p), gc2d), (O.mass_element(gm, p), gc2d)], parts[0].gdof)
    seen = np.zeros(parts[0].gdof, dtype=int)
    for s in parts:
        _, _, (crow, col, val), _ = D.oracle_slab_problem(s, p)
        l2g = s.local_to_global(np.arange(s.n_local))
        assert np.all(np.diff(l2g) > 0), "window numbering must be monotone in the global numbering"
        for lo, hi in (s.own_nodes, s.own_edges):
            for l in range(lo, hi):
                g = l2g[l]
                seen[g] += 1
                a, b = crow[l], crow[l + 1]
                ga, gb = gcrow[g], gcrow[g + 1]
                assert np.array_equal(l2g[col[a:b]], gcol[ga:gb]), "owned row pattern == global row pattern"
                assert np.array_equal(val[a:b], gval[ga:gb]), "owned row values are bit-identical"
        # halo slices pair up with the neighbour's send slices (same global ids)
        for ex in s.exchanges:
            peer = parts[ex.peer]
            back = [e for e in peer.exchanges if e.peer == s.rank][0]
            for (rl, rh), (sl, sh) in zip(ex.recv, back.send):
                assert np.array_equal(l2g[rl:rh], peer.local_to_global(np.arange(sl, sh)))
    assert np.all(seen == 1), "every global row is owned exactly once"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, p, dims, out):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    import dist_util as D
    from fealpy_b200.parallel.dist_cg import dist_cg
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        part = BoxSlab(*dims, p, world, rank)
        gm, gc2d, (crow, col, val), _ = D.oracle_slab_problem(part, p)
        gcrow, gcol, gval = O.assemble([(O.diffusion_element(gm, p), gc2d), (O.mass_element(gm, p), gc2d)], part.gdof)
        gb = O.csr_matvec(gcrow, gcol, gval, np.ones(part.gdof))
        l2g = part.local_to_global(np.arange(part.n_local))
        b = torch.from_numpy(gb[l2g].copy())
        ops = D.NumpyCgOps(crow, col, val, part.own_ranges)
        x, info = dist_cg(ops, b, torch.zeros_like(b), part.exchanges, check_every=4)
        xo, oinfo = O.cg(lambda v: O.csr_matvec(gcrow, gcol, gval, v), gb)
        own = np.zeros(part.n_local, dtype=bool)
        own[part.own_nodes[0]:part.own_nodes[1]] = True
        own[part.own_edges[0]:part.own_edges[1]] = True
        err = np.linalg.norm(x.numpy()[own] - xo[l2g[own]]) / np.linalg.norm(xo)
        out[rank] = (float(err), info["niter"], oinfo["niter"])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,dims,p", [(2, (6, 3, 3), 1), (2, (6, 3, 3), 2), (4, (9, 3, 2), 2)])
def test_distributed_cg_gloo(world, dims, p):
    """the distributed driver over gloo (interior ranks have two neighbours at world 4)"""
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, p, dims, out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        err, niter, oniter = out[r]
        assert err < 1e-10, (r, err)
        assert abs(niter - oniter) <= 1, (niter, oniter)
    assert len({out[r][1] for r in range(world)}) == 1, "all ranks stop at the same iteration"


@pytest.mark.parametrize("p", [1, 2])
@pytest.mark.parametrize("nx,world", [(16, 8), (13, 5), (16, 4), (19, 8), (128, 8)])
def test_slab_partition_structure_many_ranks(nx, world, p):
    """ownership and halo plan for the rank counts of the scaling run (structure only, no matrices): every global dof is
    owned exactly once, window numbering is monotone in the global numbering, and every receive slice names exactly
    the global ids of the peer's matching send slice"""
    ny, nz = (3, 2) if nx < 100 else (4, 4)
    parts = [BoxSlab(nx, ny, nz, p, world, r) for r in range(world)]
    gdof = parts[0].gdof
    assert sum(s.n_owned for s in parts) == gdof
    seen = np.zeros(gdof, dtype=np.int64)
    for s in parts:
        l2g = s.local_to_global(np.arange(s.n_local))
        assert l2g.min() >= 0 and l2g.max() < gdof and np.all(np.diff(l2g) > 0)
        for lo, hi in (s.own_nodes, s.own_edges):
            seen[l2g[lo:hi]] += 1
        assert len(s.exchanges) == (s.rank > 0) + (s.rank < world - 1)
        for ex in s.exchanges:
            assert abs(ex.peer - s.rank) == 1
            peer = parts[ex.peer]
            back = [e for e in peer.exchanges if e.peer == s.rank]
            assert len(back) == 1
            assert len(ex.recv) == len(back[0].send) and len(ex.send) == len(back[0].recv)
            for (rl, rh), (sl, sh) in zip(ex.recv, back[0].send):
                assert rh - rl == sh - sl
                assert np.array_equal(l2g[rl:rh], peer.local_to_global(np.arange(sl, sh)))
                # what a rank receives it does not own; what it sends it owns
                own = np.zeros(s.n_local, dtype=bool)
                own[s.own_nodes[0]:s.own_nodes[1]] = True
                own[s.own_edges[0]:s.own_edges[1]] = True
                assert not own[rl:rh].any()
            pown = np.zeros(peer.n_local, dtype=bool)
            pown[peer.own_nodes[0]:peer.own_nodes[1]] = True
            pown[peer.own_edges[0]:peer.own_edges[1]] = True
            for sl, sh in back[0].send:
                assert pown[sl:sh].all()
    assert np.all(seen == 1)


def test_slab_partition_rejects_too_many_ranks():
    with pytest.raises(ValueError):
        BoxSlab(8, 3, 2, 2, 8, 0)          # 9 node planes cannot give every rank two
