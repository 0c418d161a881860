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


# ---------------------------------------------------------------------------------------------------------------
# general meshes: Morton partition (fealpy_b200.parallel.mesh_partition) -- relabelled / shuffled meshes, p = 1..3
# ---------------------------------------------------------------------------------------------------------------
def _shuffled_mesh(kind, dims, seed):
    """a from_box mesh with randomly relabelled nodes, shuffled cells and jittered geometry: nothing of the box numbering is left"""
    import cases as C
    rng = np.random.default_rng(seed)
    if kind == "tet":
        node, cell = O.tet_from_box([0, 1, 0, 1, 0, 1], *dims)
    else:
        node, cell = O.tri_from_box([0, 1, 0, 1], *dims)
    node = C.perturb(node, dims, seed)
    perm = rng.permutation(len(node))
    inv = np.empty_like(perm); inv[perm] = np.arange(len(perm))
    node2 = node[perm]
    cell2 = inv[cell][rng.permutation(len(cell))].astype(np.int32)
    return node2, cell2


def _oracle_local(part, node, cell, gc2d, p, Ke=None):
    """what rank `part.rank` assembles: its local cells (same vertices, same order) with local dof ids.  The element
    matrices are taken from ONE global evaluation: numpy's einsum(optimize=True) may pick another contraction order for
    another cell count, which changes last bits (the CUDA kernels compute a cell's matrix independently of the others)."""
    lnode, lcell = part.local_mesh(torch.from_numpy(node), torch.from_numpy(cell))
    assert np.array_equal(lnode.numpy()[lcell.numpy()], node[cell[part.cells.numpy()]]), "local mesh = the same cells, same vertex order"
    c2d = part.local_cell2dof(torch.from_numpy(gc2d)).numpy()
    if Ke is None:
        gm = O.Mesh(node, cell)
        Ke = (O.diffusion_element(gm, p), O.mass_element(gm, p))
    cells = part.cells.numpy()
    # local dof ids come from the GLOBAL numbering (the local mesh would number its own edges differently)
    return O.assemble([(Ke[0][cells], c2d), (Ke[1][cells], c2d)], part.n_local)


@pytest.mark.parametrize("kind,dims,p", [("tet", (4, 3, 3), 1), ("tet", (4, 3, 2), 2), ("tri", (7, 6), 3), ("tet", (3, 2, 2), 3)])
@pytest.mark.parametrize("world", [2, 3, 5])
def test_morton_partition_matches_global_matrix(kind, dims, p, world):
    """every dof is owned exactly once; every rank's owned rows -- columns mapped back to global ids -- hold the same
    pattern and BIT-identical values as the single-process matrix; receive ranges pair up with the peers' packed sends"""
    from fealpy_b200.parallel.mesh_partition import MeshPartition
    node, cell = _shuffled_mesh(kind, dims, seed=31 + p)
    gm = O.Mesh(node, cell)
    gc2d = gm.cell_to_ipoint(p)
    gdof = gm.number_of_global_ipoints(p)
    Ke = (O.diffusion_element(gm, p), O.mass_element(gm, p))
    gcrow, gcol, gval = O.assemble([(Ke[0], gc2d), (Ke[1], gc2d)], gdof)
    parts = [MeshPartition(torch.from_numpy(node), torch.from_numpy(cell), torch.from_numpy(gc2d), gdof, world, r) for r in range(world)]
    seen = np.zeros(gdof, dtype=int)
    for s in parts:
        l2g = s.l2g.numpy()
        assert np.unique(l2g).size == s.n_local
        assert np.all(np.diff(l2g[:s.n_owned]) > 0), "owned dofs ascend in the global numbering"
        seen[l2g[:s.n_owned]] += 1
        assert np.all(np.diff(s.cells.numpy()) > 0), "local cells keep their global order (summation order of the owned rows)"
        crow, col, val = _oracle_local(s, node, cell, gc2d, p, Ke)
        for l in range(s.n_owned):
            g = l2g[l]
            a, b = crow[l], crow[l + 1]
            ga, gb = gcrow[g], gcrow[g + 1]
            gc = l2g[col[a:b]]
            o = np.argsort(gc, kind="stable")
            assert np.array_equal(gc[o], gcol[ga:gb]), "owned row pattern == global row pattern"
            assert np.array_equal(val[a:b][o], gval[ga:gb]), "owned row values are bit-identical"
        for ex in s.exchanges:
            peer = parts[ex.peer]
            back = [e for e in peer.exchanges if e.peer == s.rank]
            assert len(back) == 1
            n_recv = sum(hi - lo for lo, hi in ex.recv)
            assert n_recv == back[0].send_idx.numel()
            if n_recv:
                (lo, hi), = ex.recv
                assert lo >= s.n_owned, "halo ranges follow the owned block"
                assert np.array_equal(l2g[lo:hi], peer.l2g.numpy()[back[0].send_idx.numpy()]), "recv range == the peer's packed send, same order"
                assert np.all(back[0].send_idx.numpy() < peer.n_owned), "a rank only sends values it owns"
    assert np.all(seen == 1), "every global dof is owned exactly once"


def _worker_morton(rank, world, port, kind, dims, p, out):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    import dist_util as D
    from fealpy_b200.parallel.dist_cg import dist_cg
    from fealpy_b200.parallel.mesh_partition import MeshPartition
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        node, cell = _shuffled_mesh(kind, dims, seed=77)
        gm = O.Mesh(node, cell)
        gc2d = gm.cell_to_ipoint(p)
        gdof = gm.number_of_global_ipoints(p)
        gcrow, gcol, gval = O.assemble([(O.diffusion_element(gm, p), gc2d), (O.mass_element(gm, p), gc2d)], gdof)
        part = MeshPartition(torch.from_numpy(node), torch.from_numpy(cell), torch.from_numpy(gc2d), gdof, world, rank)
        crow, col, val = _oracle_local(part, node, cell, gc2d, p)
        xs = np.sin(0.37 * np.arange(gdof)) + 1.5
        gb = O.csr_matvec(gcrow, gcol, gval, xs)
        l2g = part.l2g.numpy()
        b = torch.from_numpy(gb[l2g].copy())
        ops = D.NumpyCgOps(crow, col, val, part.own_ranges)
        # tight tolerances: the two runs sum their inner products in different orders (local vs global numbering), so
        # they are compared where both have converged to the solution, not at the reference's loose default stop
        x, info = dist_cg(ops, b, torch.zeros_like(b), part.exchanges, check_every=4, atol=1e-14, rtol=1e-12)
        xo, oinfo = O.cg(lambda v: O.csr_matvec(gcrow, gcol, gval, v), gb, atol=1e-14, rtol=1e-12)
        own = slice(0, part.n_owned)
        err = np.linalg.norm(x.numpy()[own] - xo[l2g[own]]) / np.linalg.norm(xo)
        out[rank] = (float(err), info["niter"], oinfo["niter"], len(part.exchanges))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,kind,dims,p", [(2, "tet", (4, 3, 3), 2), (3, "tri", (8, 7), 3), (3, "tet", (4, 4, 3), 1)])
def test_distributed_cg_morton_gloo(world, kind, dims, p):
    """the distributed CG driver on a Morton partition of a shuffled mesh (packed sends, any number of neighbours), over gloo"""
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_morton, args=(world, port, kind, dims, p, out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        err, niter, oniter, nex = out[r]
        assert err < 1e-10, (r, err)
        assert abs(niter - oniter) <= 2, (niter, oniter)
    assert len({out[r][1] for r in range(world)}) == 1, "all ranks stop at the same iteration"


def test_morton_codes_order():
    from fealpy_b200.parallel.mesh_partition import morton_codes
    pts = torch.tensor([[0.0, 0.0], [1.0, 1.0], [0.0, 1.0], [1.0, 0.0], [0.49, 0.49]], dtype=torch.float64)
    c = morton_codes(pts)
    assert c[0] < c[4] < c[3] < c[2] < c[1]          # Z-order: (0,0) < near-centre < (1,0) < (0,1) < (1,1)
    pts3 = torch.rand(100, 3, dtype=torch.float64)
    assert morton_codes(pts3).unique().numel() == 100
