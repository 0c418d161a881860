"""Hardware parity check of the multi-GPU path, run by every rank of a `torch.distributed` job before anything is timed
(bench.py --gpus N emits its result as `"verify": {...}`; tests/test_multi_gpu.py asserts on it).

A small slab problem (default: a (24*world) x 12 x 12 box, P2) is assembled by every rank on its slab AND, redundantly,
as the whole box on the same GPU (the single-GPU path, itself pinned against the reference's golden matrices).  Checked:
  * the rank's owned rows -- pattern mapped back to global ids, and values -- are BIT-identical to the single-GPU rows;
  * the distributed CG (halo exchange + reductions over the process group) run to the reference tolerances needs the
    same number of iterations (+-1) as the single-GPU fb2_cg and returns the same solution (relative L2 <= 1e-10).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def verify_slab(world, rank, device, *, p=2, dims_per_rank=(24, 12, 12), solver_factory=None, group=None):
    from ..fem import BilinearForm, ScalarDiffusionIntegrator, ScalarMassIntegrator
    from ..functionspace import LagrangeFESpace
    from ..mesh import TetrahedronMesh
    from ..solver import cg
    from .dist_cg import DistCG
    from .slab_problem import SlabProblem

    nx, ny, nz = dims_per_rank[0] * world, dims_per_rank[1], dims_per_rank[2]
    box = [0.0, float(world), 0.0, 1.0, 0.0, 1.0]

    def form(space):
        bf = BilinearForm(space)
        bf.add_integrator(ScalarDiffusionIntegrator())
        bf.add_integrator(ScalarMassIntegrator())
        return bf.assembly()

    # single-GPU matrix of the whole box
    gmesh = TetrahedronMesh.from_box(box, nx, ny, nz, device=device)
    G = form(LagrangeFESpace(gmesh, p))
    # this rank's slab
    sp = SlabProblem(box, nx, ny, nz, p, world, rank, device=device)
    part = sp.part
    A = form(sp.space)
    l2g = part.local_to_global(torch.arange(part.n_local, device=device, dtype=torch.int64))
    same = True
    for lo, hi in (part.own_nodes, part.own_edges):
        if hi <= lo:
            continue
        g0, g1 = int(l2g[lo]), int(l2g[hi - 1]) + 1
        a, b = int(A.crow[lo]), int(A.crow[hi])
        ga, gb = int(G.crow[g0]), int(G.crow[g1])
        same &= (b - a) == (gb - ga)
        if not same:
            break
        same &= bool(torch.equal(A.crow[lo:hi + 1] - a, G.crow[g0:g1 + 1] - ga))
        same &= bool(torch.equal(l2g[A.col[a:b].long()], G.col[ga:gb].long()))
        same &= bool(torch.equal(A.values[a:b], G.values[ga:gb]))              # bit-identical
    # CG: distributed vs single GPU, reference tolerances, right-hand side with a non-trivial solution
    gdof = G.shape[0]
    xs = torch.sin(torch.arange(gdof, device=device, dtype=torch.float64) * 0.37) + 1.5
    bg = G @ xs
    xg, ginfo = cg(G, bg, returninfo=True)
    own = torch.zeros(part.n_local, dtype=torch.bool, device=device)
    own[part.own_nodes[0]:part.own_nodes[1]] = True
    own[part.own_edges[0]:part.own_edges[1]] = True
    bl = torch.zeros(part.n_local, dtype=torch.float64, device=device)
    bl[own] = bg[l2g[own]]
    solver = (solver_factory or DistCG)(A, part, group=group)
    xl, linfo = solver.solve(bl)
    num = (xl[own] - xg[l2g[own]]).pow(2).sum()
    den = xg[l2g[own]].pow(2).sum()
    acc = torch.stack([num, den])
    flags = torch.tensor([int(same), int(abs(linfo["niter"] - ginfo["niter"]) <= 1)], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN, group=group)
    x_rel = float((acc[0] / acc[1]).sqrt())
    out = {"problem": f"tet P{p} from_box {nx}x{ny}x{nz} in {world} x-slabs", "gdof": gdof, "nnz": G.nnz,
           "owned_rows_bit_identical": bool(flags[0]), "niter": int(linfo["niter"]), "niter_single_gpu": int(ginfo["niter"]),
           "niter_ok": bool(flags[1]), "x_rel": x_rel, "mode": getattr(solver, "mode", "nccl")}
    out["ok"] = out["owned_rows_bit_identical"] and out["niter_ok"] and x_rel <= 1e-10
    return out


def verify_partition(mesh, space, world, rank, device, *, solver_factory=None, group=None, solve=True):
    """The same check for a GENERAL mesh split by parallel/mesh_partition.py (Morton order, packed halos, any number of
    neighbours): this rank's owned rows are bit-identical to the rows of the single-GPU matrix of (mesh, space), and the
    distributed CG agrees with the single-GPU fb2_cg.  With world == 1 process groups are not needed (solve=False checks the
    rows only: one GPU can play every rank in turn)."""
    from ..fem import BilinearForm, ScalarDiffusionIntegrator, ScalarMassIntegrator
    from ..solver import cg
    from .dist_cg import DistCG
    from .mesh_partition import PartitionedProblem

    def form(sp):
        bf = BilinearForm(sp)
        bf.add_integrator(ScalarDiffusionIntegrator())
        bf.add_integrator(ScalarMassIntegrator())
        return bf.assembly()

    G = form(space)
    pp = PartitionedProblem(mesh, space, world, rank)
    part = pp.part
    A = form(pp.space)
    l2g = part.l2g
    no = part.n_owned
    a1 = int(A.crow[no])
    grow = l2g[:no]                                             # owned rows ascend in the global numbering
    glen = G.crow[grow + 1] - G.crow[grow]
    same = bool(torch.equal(A.crow[1:no + 1] - A.crow[:no], glen))
    if same:
        # global value index of every owned local entry: row start in G + position inside the row (local columns, mapped to
        # global ids, are a permutation of the global row: sort them per row)
        lrow = torch.repeat_interleave(torch.arange(no, device=device), (A.crow[1:no + 1] - A.crow[:no]))
        gcol = l2g[A.col[:a1].long()]
        key = lrow * (G.shape[0] + 1) + gcol
        order = torch.argsort(key, stable=True)
        gidx_start = G.crow[grow][lrow[order]]
        pos = torch.arange(a1, device=device) - A.crow[:no][lrow[order]]
        gidx = gidx_start + pos
        same &= bool(torch.equal(gcol[order].to(G.col.dtype), G.col[gidx]))
        same &= bool(torch.equal(A.values[:a1][order], G.values[gidx]))          # bit-identical
    out = {"problem": f"{type(mesh).__name__} P{space.p}, {mesh.number_of_cells()} cells in {world} Morton parts", "gdof": G.shape[0],
           "nnz": G.nnz, "n_owned": no, "n_halo": part.n_local - no, "neighbours": len(part.exchanges), "owned_rows_bit_identical": same}
    if not solve:
        out["ok"] = same
        return out
    gdof = G.shape[0]
    xs = torch.sin(torch.arange(gdof, device=device, dtype=torch.float64) * 0.37) + 1.5
    bg = G @ xs
    xg, ginfo = cg(G, bg, atol=1e-14, rtol=1e-12, returninfo=True)
    bl = torch.zeros(part.n_local, dtype=torch.float64, device=device)
    bl[:no] = bg[grow]
    solver = (solver_factory or DistCG)(A, part, group=group)
    xl, linfo = solver.solve(bl, atol=1e-14, rtol=1e-12)
    acc = torch.stack([(xl[:no] - xg[grow]).pow(2).sum(), xg[grow].pow(2).sum()])
    flags = torch.tensor([int(same), int(abs(linfo["niter"] - ginfo["niter"]) <= 2)], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN, group=group)
    x_rel = float((acc[0] / acc[1]).sqrt())
    out.update({"owned_rows_bit_identical": bool(flags[0]), "niter": int(linfo["niter"]), "niter_single_gpu": int(ginfo["niter"]),
                "niter_ok": bool(flags[1]), "x_rel": x_rel})
    out["ok"] = out["owned_rows_bit_identical"] and out["niter_ok"] and x_rel <= 1e-10
    return out
