"""Row partition of an ARBITRARY simplex mesh + Lagrange space over the ranks of a job (replaces
what fealpy/mesh/parallel.py:15-115 does for the reference's MPI model: split cells, find shared entities, local ids).

    cells  -> ranks : Morton order of the cell barycentres, cut into `world` chunks of equal cell count
    dofs   -> ranks : owner(dof) = smallest rank among the cells that touch it
    rank r assembles every cell that touches a dof it owns (its own cells + one ghost layer, recomputed redundantly:
    no assembly communication) and keeps the owned rows; local cells keep their GLOBAL relative order, so every owned
    row is summed in the same (local index, cell) order as on one GPU and its values are bit-identical.
    local dof numbering : [ owned (ascending global id) | halo of neighbour q0 | halo of neighbour q1 | ... ]
    CG halo exchange    : per neighbour one packed send (gather of the owned dofs it needs) and one receive that lands
                          directly in that neighbour's contiguous halo range -- any number of neighbours.

Host logic on integer tensors (torch ops on whatever device the mesh lives on; CPU in the gloo tests).  Every rank holds
the global mesh while partitioning -- fine up to config-2 sizes; the 200 M-cell box of config 5 uses the closed-form slabs
of box_partition.py instead.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch


def morton_codes(points: torch.Tensor) -> torch.Tensor:
    """interleaved-bit (Z-order) code of points (N, 2 or 3), quantised to 21 (3-D) / 31 (2-D) bits per axis"""
    GD = points.shape[1]
    bits = 21 if GD == 3 else 31
    lo, hi = points.min(dim=0).values, points.max(dim=0).values
    span = torch.where(hi > lo, hi - lo, torch.ones_like(hi))
    q = ((points - lo) / span * float((1 << bits) - 1)).to(torch.int64).clamp_(0, (1 << bits) - 1)
    code = torch.zeros(points.shape[0], dtype=torch.int64, device=points.device)
    for b in range(bits):
        for d in range(GD):
            code |= ((q[:, d] >> b) & 1) << (GD * b + d)
    return code


@dataclass
class PackedExchange:
    peer: int
    send_idx: torch.Tensor                         # local ids (owned) whose values the peer needs, ascending global id
    recv: list = field(default_factory=list)      # [(lo, hi)] the peer's contiguous halo range in my numbering
    send: list = field(default_factory=list)      # unused (slab partitions send contiguous slices instead)


class MeshPartition:
    """ownership, local numbering and exchange lists of rank `rank` of `world` for (cell, cell2dof)"""

    def __init__(self, node, cell, cell2dof, gdof, world, rank):
        dev = cell.device
        self.world, self.rank, self.gdof = world, rank, int(gdof)
        NC = cell.shape[0]
        c2d = cell2dof.long()
        # cells -> ranks (Morton order, equal counts)
        bary = node[cell.long()].mean(dim=1)
        order = torch.argsort(morton_codes(bary), stable=True)
        cell_rank = torch.empty(NC, dtype=torch.int64, device=dev)
        cell_rank[order] = (torch.arange(NC, device=dev, dtype=torch.int64) * world) // max(NC, 1)
        # dofs -> ranks
        owner = torch.full((self.gdof,), world, dtype=torch.int64, device=dev)
        owner.scatter_reduce_(0, c2d.reshape(-1), cell_rank.repeat_interleave(c2d.shape[1]), reduce="amin")
        self.owner, self.cell_rank = owner, cell_rank

        def cells_of(r):                            # cells that touch a dof owned by r, in global order
            return ((owner[c2d] == r).any(dim=1)).nonzero().reshape(-1)

        def dofs_of(cells):
            f = torch.zeros(self.gdof, dtype=torch.bool, device=dev)
            f[c2d[cells].reshape(-1)] = True
            return f
        self.cells = cells_of(rank)
        touched = dofs_of(self.cells)
        owned = (touched & (owner == rank)).nonzero().reshape(-1)
        self.n_owned = int(owned.numel())
        parts, self.exchanges, off = [owned], [], self.n_owned
        nbrs = torch.unique(owner[touched & (owner != rank)]).tolist()
        for q in nbrs:                              # my halo: dofs owned by q that my cells touch
            h = (touched & (owner == q)).nonzero().reshape(-1)
            parts.append(h)
            self.exchanges.append(PackedExchange(int(q), None, recv=[(off, off + int(h.numel()))]))
            off += int(h.numel())
        self.l2g = torch.cat(parts)
        self.n_local = int(self.l2g.numel())
        g2l = torch.full((self.gdof,), -1, dtype=torch.int64, device=dev)
        g2l[self.l2g] = torch.arange(self.n_local, device=dev, dtype=torch.int64)
        self.g2l = g2l
        # what each neighbour needs from me = its halo owned by me (it computes the same set as its recv list, same order).
        # A rank that needs my dofs also owns dofs I need? not necessarily -- so the senders are found from the peers' side.
        for q in range(world):
            if q == rank:
                continue
            need = (dofs_of(cells_of(q)) & (owner == rank)).nonzero().reshape(-1)       # ascending global id
            if need.numel() == 0:
                continue
            ex = next((e for e in self.exchanges if e.peer == q), None)
            if ex is None:
                ex = PackedExchange(q, None, recv=[])
                self.exchanges.append(ex)
            ex.send_idx = g2l[need].contiguous()
        self.exchanges.sort(key=lambda e: e.peer)
        for ex in self.exchanges:
            if ex.send_idx is None:
                ex.send_idx = torch.empty(0, dtype=torch.int64, device=dev)
        self.own_ranges = (0, self.n_owned, 0, 0)

    def local_to_global(self, l):
        return self.l2g[l]

    def local_cell2dof(self, cell2dof):
        return self.g2l[cell2dof[self.cells].long()].to(cell2dof.dtype).contiguous()

    def local_mesh(self, node, cell):
        """(node_local, cell_local): the nodes of my cells renumbered in ascending global order"""
        lc = cell[self.cells].long()
        used = torch.unique(lc.reshape(-1))
        remap = torch.full((node.shape[0],), -1, dtype=torch.int64, device=cell.device)
        remap[used] = torch.arange(used.numel(), device=cell.device, dtype=torch.int64)
        return node[used].contiguous(), remap[lc].to(cell.dtype).contiguous()


class PartitionedProblem:
    """local mesh + space of one rank of a Morton partition of a global fealpy_b200 mesh / LagrangeFESpace (p = 1..3)"""

    def __init__(self, mesh, space, world, rank):
        from ..mesh import TetrahedronMesh, TriangleMesh
        from .slab_problem import WindowSpace
        c2d = space.cell_to_dof()
        self.part = part = MeshPartition(mesh.node, mesh.cell, c2d, space.number_of_global_dofs(), world, rank)
        node_l, cell_l = part.local_mesh(mesh.node, mesh.cell)
        cls = TetrahedronMesh if mesh.TD == 3 else TriangleMesh
        self.mesh = cls(node_l, cell_l)
        self.space = WindowSpace(self.mesh, space.p, part.local_cell2dof(c2d), part.n_local)
