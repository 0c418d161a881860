"""Row partition of `TetrahedronMesh.from_box` problems into x-slabs (host logic, pure Python).

SURVEY.md section 8(e): rows are owned in the reference's GLOBAL numbering.  For from_box meshes
node ids are x-major and edge ids are ordered by their smaller node, so whole x-planes of
nodes -- and the edges leaving them -- are contiguous id ranges (appendix D).  Rank r owns
node planes [P_r, P_{r+1}) and the edges whose smaller node lies there; it assembles the cube
layers [P_r - 1, P_{r+1}) (one ghost layer of cells below, none above: the layer above touches
no owned row), and numbers its DOFs in a *window* of the global numbering:

    window-local id = global node id - first window node id                 (nodes)
                    = NNw + (global edge id - first window edge id)         (P2 edge dofs)

so local ids are monotone in global ids (column order inside a row is the global one), owned
rows are two contiguous ranges, and every halo region is a contiguous slice: the CG halo
exchange needs no pack/unpack kernels.
"""
from __future__ import annotations

from dataclasses import dataclass, field


def box_edges_before(nx, ny, nz, i, j, k):
    """edges leaving all nodes that precede node (i,j,k) in id order (csrc/topo.cu, same formula)"""
    Sb, Sc = 2 * ny + 1, 2 * nz + 1
    plane_full = 2 * Sb * Sc - (ny + 1) * (nz + 1)
    a1 = 2 if i < nx else 1
    b1 = 2 if j < ny else 1
    return i * plane_full + j * (a1 * 2 * Sc - (nz + 1)) + k * (a1 * b1 * 2 - 1)


def box_number_of_edges(nx, ny, nz):
    return box_edges_before(nx, ny, nz, nx, ny, nz)          # the last node has no outgoing edge


def box_number_of_nodes(nx, ny, nz):
    return (nx + 1) * (ny + 1) * (nz + 1)


@dataclass
class Exchange:
    peer: int
    send: list = field(default_factory=list)      # [(lo, hi)] window-local slices of owned values
    recv: list = field(default_factory=list)      # [(lo, hi)] window-local halo slices


class BoxSlab:
    """ownership, window and halo slices of rank `rank` of `world` for a (nx,ny,nz) box, degree p"""

    def __init__(self, nx, ny, nz, p, world, rank):
        if p not in (1, 2):
            raise NotImplementedError("slab partition: closed-form numbering exists for p = 1, 2")
        planes = nx + 1
        if world > planes // 2:
            raise ValueError(f"too many ranks ({world}) for {planes} node planes (need >= 2 planes per rank)")
        self.dims, self.p, self.world, self.rank = (nx, ny, nz), p, world, rank
        self.P = [(r * planes) // world for r in range(world + 1)]
        P0, P1 = self.P[rank], self.P[rank + 1]
        self.own_planes = (P0, P1)
        self.cl0, self.cl1 = max(P0 - 1, 0), min(P1, nx)             # cube layers [cl0, cl1)
        self.nyz = (ny + 1) * (nz + 1)
        self.NN, self.NE = box_number_of_nodes(nx, ny, nz), box_number_of_edges(nx, ny, nz)
        self.NNw = (self.cl1 - self.cl0 + 1) * self.nyz
        self.E0 = self.eoff(self.cl0)
        self.NEw = (self.eoff(self.cl1 + 1) - self.E0) if p == 2 else 0
        self.n_local = self.NNw + self.NEw
        self.gdof = self.NN + (self.NE if p == 2 else 0)
        self.own_nodes = ((P0 - self.cl0) * self.nyz, (P1 - self.cl0) * self.nyz)
        self.own_edges = (self._e(P0), self._e(P1)) if p == 2 else (0, 0)
        self.exchanges = []
        if rank > 0:           # lower neighbour owns plane P0 - 1 (= my first window plane)
            ex = Exchange(rank - 1)
            ex.send.append(self._node_plane(P0)); ex.recv.append(self._node_plane(P0 - 1))
            if p == 2:
                ex.send.append(self._edge_plane(P0)); ex.recv.append(self._edge_plane(P0 - 1))
            self.exchanges.append(ex)
        if rank < world - 1:   # upper neighbour owns plane P1 (= my last window plane)
            ex = Exchange(rank + 1)
            ex.send.append(self._node_plane(P1 - 1)); ex.recv.append(self._node_plane(P1))
            if p == 2:
                ex.send.append(self._edge_plane(P1 - 1)); ex.recv.append(self._edge_plane(P1))
            self.exchanges.append(ex)

    # ---- helpers ------------------------------------------------------------------------
    def eoff(self, plane):
        """global id of the first edge whose smaller node lies in `plane` (NE past the end)"""
        nx, ny, nz = self.dims
        return self.NE if plane > nx else box_edges_before(nx, ny, nz, plane, 0, 0)

    def _e(self, plane):
        return self.NNw + self.eoff(plane) - self.E0

    def _node_plane(self, plane):
        return ((plane - self.cl0) * self.nyz, (plane - self.cl0 + 1) * self.nyz)

    def _edge_plane(self, plane):
        return (self._e(plane), self._e(plane + 1))

    def row_split(self):
        """(interior, boundary): window-local (lo, hi) ranges of the OWNED rows.  Boundary rows are the dofs of the first / last
        owned node plane next to a neighbour (their cells reach into the halo plane); interior rows touch owned columns only,
        so their part of the SpMV needs no halo value and overlaps the exchange."""
        P0, P1 = self.own_planes
        a = min(P0 + (1 if self.rank > 0 else 0), P1)
        b = max(P1 - (1 if self.rank < self.world - 1 else 0), a)

        def ranges(x0, x1):
            out = []
            if x1 > x0:
                out.append((self._node_plane(x0)[0], self._node_plane(x1 - 1)[1]))
                if self.p == 2:
                    out.append((self._e(x0), self._e(x1)))
            return out
        return ranges(a, b), ranges(P0, a) + ranges(b, P1)

    @property
    def own_ranges(self):
        """(lo0, hi0, lo1, hi1) of the owned rows in window-local ids"""
        return (*self.own_nodes, *self.own_edges)

    @property
    def n_owned(self):
        return (self.own_nodes[1] - self.own_nodes[0]) + (self.own_edges[1] - self.own_edges[0])

    def local_to_global(self, l):
        """window-local dof ids -> global dof ids (works on ints, numpy arrays and torch tensors)"""
        node_off = self.cl0 * self.nyz
        is_edge = l >= self.NNw
        return l + node_off + is_edge * (self.NN + self.E0 - self.NNw - node_off)

    def owned_global_ranges(self):
        P0, P1 = self.own_planes
        nodes = (P0 * self.nyz, P1 * self.nyz)
        edges = (self.NN + self.eoff(P0), self.NN + self.eoff(P1)) if self.p == 2 else (0, 0)
        return nodes, edges
