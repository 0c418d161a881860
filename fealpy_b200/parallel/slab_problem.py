"""Local (per-rank) mesh + space of a slab partition, built on the device in closed form."""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib
from ..basis import number_of_local_dofs
from ..mesh import TetrahedronMesh
from .box_partition import BoxSlab


class WindowSpace:
    """LagrangeFESpace-shaped view of the slab: dofs numbered in the rank's window of the global
    numbering (see box_partition.py).  BilinearForm / integrators only read the members below."""

    def __init__(self, mesh, p, cell2dof, n_local):
        self.mesh, self.p = mesh, p
        self._c2d, self._n = cell2dof, n_local
        self.itype, self.ftype, self.device = mesh.itype, mesh.ftype, mesh.device
        self.TD, self.GD = mesh.TD, mesh.node.shape[1]
        self.ctype = "C"

    def number_of_global_dofs(self):
        return self._n

    def number_of_local_dofs(self, doftype="cell"):
        return number_of_local_dofs(self.TD, self.p)

    def cell_to_dof(self, index=None):
        return self._c2d if index is None else self._c2d[index]

    def geo_dimension(self): return self.GD
    def top_dimension(self): return self.TD


class SlabProblem:
    def __init__(self, box, nx, ny, nz, p, world, rank, device="cuda"):
        _lib.require_cuda()
        self.part = part = BoxSlab(nx, ny, nz, p, world, rank)
        dev = torch.device(device)
        NC = 6 * (part.cl1 - part.cl0) * ny * nz
        L = number_of_local_dofs(3, p)
        node = torch.empty((part.NNw, 3), dtype=torch.float64, device=dev)
        cell = torch.empty((NC, 4), dtype=torch.int32, device=dev)
        c2d = torch.empty((NC, L), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            b = (C.c_double * 6)(*[float(v) for v in box])
            _lib.call("fb2_tet_box_slab", b, nx, ny, nz, part.cl0, part.cl1, p, _lib.ptr(node), _lib.ptr(cell), _lib.ptr(c2d),
                      _lib.stream())
        self.mesh = TetrahedronMesh(node, cell)
        self.space = WindowSpace(self.mesh, p, c2d, part.n_local)
