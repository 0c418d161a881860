"""Simplex meshes living on the GPU: TriangleMesh / TetrahedronMesh.

Mirrors the part of the reference mesh API the assembly path consumes
(fealpy/mesh/triangle_mesh.py, tetrahedron_mesh.py, mesh_data_structure.py, mesh_base.py):
node/cell storage, `from_box`, the edge/face topology `construct()` builds, and the global
interpolation-point (DOF) numbering `cell_to_ipoint(p)`.  All integer work runs in the CUDA
library (csrc/topo.cu); torch tensors are only containers.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from .. import _lib
from ..basis import grad_shape_function, multi_index_matrix, number_of_local_dofs, shape_function
from ..quadrature import simplex_quadrature


class SimplexMesh:
    TD = None

    def __init__(self, node: torch.Tensor, cell: torch.Tensor):
        if not (isinstance(node, torch.Tensor) and isinstance(cell, torch.Tensor)):
            raise TypeError("node and cell must be torch tensors (PyTorch tensors are the only device containers)")
        if node.device.type != "cuda" or cell.device != node.device:
            raise RuntimeError("fealpy_b200 meshes live on a CUDA device; there is no CPU path")
        if node.dtype != torch.float64:
            raise TypeError("node must be float64")
        if cell.dtype != torch.int32:
            if cell.dtype == torch.int64:
                cell = cell.to(torch.int32)
            else:
                raise TypeError("cell must be int32 (or int64, converted)")
        if cell.ndim != 2 or cell.shape[1] != self.TD + 1 or node.ndim != 2 or node.shape[1] != self.TD:
            raise ValueError(f"expected node (NN,{self.TD}) and cell (NC,{self.TD + 1})")
        self.node = node.contiguous()
        self.cell = cell.contiguous()
        self.itype = torch.int32
        self.ftype = torch.float64
        self.device = node.device
        self._edge = None
        self._cell2edge = None
        self._face = None
        self._cell2face = None
        self._c2ip = {}
        self.meshdata = {}

    # ---- sizes -----------------------------------------------------------------------
    def top_dimension(self): return self.TD
    def geo_dimension(self): return self.node.shape[1]
    def number_of_nodes(self): return self.node.shape[0]
    def number_of_cells(self): return self.cell.shape[0]
    def number_of_edges(self):
        if self._edge is None and self._closed_form_box():
            nx, ny, nz = self._box_dims
            return int(_lib.load().fb2_box_edges_before(nx, ny, nz, nx, ny, nz))
        return self.edge.shape[0]

    def _closed_form_box(self):
        """True for an untouched TetrahedronMesh.from_box mesh (FB2_BOX_CLOSED_FORM=0 forces the generic sort path)"""
        return (self.TD == 3 and getattr(self, "_box_dims", None) is not None
                and os.environ.get("FB2_BOX_CLOSED_FORM", "1") != "0")
    def number_of_faces(self): return self.face.shape[0]

    def entity(self, etype):
        if etype in ("node", 0): return self.node
        if etype in ("cell", self.TD): return self.cell
        if etype in ("edge", 1): return self.edge
        if etype in ("face", self.TD - 1): return self.face
        raise ValueError(f"unknown entity type {etype}")

    # ---- topology (MeshDS.construct, mesh/mesh_data_structure.py:428-464) ----------------
    def _build_entities(self, kind):
        lib = _lib.load()
        NC, NN = self.number_of_cells(), self.number_of_nodes()
        per_cell = {(2, 1): 3, (3, 1): 6, (3, 2): 4}[(self.TD, kind)]
        nve = 2 if kind == 1 else 3
        ws = _lib.workspace(lib.fb2_entity_workspace_bytes(NC, per_cell), self.device)
        c2e = torch.empty((NC, per_cell), dtype=torch.int32, device=self.device)
        cnt = C.c_int64(0)
        _lib.call("fb2_build_entities", _lib.ptr(self.cell), NC, self.TD, kind, NN, _lib.ptr(c2e), C.byref(cnt),
                  _lib.ptr(ws), _lib.stream())
        ent = torch.empty((cnt.value, nve), dtype=torch.int32, device=self.device)
        _lib.call("fb2_entities_emit", _lib.ptr(self.cell), NC, self.TD, kind, _lib.ptr(c2e), _lib.ptr(ent), _lib.ptr(ws),
                  _lib.stream())
        return ent, c2e

    @property
    def edge(self):
        if self._edge is None:
            self._edge, self._cell2edge = self._build_entities(1)
        return self._edge

    @property
    def cell2edge(self):
        self.edge
        return self._cell2edge

    @property
    def face(self):
        if self.TD == 2:
            return self.edge
        if self._face is None:
            self._face, self._cell2face = self._build_entities(2)
        return self._face

    @property
    def cell2face(self):
        if self.TD == 2:
            return self.cell2edge
        self.face
        return self._cell2face

    def cell_to_edge(self): return self.cell2edge
    def cell_to_face(self): return self.cell2face

    # ---- interpolation points / dof numbering -----------------------------------------------
    def multi_index_matrix(self, p, TD=None):
        return multi_index_matrix(p, self.TD if TD is None else TD)

    def number_of_local_ipoints(self, p, iptype="cell"):
        if iptype in ("cell", self.TD):
            return number_of_local_dofs(self.TD, p)
        if iptype in ("face", self.TD - 1) and self.TD == 3:
            return (p + 1) * (p + 2) // 2
        return p + 1

    def number_of_global_ipoints(self, p):
        NN, NC = self.number_of_nodes(), self.number_of_cells()
        if p == 1:
            return NN
        n = NN + (p - 1) * self.number_of_edges()
        if self.TD == 2:
            return n + (p - 1) * (p - 2) // 2 * NC
        if p >= 3:
            n += (p - 1) * (p - 2) // 2 * self.number_of_faces() + (p - 1) * (p - 2) * (p - 3) // 6 * NC
        return n

    def cell_to_ipoint(self, p, index=None):
        """(NC, ldof) int32 global dof ids (mesh/triangle_mesh.py:218-270, tetrahedron_mesh.py:388-441)."""
        if p not in self._c2ip:
            if p == 1:
                self._c2ip[p] = self.cell
            elif p == 2 and self._closed_form_box():
                # from_box tetrahedra: edge id = number of edges leaving smaller nodes (csrc/topo.cu) -- the same ids the
                # sorted-unique construction gives (tests: test_from_box_closed_form_numbering), without sorting 75 M keys
                nx, ny, nz = self._box_dims
                out = torch.empty((self.number_of_cells(), 10), dtype=torch.int32, device=self.device)
                b = (C.c_double * 6)(*[float(v) for v in self.box])
                with torch.cuda.device(self.device):
                    _lib.call("fb2_tet_box_slab", b, nx, ny, nz, 0, nx, 2, None, None, _lib.ptr(out), _lib.stream())
                self._c2ip[p] = out
            else:
                if p > 3:
                    raise NotImplementedError("fealpy_b200 supports Lagrange degree p = 1..3")
                NC, L = self.number_of_cells(), number_of_local_dofs(self.TD, p)
                need_face = self.TD == 3 and p >= 3
                c2f = self.cell2face if need_face else None
                NF = self.number_of_faces() if need_face else 0
                mi = np.ascontiguousarray(self.multi_index_matrix(p).astype(np.uint8))
                out = torch.empty((NC, L), dtype=torch.int32, device=self.device)
                _lib.call("fb2_cell_to_dof", _lib.ptr(self.cell), _lib.ptr(self.cell2edge), _lib.ptr(self.edge), _lib.ptr(c2f),
                          NC, self.TD, p, self.number_of_nodes(), self.number_of_edges(), NF,
                          mi.ctypes.data_as(C.c_void_p), L, _lib.ptr(out), _lib.stream())
                self._c2ip[p] = out
        c2d = self._c2ip[p]
        return c2d if index is None else c2d[index]

    def quadrature_formula(self, q, etype="cell"):
        if etype not in ("cell", self.TD):
            raise NotImplementedError("only cell quadrature is on the accelerated path")
        return simplex_quadrature(self.TD, q)

    def bc_to_point(self, bcs, index=None):
        """physical points of barycentric points, (NC, NQ, GD) (mesh/mesh_base.py:454-478, backend/numpy_backend.py:401-407)."""
        cell = (self.cell if index is None else self.cell[index]).contiguous()
        b = torch.as_tensor(bcs, dtype=torch.float64, device=self.device).contiguous()
        if b.ndim != 2 or b.shape[1] != self.TD + 1:
            raise ValueError(f"bc_to_point: barycentric points must be (NQ, {self.TD + 1})")
        out = torch.empty((cell.shape[0], b.shape[0], self.TD), dtype=torch.float64, device=self.device)
        _lib.call("fb2_bc_to_points", self.TD, cell.shape[0], b.shape[0], _lib.ptr(self.node), _lib.ptr(cell), _lib.ptr(b),
                  _lib.ptr(out), _lib.stream())
        return out

    def interpolation_points(self, p, index=None):
        """(gdof, GD) coordinates of the global interpolation points.  `index` (optional, a bool flag over the points): only
        the flagged rows are needed -- they are computed from the cells that touch them, the other rows stay zero (the
        Dirichlet set-up asks for the boundary points of a 17 M-dof space: ~1 % of the cells)."""
        if p == 1:                          # the nodes themselves (a 12.6 M-cell gather + scatter took 212 ms for nothing)
            return self.node.clone()
        gdof = self.number_of_global_ipoints(p)
        mi = torch.as_tensor(self.multi_index_matrix(p) / p, dtype=torch.float64, device=self.device)
        cell, c2i = self.cell.long(), self.cell_to_ipoint(p).long()
        if index is not None:
            sel = index[c2i].any(dim=1).nonzero().reshape(-1)
            cell, c2i = cell[sel], c2i[sel]
        pts = torch.einsum("cjk,ij->cik", self.node[cell], mi)
        ip = torch.zeros((gdof, self.geo_dimension()), dtype=torch.float64, device=self.device)
        ip[c2i.reshape(-1)] = pts.reshape(-1, self.geo_dimension())
        ip[: self.number_of_nodes()] = self.node
        return ip

    def boundary_face_flag(self):
        cnt = torch.bincount(self.cell2face.reshape(-1).long(), minlength=self.number_of_faces())
        return cnt == 1

    def boundary_face_index(self):
        return self.boundary_face_flag().nonzero().reshape(-1)

    def entity_barycenter(self, etype="cell", index=None):
        """mean of the entity's vertices (mesh/mesh_base.py entity_barycenter)"""
        if etype in ("node", 0):
            return self.node if index is None else self.node[index]
        ent = self.entity(etype)
        ent = ent if index is None else ent[index]
        return self.node[ent.long()].mean(dim=1)

    # ---- cell geometry as public device arrays ---------------------------------
    # The arithmetic is the element kernels' own (csrc/elem.cu load_geo): one record per cell holds
    # grad(lambda_k) for k = 0..TD and the signed measure.
    def _cells(self, index):
        if index is None or (isinstance(index, slice) and index == slice(None)):
            return self.cell
        return self.cell[index].reshape(-1, self.TD + 1).contiguous()

    def cell_gradient_records(self, index=None):
        """(NC, (TD+1)*TD + 1) float64: grad lambda (k-major) then the signed cell measure"""
        cell = self._cells(index)
        NC, TD = cell.shape[0], self.TD
        out = torch.empty((NC, (TD + 1) * TD + 1), dtype=torch.float64, device=self.device)
        _lib.call("fb2_cell_gradients", TD, NC, _lib.ptr(self.node), _lib.ptr(cell), _lib.ptr(out), _lib.stream())
        return out

    def entity_measure(self, etype="cell", index=None):
        """mesh/triangle_mesh.py:45-70, mesh/tetrahedron_mesh.py:144-175 (signed cell measure, never abs'ed)"""
        if etype in ("cell", self.TD):
            return self.cell_gradient_records(index)[:, -1].contiguous()
        if etype in ("node", 0):
            return torch.zeros(1, dtype=torch.float64, device=self.device)
        if etype in ("edge", 1):
            e = self.edge if index is None else self.edge[index]
            v = self.node[e[:, 1].long()] - self.node[e[:, 0].long()]
            return torch.sqrt((v * v).sum(dim=1))
        if etype in ("face", 2) and self.TD == 3:
            f = self.face if index is None else self.face[index]
            v01 = self.node[f[:, 1].long()] - self.node[f[:, 0].long()]
            v02 = self.node[f[:, 2].long()] - self.node[f[:, 0].long()]
            nv = torch.linalg.cross(v01, v02)
            return torch.sqrt((nv * nv).sum(dim=1)) / 2.0
        raise ValueError(f"entity type: {etype} is wrong!")

    def grad_lambda(self, index=None, TD=None):
        """(NC, TD+1, GD) gradients of the barycentric coordinates (mesh/triangle_mesh.py:115-129,
        mesh/tetrahedron_mesh.py:208-219)"""
        if TD is not None and TD != self.TD:
            raise NotImplementedError("grad_lambda of sub-entities is not on the accelerated path")
        rec = self.cell_gradient_records(index)
        return rec[:, :-1].reshape(rec.shape[0], self.TD + 1, self.TD).contiguous()

    def shape_function(self, bcs, p=1, *, index=None, mi=None):
        """phi (NQ, ldof) at barycentric points (mesh/mesh_base.py:687-711 -> bm.simplex_shape_function);
        cell independent, evaluated once on the host in float64 and uploaded"""
        b = bcs.detach().cpu().numpy() if isinstance(bcs, torch.Tensor) else np.asarray(bcs, dtype=np.float64)
        return torch.from_numpy(np.ascontiguousarray(shape_function(b, p))).to(self.device)

    GRAD_SHAPE_DEFAULT = "u"

    def grad_shape_function(self, bcs, p=1, *, index=None, variables=None, mi=None):
        """variables='u': R (NQ, ldof, TD+1) = dphi/dlambda; 'x': (NC, NQ, ldof, GD) physical gradients
        (mesh/mesh_base.py:713-749; TriangleMesh defaults to 'x', mesh/triangle_mesh.py:142-153)"""
        variables = self.GRAD_SHAPE_DEFAULT if variables is None else variables
        b = bcs.detach().cpu().numpy() if isinstance(bcs, torch.Tensor) else np.asarray(bcs, dtype=np.float64)
        if b.ndim != 2 or b.shape[1] != self.TD + 1:
            raise ValueError(f"barycentric points must be (NQ, {self.TD + 1})")
        R = torch.from_numpy(np.ascontiguousarray(grad_shape_function(b, p))).to(self.device)
        if variables == "u":
            return R
        if variables != "x":
            raise ValueError(f"Variables type is expected to be 'u' or 'x', but got '{variables}'.")
        rec = self.cell_gradient_records(index)
        NC, NQ, L = rec.shape[0], R.shape[0], R.shape[1]
        out = torch.empty((NC, NQ, L, self.TD), dtype=torch.float64, device=self.device)
        _lib.call("fb2_grad_basis", self.TD, NC, NQ, L, _lib.ptr(rec), _lib.ptr(R), _lib.ptr(out), _lib.stream())
        return out


class TriangleMesh(SimplexMesh):
    TD = 2
    GRAD_SHAPE_DEFAULT = "x"

    def cell_area(self, index=None):
        return self.entity_measure("cell", index)

    @classmethod
    def from_box(cls, box=(0, 1, 0, 1), nx=10, ny=10, *, threshold=None, itype=None, ftype=None, device="cuda"):
        """mesh/triangle_mesh.py:1386-1435"""
        if threshold is not None:
            raise NotImplementedError("from_box(threshold=...) is not on the accelerated path")
        _lib.require_cuda()
        device = torch.device(device if device is not None else "cuda")
        NN, NC = (nx + 1) * (ny + 1), 2 * nx * ny
        node = torch.empty((NN, 2), dtype=torch.float64, device=device)
        cell = torch.empty((NC, 3), dtype=torch.int32, device=device)
        with torch.cuda.device(device):
            b = (C.c_double * 4)(*[float(v) for v in box])
            _lib.call("fb2_tri_from_box", b, nx, ny, _lib.ptr(node), _lib.ptr(cell), _lib.stream())
        return cls(node, cell)


class TetrahedronMesh(SimplexMesh):
    TD = 3

    def cell_volume(self, index=None):
        return self.entity_measure("cell", index)

    @classmethod
    def from_box(cls, box=(0, 1, 0, 1, 0, 1), nx=10, ny=10, nz=10, threshold=None, device="cuda"):
        """mesh/tetrahedron_mesh.py:1016-1086"""
        if threshold is not None:
            raise NotImplementedError("from_box(threshold=...) is not on the accelerated path")
        _lib.require_cuda()
        device = torch.device(device if device is not None else "cuda")
        NN, NC = (nx + 1) * (ny + 1) * (nz + 1), 6 * nx * ny * nz
        node = torch.empty((NN, 3), dtype=torch.float64, device=device)
        cell = torch.empty((NC, 4), dtype=torch.int32, device=device)
        with torch.cuda.device(device):
            b = (C.c_double * 6)(*[float(v) for v in box])
            _lib.call("fb2_tet_from_box", b, nx, ny, nz, _lib.ptr(node), _lib.ptr(cell), _lib.stream())
        m = cls(node, cell)
        m.box = list(box)
        m._box_dims = (int(nx), int(ny), int(nz))      # closed-form edge count / P2 numbering (no edge sort on the assembly path)
        return m
