"""LagrangeFESpace / TensorFunctionSpace on simplex meshes.

Mirrors fealpy/functionspace/lagrange_fe_space.py:17-157 (+ dofs.py:16-84) and
tensor_space.py:13-102 for the members the assembly path reads: p, mesh, itype/ftype,
number_of_local_dofs / number_of_global_dofs, cell_to_dof, is_boundary_dof,
interpolation_points, dof_priority.
"""
from __future__ import annotations

import torch

from .. import _lib


class LagrangeFESpace:
    def __init__(self, mesh, p: int = 1, ctype: str = "C"):
        if ctype != "C":
            raise NotImplementedError("discontinuous spaces (ctype='D') are not on the accelerated path")
        if p < 1 or p > 3:
            raise NotImplementedError("fealpy_b200 supports Lagrange degree p = 1..3")
        self.mesh = mesh
        self.p = p
        self.ctype = ctype
        self.itype = mesh.itype
        self.ftype = mesh.ftype
        self.device = mesh.device
        self.TD = mesh.top_dimension()
        self.GD = mesh.geo_dimension()

    def __str__(self):
        return "Lagrange finite element space on linear mesh!"

    def number_of_local_dofs(self, doftype="cell"):
        return self.mesh.number_of_local_ipoints(self.p, doftype)

    def number_of_global_dofs(self):
        return self.mesh.number_of_global_ipoints(self.p)

    def cell_to_dof(self, index=None):
        return self.mesh.cell_to_ipoint(self.p, index=index)

    def interpolation_points(self, index=None):
        return self.mesh.interpolation_points(self.p, index=index)

    def geo_dimension(self): return self.GD
    def top_dimension(self): return self.TD

    def basis(self, bc, index=None):
        """phi (1, NQ, ldof) (functionspace/lagrange_fe_space.py:146-148)"""
        return self.mesh.shape_function(bc, self.p, index=index)[None, ...]

    face_basis = basis
    edge_basis = basis

    def grad_basis(self, bc, index=None, variable="x"):
        """(NC, NQ, ldof, GD) for variable='x', R (NQ, ldof, TD+1) for 'u' (functionspace/lagrange_fe_space.py:153-154)"""
        return self.mesh.grad_shape_function(bc, self.p, index=index, variables=variable)

    def _boundary_face_dofs(self, keep_face=None):
        """flag of the dofs lying on boundary faces (optionally only the faces flagged in keep_face): the dofs of
        local face lf of a cell are the multi-indices with a zero in position lf (= face_to_dof of that face)"""
        mesh, p = self.mesh, self.p
        gdof = self.number_of_global_dofs()
        bd_face = mesh.boundary_face_flag()
        if keep_face is not None:
            bd_face = bd_face & keep_face
        c2d = self.cell_to_dof().long()
        c2f = mesh.cell2face.long()
        mi = torch.as_tensor(mesh.multi_index_matrix(p), device=self.device)
        flag = torch.zeros(gdof, dtype=torch.bool, device=self.device)
        for lf in range(self.TD + 1):                 # local face lf is opposite vertex lf
            on_face = (mi[:, lf] == 0).nonzero().reshape(-1)
            cells = bd_face[c2f[:, lf]].nonzero().reshape(-1)
            if cells.numel():
                flag[c2d[cells][:, on_face].reshape(-1)] = True
        return flag

    def is_boundary_dof(self, threshold=None, method=None):
        """functionspace/dofs.py:23-55.  method None / 'centroid': a callable threshold is evaluated on the
        barycentres of the boundary FACES and every dof of a kept face is marked; 'interp': it is evaluated on
        the interpolation points of the boundary dofs.  (threshold callables run as torch ops.)"""
        gdof = self.number_of_global_dofs()
        if isinstance(threshold, torch.Tensor):
            if threshold.dtype == torch.bool and threshold.numel() == gdof:
                return threshold
            raise ValueError(f"Unknown threshold: {threshold}")
        if method is None or method == "centroid":
            keep = None
            if callable(threshold):
                mesh = self.mesh
                idx = mesh.boundary_face_index()
                sel = torch.as_tensor(threshold(mesh.entity_barycenter("face", idx)), device=self.device).to(torch.bool)
                keep = torch.zeros(mesh.number_of_faces(), dtype=torch.bool, device=self.device)
                keep[idx[sel]] = True
            return self._boundary_face_dofs(keep)
        if method == "interp":
            flag = self._boundary_face_dofs()
            if callable(threshold):
                idx = flag.nonzero().reshape(-1)
                sel = torch.as_tensor(threshold(self.interpolation_points(index=flag)[idx]), device=self.device).to(torch.bool)
                flag = torch.zeros_like(flag)
                flag[idx[sel]] = True
            return flag
        raise ValueError(f"Unknown method: {method}")

    def boundary_interpolate(self, gd, uh=None, *, threshold=None, method=None):
        """functionspace/lagrange_fe_space.py:111-142"""
        isD = self.is_boundary_dof(threshold=threshold, method="interp")
        if isinstance(gd, torch.Tensor):
            if uh is None:
                uh = torch.zeros_like(gd)
            uh[..., isD] = gd[isD]
        elif callable(gd):
            val = gd(self.interpolation_points(index=isD)[isD])
            if uh is None:
                uh = torch.zeros(self.number_of_global_dofs(), dtype=self.ftype, device=self.device)
            uh[..., isD] = val
        elif isinstance(gd, (int, float)):
            if uh is None:
                uh = torch.zeros(self.number_of_global_dofs(), dtype=self.ftype, device=self.device)
            uh[..., isD] = float(gd)
        else:
            raise TypeError("gd must be a tensor or a callable function")
        return uh, isD

    set_dirichlet_bc = boundary_interpolate


class TensorFunctionSpace:
    """Vector-valued space over a scalar Lagrange space (functionspace/tensor_space.py:13-102)."""

    def __init__(self, scalar_space: LagrangeFESpace, shape):
        self.scalar_space = scalar_space
        self.shape = tuple(shape)
        if len(self.shape) < 2:
            raise ValueError("shape must be a tuple of at least two element")
        if self.shape[0] == -1:
            self.dof_shape = tuple(self.shape[1:])
            self.dof_priority = False
        elif self.shape[-1] == -1:
            self.dof_shape = tuple(self.shape[:-1])
            self.dof_priority = True
        else:
            raise ValueError("`-1` is required as the first or last element of the shape")
        if len(self.dof_shape) != 1:
            raise NotImplementedError("only vector-valued tensor spaces (one dof axis) are on the accelerated path")
        self._c2d = None

    mesh = property(lambda s: s.scalar_space.mesh)
    device = property(lambda s: s.scalar_space.device)
    ftype = property(lambda s: s.scalar_space.ftype)
    itype = property(lambda s: s.scalar_space.itype)
    p = property(lambda s: s.scalar_space.p)

    @property
    def dof_numel(self):
        n = 1
        for v in self.dof_shape:
            n *= v
        return n

    @property
    def dof_ndim(self):
        return len(self.dof_shape)

    def number_of_global_dofs(self):
        return self.dof_numel * self.scalar_space.number_of_global_dofs()

    def number_of_local_dofs(self, doftype="cell"):
        return self.dof_numel * self.scalar_space.number_of_local_dofs(doftype)

    def cell_to_dof(self, index=None):
        """functionspace/utils.py:83-95"""
        if self._c2d is None:
            s = self.scalar_space
            c2d = s.cell_to_dof()
            NC, L = c2d.shape
            out = torch.empty((NC, L * self.dof_numel), dtype=torch.int32, device=self.device)
            _lib.call("fb2_tensor_cell_to_dof", _lib.ptr(c2d), NC, L, self.dof_numel, s.number_of_global_dofs(),
                      int(self.dof_priority), _lib.ptr(out), _lib.stream())
            self._c2d = out
        return self._c2d if index is None else self._c2d[index]

    def interpolation_points(self, index=None):
        return self.scalar_space.interpolation_points(index=index)

    def is_boundary_dof(self, threshold=None, method="interp"):
        """functionspace/tensor_space.py:159-188: the scalar flag repeated over the components (or one threshold per
        component when a tuple is given)"""
        s, n = self.scalar_space, self.dof_numel
        if isinstance(threshold, torch.Tensor):
            if threshold.dtype == torch.bool and threshold.numel() == self.number_of_global_dofs():
                return threshold
            raise ValueError("len(threshold) must equal tensorspace gdof")
        if threshold is None or callable(threshold):
            f = s.is_boundary_dof(threshold, method=method)
            return f.repeat(n) if self.dof_priority else f.repeat_interleave(n)
        if isinstance(threshold, tuple):
            assert n == len(threshold)
            fs = [s.is_boundary_dof(t, method=method) for t in threshold]
            return torch.cat(fs) if self.dof_priority else torch.stack(fs, dim=1).reshape(-1)
        raise ValueError(f"Unknown type of threshold {type(threshold)}")

    def boundary_interpolate(self, gd, uh=None, *, threshold=None, method=None):
        """functionspace/tensor_space.py:190-260 for the cases the path uses: gd a number, a full-size tensor, or a
        callable returning (npoints, ncomp) values at the boundary interpolation points"""
        s, n = self.scalar_space, self.dof_numel
        gdof = self.number_of_global_dofs()
        if uh is None:
            uh = torch.zeros(gdof, dtype=self.ftype, device=self.device)
        isbd = self.is_boundary_dof(threshold) if isinstance(threshold, torch.Tensor) else self.is_boundary_dof(threshold, method=method)
        if isinstance(gd, (int, float)):
            uh[isbd] = float(gd)          # (the reference writes uh[threshold] = gd here, :197-204, which only works for mask thresholds)
        elif isinstance(gd, torch.Tensor):
            assert gd.numel() == gdof
            uh[isbd] = gd.reshape(-1)[isbd]
        elif callable(gd):
            if isinstance(threshold, (tuple, torch.Tensor)):
                raise NotImplementedError("callable gd with per-component / tensor thresholds is not on the accelerated path")
            sflag = s.is_boundary_dof(threshold, method=method)
            val = gd(s.interpolation_points(index=sflag)[sflag])     # (nbd, ncomp)
            sg = s.number_of_global_dofs()
            view = uh.view(n, sg) if self.dof_priority else uh.view(sg, n)
            if self.dof_priority:
                view[:, sflag] = val.T.to(view.dtype)
            else:
                view[sflag, :] = val.to(view.dtype)
        else:
            raise ValueError("Unsupported type for gd. Must be a callable, int, float, or tensor.")
        return uh, isbd

    set_dirichlet_bc = boundary_interpolate
