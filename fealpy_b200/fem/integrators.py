"""Cell integrators of the accelerated path: K1 element matrices through the C ABI.

Same constructor signatures, attributes (`coef`, `q`, `material`, `method`) and protocol as the
reference: `assembly(space, indices=None) -> (NC, ldof, ldof)` and `to_global_dof(space)`
(fem/integrator.py:67-207, fem/scalar_diffusion_integrator.py:14-95,
fem/scalar_mass_integrator.py:13-54, fem/linear_elasticity_integrator.py:14-181).
Coefficient rules follow process_coef_func / bilinear_integral (utils/utils.py:21-53,
functional.py:68-106).
"""
from __future__ import annotations

import torch

from .. import _lib
from ..basis import device_tables, host_tables, number_of_local_dofs


class _Variant:
    """`integrator.assembly.set('fast')` plug-in point (decorator/variantmethod.py:64-106)."""

    def __init__(self, owner, table):
        self._owner, self._table, self._key = owner, table, None

    def set(self, key):
        if key not in self._table:
            raise NotImplementedError(f"variant {key!r} of {type(self._owner).__name__}.assembly is not on the "
                                      f"accelerated path (available: {list(self._table)})")
        self._key = key

    def get_key(self):
        return self._key

    def __call__(self, *a, **k):
        return self._table[self._key](*a, **k)


class Integrator:
    def __init__(self):
        self._region = None

    def set_region(self, region):
        if region is not None:
            raise NotImplementedError("region / sub-domain integration is not on the accelerated path")
        self._region = region

    def to_global_dof(self, space, indices=None):
        if indices is not None:
            raise NotImplementedError("chunked (indices=...) assembly is not needed on the GPU path")
        return space.cell_to_dof()

    def __call__(self, space, indices=None):
        return self.assembly(space, indices)


def _check_space(space):
    mesh = getattr(space, "mesh", None)
    if mesh is None or getattr(mesh, "TD", None) not in (2, 3):
        raise RuntimeError(f"{type(mesh).__name__} is not a simplex mesh of the accelerated path")
    if mesh.node.device.type != "cuda":
        raise RuntimeError("fealpy_b200 integrators need a mesh on a CUDA device; there is no CPU fallback")
    return mesh


def process_coef(coef, mesh, bcs, batched=False):
    """utils/utils.py:21-53 + the shape rules of functional.py:94-106.
    Returns (kind, payload): ('scalar', float) | ('cell', (NC,)) | ('quad', (NC,NQ)) | ('matrix', (NC,NQ,GD,GD))."""
    if batched:
        raise NotImplementedError("batched coefficients are not on the accelerated path")
    NC, NQ, GD = mesh.number_of_cells(), bcs.shape[0], mesh.geo_dimension()
    if callable(coef):
        if getattr(coef, "coordtype", "barycentric") == "barycentric":
            coef = coef(torch.as_tensor(bcs, dtype=torch.float64, device=mesh.device), index=slice(None))
        else:
            coef = coef(mesh.bc_to_point(bcs))
    if coef is None:
        return "scalar", 1.0
    if isinstance(coef, (int, float)):
        return "scalar", float(coef)
    if not isinstance(coef, torch.Tensor):
        raise TypeError(f"coef should be int, float or TensorLike, but got {type(coef)}.")
    if coef.numel() == 1:
        return "scalar", float(coef.reshape(()).item())
    coef = coef.to(device=mesh.device, dtype=torch.float64)
    if coef.ndim == 4:
        return "matrix", coef.expand(NC, NQ, GD, GD).contiguous()
    if coef.ndim > 4:
        raise RuntimeError(f"The dimension of the input should be smaller than 4, but got shape {tuple(coef.shape)}.")
    c = coef
    while c.ndim < 3:
        c = c[..., None]                       # fill_axis(coef, 3): (C,) -> (C,1,1), (C,Q) -> (C,Q,1)
    if c.shape[2] == 1:
        if c.shape[1] == 1:
            return "cell", c.expand(NC, 1, 1).reshape(NC).contiguous()
        return "quad", c.expand(NC, NQ, 1).reshape(NC, NQ).contiguous()
    # (C,Q,d): per-component coefficient == diagonal matrix coefficient
    return "matrix", torch.diag_embed(c.expand(NC, NQ, GD)).contiguous()


class _ScalarCellIntegrator(Integrator):
    KIND = None            # 'diffusion' | 'mass'

    def _q(self, space):
        return space.p + 3 if self.q is None else self.q

    def describe(self, space):
        """classification used by BilinearForm's fused path"""
        memo = getattr(self, "_describe_memo", None)      # set by BilinearForm.assembly() for the duration of one call:
        if memo is not None and memo[0] is space:         # a callable coefficient is evaluated once per assembly
            return memo[1]
        d = self._describe(space)
        if memo is not None:
            self._describe_memo = (space, d)
        return d

    def _describe(self, space):
        mesh = _check_space(space)
        q = self._q(space)
        tabs = device_tables(mesh.TD, space.p, q, mesh.device)
        if self.KIND == "diffusion" and self.assembly.get_key() == "fast":
            kind, payload = "scalar", 1.0      # the reference's 'fast' variant ignores coef (:65-79)
        else:
            # host copy of the quadrature points (cached numpy table): no device->host copy / sync on the assembly path
            kind, payload = process_coef(self.coef, mesh, host_tables(mesh.TD, space.p, q)["bcs"], self.batched)
        if kind == "matrix" and self.KIND == "mass":
            raise RuntimeError("matrix coefficients are not valid for the mass integrator")
        return dict(kind=self.KIND, q=q, coef_kind=kind, coef=payload, tabs=tabs)

    def _assembly_default(self, space, indices=None):
        if indices is not None:
            raise NotImplementedError("chunked (indices=...) assembly is not needed on the GPU path")
        d = self.describe(space)
        mesh = space.mesh
        TD, p, NC = mesh.TD, space.p, mesh.number_of_cells()
        L = number_of_local_dofs(TD, p)
        out = torch.empty((NC, L, L), dtype=torch.float64, device=mesh.device)
        tabs, ck, cf = d["tabs"], d["coef_kind"], d["coef"]
        is_mass = self.KIND == "mass"
        if ck in ("scalar", "cell"):
            scal = cf if ck == "scalar" else 1.0
            arr = cf if ck == "cell" else None
            if is_mass:
                _lib.call("fb2_elem_scalar_const", TD, p, NC, _lib.ptr(mesh.node), _lib.ptr(mesh.cell), None,
                          _lib.ptr(tabs["Mm"]), 0.0, None, scal, _lib.ptr(arr), _lib.ptr(out), _lib.stream())
            else:
                _lib.call("fb2_elem_scalar_const", TD, p, NC, _lib.ptr(mesh.node), _lib.ptr(mesh.cell), _lib.ptr(tabs["Ms"]),
                          None, scal, _lib.ptr(arr), 0.0, None, _lib.ptr(out), _lib.stream())
        else:
            table = tabs["phi"] if is_mass else tabs["R"]
            _lib.call("fb2_elem_scalar_quad", TD, p, NC, _lib.ptr(mesh.node), _lib.ptr(mesh.cell), int(is_mass),
                      tabs["ws"].shape[0], _lib.ptr(tabs["ws"]), _lib.ptr(table), 2 if ck == "quad" else 3, _lib.ptr(cf),
                      _lib.ptr(out), _lib.stream())
        return out


class ScalarDiffusionIntegrator(_ScalarCellIntegrator):
    """(kappa grad u, grad v); fem/scalar_diffusion_integrator.py:14-95"""
    KIND = "diffusion"

    def __init__(self, coef=None, q=None, *, region=None, batched=False, method=None):
        super().__init__()
        self.coef, self.q, self.batched = coef, q, batched
        self.set_region(region)
        self.assembly = _Variant(self, {None: self._assembly_default, "fast": self._assembly_default})
        self.assembly.set(method)


class ScalarMassIntegrator(_ScalarCellIntegrator):
    """(c u, v); fem/scalar_mass_integrator.py:13-54"""
    KIND = "mass"

    def __init__(self, coef=None, q=None, *, index=slice(None), batched=False, method=None):
        super().__init__()
        if index != slice(None):
            raise NotImplementedError("index= sub-selection is not on the accelerated path")
        self.coef, self.q, self.index, self.batched = coef, q, index, batched
        self.assembly = _Variant(self, {None: self._assembly_default})
        self.assembly.set(method)


class LinearElasticityIntegrator(Integrator):
    """isotropic linear elasticity on a TensorFunctionSpace; fem/linear_elasticity_integrator.py:14-181
    (the reference ignores its ctor `method`, :24 -- so does this class)"""
    KIND = "elasticity"

    def __init__(self, material, q=None, *, index=slice(None), method=None):
        super().__init__()
        if index != slice(None):
            raise NotImplementedError("index= sub-selection is not on the accelerated path")
        self.material, self.q, self.index = material, q, index
        self.assembly = _Variant(self, {None: self._assembly_default})

    def coefficients(self, space):
        D = getattr(self.material, "_D", None)          # host copy: no device round trip on the assembly path
        if D is None:
            D = self.material.elastic_matrix()[0, 0].cpu()
        GD = space.mesh.geo_dimension()
        if GD == 2:
            D00, D01, Dss = float(D[0, 0]), float(D[0, 1]), float(D[2, 2])
            d_diag = D00
        else:
            D00, D01, Dss = float(D[0, 0]), float(D[0, 1]), float(D[5, 5])
            # the interleaved 3-D branch uses (2*D55 + D01) on the diagonal blocks (:160-165)
            d_diag = D00 if space.dof_priority else (2 * Dss + D01)
        return d_diag, D01, Dss

    def _assembly_default(self, space, indices=None):
        if indices is not None:
            raise NotImplementedError("chunked (indices=...) assembly is not needed on the GPU path")
        sspace = getattr(space, "scalar_space", None)
        if sspace is None:
            raise RuntimeError("LinearElasticityIntegrator needs a TensorFunctionSpace")
        mesh = _check_space(sspace)
        TD, p, NC, GD = mesh.TD, sspace.p, mesh.number_of_cells(), mesh.geo_dimension()
        if space.dof_numel != GD:
            raise ValueError("the tensor space must have GD components")
        q = p + 3 if self.q is None else self.q
        tabs = device_tables(TD, p, q, mesh.device)
        L = number_of_local_dofs(TD, p)
        d_diag, d_lam, d_shear = self.coefficients(space)
        out = torch.empty((NC, GD * L, GD * L), dtype=torch.float64, device=mesh.device)
        _lib.call("fb2_elem_elasticity", TD, p, NC, _lib.ptr(mesh.node), _lib.ptr(mesh.cell), _lib.ptr(tabs["M4"]), d_diag, d_lam,
                  d_shear, int(space.dof_priority), _lib.ptr(out), _lib.stream())
        return out
