"""BilinearForm: global assembly of cell integrators into a CSRTensor on the GPU.

Drop-in for the reference call surface (fem/form.py:39-188, fem/bilinear_form.py:11-158):
    BilinearForm(space).add_integrator(I, ...).assembly(format='csr') -> CSRTensor
Integrators added in ONE add_integrator call form a group (their element matrices are
summed, fem/integrator.py:295-378); separate calls are separate groups whose COO blocks the
reference concatenates and sums in coalesce().  Three device paths produce the same matrix:

  'fused'  symbolic pattern (cached per space) + row-owner numeric kernel that recomputes the
           element rows in registers (constant / per-cell coefficients, scalar spaces)
  'gather' symbolic pattern + K1 element matrices + row-owner gather (any integrator,
           variable coefficients, tensor spaces)
  'coo'    K1 + the literal COO -> stable (row, col) sort -> ordered segmented reduce (K2)

`assembly_path='auto'` picks fused, else gather.  The pattern (explicit zeros included) is
identical for all three and bit-identical to the reference's coalesce().tocsr().
"""
from __future__ import annotations

import ctypes as C
import os as _os
from collections.abc import Sequence

import torch

from .. import _lib
from ..basis import host_tables
from ..sparse import CSRTensor
from .integrators import Integrator


class GroupIntegrator(Integrator):
    """sum of the member element matrices, in list order (fem/integrator.py:295-378)"""

    def __init__(self, *ints, region=None):
        super().__init__()
        if len(ints) == 0:
            raise ValueError("No integrators provided.")
        self.ints = []
        for it in ints:
            if isinstance(it, GroupIntegrator):
                self.ints.extend(it.ints)
            elif hasattr(it, "assembly"):
                self.ints.append(it)
            else:
                raise TypeError(f"Unsupported type {it.__class__.__name__} found in the inputs.")
        self.set_region(region)

    def __iter__(self):
        return iter(self.ints)

    def __len__(self):
        return len(self.ints)

    def __iadd__(self, other):
        if isinstance(other, GroupIntegrator):
            self.ints.extend(other.ints)
        else:
            self.ints.append(other)
        return self

    def to_global_dof(self, space, indices=None):
        return self.ints[0].to_global_dof(space)

    def assembly(self, space, indices=None):
        ct = self.ints[0].assembly(space)
        for it in self.ints[1:]:
            ct = ct + it.assembly(space)
        return ct


def _bits(n: int) -> int:
    b = 1
    while (1 << b) < n:
        b += 1
    return b


def symbolic_pattern(space):
    """cached per scalar space: adjacency, CSR pattern and slot map (csrc/assemble.cu)"""
    cache = getattr(space, "_b200_symbolic", None)
    if cache is not None:
        return cache
    lib = _lib.load()
    c2d = space.cell_to_dof().contiguous()
    NC, L = c2d.shape
    gdof = space.number_of_global_dofs()
    dev = c2d.device
    adj_ptr = torch.empty(gdof + 1, dtype=torch.int64, device=dev)
    adj_pair = torch.empty(NC * L, dtype=torch.int32, device=dev)
    crow = torch.empty(gdof + 1, dtype=torch.int64, device=dev)
    ws = _lib.workspace(lib.fb2_sym_workspace_bytes(NC, L, gdof), dev)
    nnz, max_row = C.c_int64(0), C.c_int32(0)
    try:        # scratch for the candidates' ranks (2 bytes per (cell, i, j)): the fill pass then need not rank them again.
        # Sized so that asm4_plan() can take it over as its schedule buffer (~21 bytes per (cell, i)): a multi-GB cudaMalloc
        # costs ~1.6 ms per GB on a fresh process, which is most of what a cold assembly pays on top of its kernels
        nwords = lib.fb2_slot_stride(L, 1) // 4
        est_blocks = int(NC * L / 32 / 0.80 + 1024) * (64 + 32 * nwords + 4) * 4
        stash_bytes = max(2 * NC * L * L, est_blocks if L <= 20 else 0)
        stash = None if _os.environ.get("FB2_SYM_NO_STASH") else torch.empty(stash_bytes, dtype=torch.uint8, device=dev)
    except torch.cuda.OutOfMemoryError:
        stash = None
    _lib.call("fb2_sym_count", _lib.ptr(c2d), NC, L, gdof, _lib.ptr(adj_ptr), _lib.ptr(adj_pair), _lib.ptr(crow),
              C.byref(nnz), C.byref(max_row), _lib.ptr(stash), _lib.ptr(ws), _lib.stream())
    slot_bytes = 1 if max_row.value <= 255 else 2
    col = torch.empty(nnz.value, dtype=torch.int32, device=dev)
    stride = lib.fb2_slot_stride(L, slot_bytes)
    slots = torch.zeros(NC * L * stride, dtype=torch.uint8 if slot_bytes == 1 else torch.int16, device=dev)
    _lib.call("fb2_sym_fill", _lib.ptr(c2d), NC, L, gdof, _lib.ptr(adj_ptr), _lib.ptr(adj_pair), _lib.ptr(crow), _lib.ptr(col),
              _lib.ptr(slots), slot_bytes, _lib.ptr(stash), _lib.stream())
    # CTA tile of the gather kernels: where a dof meets few cells (tri P3: 2.2 (cell, i) pairs per row) the gather is bound by
    # latency and wants more resident CTAs -- 2048-value tiles: 4.03 against 4.76 ms on config 3; the high-valence tensor
    # gather of config 4 wants the large tile (5.36 against 6.13 ms).  profiles/r02_tune_gather.txt
    tile = ASM_TILE if ("FB2_ASM_TILE" in _os.environ or NC * L >= 4 * gdof) else 2048
    blk_row, nblk = row_tiling(crow, gdof, nnz.value, tile)
    cache = dict(adj_ptr=adj_ptr, adj_pair=adj_pair, crow=crow, col=col, slots=slots, slot_bytes=slot_bytes,
                 max_row=max_row.value, nnz=nnz.value, NC=NC, L=L, gdof=gdof, blk_row=blk_row, nblk=nblk, tile=tile,
                 scratch=stash)          # handed to asm4_plan(), dropped by the first kernel that does not want it
    space._b200_symbolic = cache
    return cache


def adjacency(space):
    """dof -> (cell, local index) adjacency of a scalar space (cached): the part of the symbolic phase a matrix-free
    product needs; the full pattern's copy is used when it already exists"""
    sym = getattr(space, "_b200_symbolic", None)
    if sym is not None:
        return sym
    adj = getattr(space, "_b200_adjacency", None)
    if adj is not None:
        return adj
    lib = _lib.load()
    c2d = space.cell_to_dof().contiguous()
    NC, L = c2d.shape
    gdof = space.number_of_global_dofs()
    adj_ptr = torch.empty(gdof + 1, dtype=torch.int64, device=c2d.device)
    adj_pair = torch.empty(NC * L, dtype=torch.int32, device=c2d.device)
    ws = _lib.workspace(lib.fb2_adjacency_workspace_bytes(gdof), c2d.device)
    _lib.call("fb2_adjacency", _lib.ptr(c2d), NC, L, gdof, _lib.ptr(adj_ptr), _lib.ptr(adj_pair), _lib.ptr(ws), _lib.stream())
    adj = space._b200_adjacency = dict(adj_ptr=adj_ptr, adj_pair=adj_pair, NC=NC, L=L, gdof=gdof)
    return adj


import os as _os
ASM_TILE = int(_os.environ.get("FB2_ASM_TILE", "4096"))      # CSR values per CTA of the numeric assembly kernels


def row_tiling(crow, nrow, nnz, tile):
    """first row of every `tile`-value CTA tile (csrc/cg.cu partition_rows_kernel)"""
    lib = _lib.load()
    nblk = lib.fb2_spmv_plan_blocks(nnz, tile)
    blk_row = torch.empty(nblk + 2, dtype=torch.int32, device=crow.device)
    mr = C.c_int32(0)
    _lib.call("fb2_spmv_plan_build", nrow, _lib.ptr(crow), tile, _lib.ptr(blk_row), nnz, C.byref(mr), _lib.stream())
    return blk_row, nblk


ASM4_TILE = int(_os.environ.get("FB2_ASM4_TILE", "2560"))     # values per WARP tile of the v4 kernel


def asm4_plan(space):
    """batch schedule of the v4 numeric kernel (cached per space; csrc/assemble.cu asm4_schedule_kernel)"""
    sym = symbolic_pattern(space)
    if "asm4" in sym:
        return sym["asm4"]
    lib = _lib.load()
    dev = sym["crow"].device
    blk_row, ntile = row_tiling(sym["crow"], sym["gdof"], sym["nnz"], ASM4_TILE)
    ws = _lib.workspace(lib.fb2_asm4_workspace_bytes(ntile), dev)
    batch_ptr = torch.empty(ntile + 1, dtype=torch.int64, device=dev)
    nb = C.c_int64(0)
    _lib.call("fb2_asm4_plan_count", ntile, _lib.ptr(blk_row), _lib.ptr(sym["crow"]), _lib.ptr(sym["adj_ptr"]), _lib.ptr(sym["adj_pair"]),
              sym["L"], _lib.ptr(batch_ptr), C.byref(nb), _lib.ptr(ws), _lib.stream())
    nb = nb.value
    nwords = lib.fb2_slot_stride(sym["L"], sym["slot_bytes"]) * sym["slot_bytes"] // 4
    batch_i = torch.empty(nb, dtype=torch.uint8, device=dev)
    # one packed block per batch (include/fealpy_b200.h): the numeric kernel fetches it with a single bulk copy
    bw = 64 + 32 * nwords + 4
    scratch = sym.pop("scratch", None)
    if scratch is not None and scratch.numel() >= nb * bw * 4:
        blocks = scratch[:nb * bw * 4].view(torch.int32).view(nb, bw).zero_()      # the symbolic phase's rank scratch, recycled
    else:
        blocks = torch.zeros((nb, bw), dtype=torch.int32, device=dev)
    del scratch
    _lib.call("fb2_asm4_plan_fill", ntile, _lib.ptr(blk_row), _lib.ptr(sym["crow"]), _lib.ptr(sym["adj_ptr"]), _lib.ptr(sym["adj_pair"]),
              sym["L"], _lib.ptr(batch_ptr), _lib.ptr(batch_i), _lib.ptr(blocks), _lib.ptr(sym["slots"]), sym["slot_bytes"], _lib.stream())
    sym["asm4"] = dict(blk_row=blk_row, ntile=ntile, tile=ASM4_TILE, batch_ptr=batch_ptr, batch_i=batch_i, blocks=blocks,
                       ent_cell=blocks[:, :32], ent_base=blocks[:, 32:64], ent_slots=blocks[:, 64:64 + 32 * nwords],
                       header_i=blocks[:, 64 + 32 * nwords], nbatch=nb)
    return sym["asm4"]


def tensor_pattern(space):
    """pattern of a TensorFunctionSpace = scalar pattern (x) dense ncomp x ncomp blocks"""
    cache = getattr(space, "_b200_pattern", None)
    if cache is not None:
        return cache
    sym = symbolic_pattern(space.scalar_space)
    nc = space.dof_numel
    dev = sym["crow"].device
    crow = torch.empty(sym["gdof"] * nc + 1, dtype=torch.int64, device=dev)
    col = torch.empty(sym["nnz"] * nc * nc, dtype=torch.int32, device=dev)
    _lib.call("fb2_expand_pattern", sym["gdof"], nc, int(space.dof_priority), _lib.ptr(sym["crow"]), _lib.ptr(sym["col"]),
              _lib.ptr(crow), _lib.ptr(col), _lib.stream())
    blk_row, nblk = row_tiling(crow, sym["gdof"] * nc, sym["nnz"] * nc * nc, ASM_TILE)
    cache = dict(crow=crow, col=col, blk_row=blk_row, nblk=nblk, tile=ASM_TILE)
    space._b200_pattern = cache
    return cache


class BilinearForm:
    def __init__(self, space, batch_size: int = 0, *, assembly_path: str = "auto", share_pattern: bool = True):
        if isinstance(space, (tuple, list)):
            if len(space) != 1:
                raise NotImplementedError("two-space (rectangular) forms are not on the accelerated path")
            space = space[0]
        if batch_size:
            raise NotImplementedError("batched forms (batch_size > 0) are not on the accelerated path")
        self.space = space
        self._spaces = (space,)
        self.batch_size = 0
        self.integrators = {}
        self.splitters = {}
        self._cursor = 0
        self._M = None
        self._transposed = False
        if assembly_path not in ("auto", "fused", "gather", "coo"):
            raise ValueError(f"unknown assembly_path {assembly_path!r}")
        self.assembly_path = assembly_path
        self.last_path = None
        # The CSR pattern depends only on the space and is cached with it.  share_pattern=True (default): every matrix
        # assembled on this space references the SAME crow/col tensors -- treat them as read-only (CSRTensor.copy()
        # gives private ones).  share_pattern=False: each returned matrix owns fresh copies, as the reference's do.
        self.share_pattern = bool(share_pattern)

    # ---- bookkeeping (fem/form.py:92-144) ----------------------------------------------------
    @property
    def shape(self):
        g = self.space.number_of_global_dofs()
        return (g, g)

    sparse_shape = shape

    def add_integrator(self, *I, region=None, splitter=None, group=None):
        if len(I) == 0:
            return self
        if len(I) == 1 and isinstance(I[0], Sequence):
            I = tuple(I[0])
        if region is not None:
            raise NotImplementedError("region / sub-domain integration is not on the accelerated path")
        I = I[0] if len(I) == 1 else GroupIntegrator(*I)
        # `splitter` chunks the element loop to bound the reference's memory; the GPU path never
        # materialises gphi, so it is accepted and ignored.
        return self._add_integrator_impl(I, group, splitter)

    def __lshift__(self, other):
        if hasattr(other, "assembly"):
            return self._add_integrator_impl(other, None)
        return NotImplemented

    def _add_integrator_impl(self, I, group=None, splitter=None):
        group = f"_group_{self._cursor}" if group is None else group
        self._cursor += 1
        if group in self.integrators:
            cur = self.integrators[group]
            if not isinstance(cur, GroupIntegrator):
                cur = GroupIntegrator(cur)
            cur += I
            self.integrators[group] = cur
        else:
            self.integrators[group] = I
        self.splitters[group] = splitter
        self._M = None
        return self

    def _flat_integrators(self):
        out = []
        for it in self.integrators.values():
            out.extend(it.ints if isinstance(it, GroupIntegrator) else [it])
        return out

    def assembly_local_iterative(self):
        for it in self.integrators.values():
            etg = it.to_global_dof(self.space)
            yield it.assembly(self.space), (etg,)

    # ---- assembly ------------------------------------------------------------------------------
    def _is_tensor_space(self):
        return hasattr(self.space, "scalar_space")

    def _plan_fused(self):
        """merge constant / per-cell scalar integrators into (<=1 diffusion, <=1 mass) kernels"""
        if self._is_tensor_space():
            return None
        merged = {}
        for it in self._flat_integrators():
            desc = getattr(it, "describe", None)
            if desc is None:
                return None
            d = desc(self.space)
            if d["coef_kind"] not in ("scalar", "cell"):
                return None
            m = merged.setdefault(d["kind"], dict(q=d["q"], tabs=d["tabs"], scal=0.0, arr=None))
            if m["q"] != d["q"]:
                return None
            if d["coef_kind"] == "scalar":
                if m["arr"] is None:
                    m["scal"] += d["coef"]
                else:
                    m["arr"] = m["arr"] + d["coef"]
            else:
                m["arr"] = d["coef"] + (m["scal"] if m["arr"] is None else m["arr"])
                m["scal"] = 0.0
        return merged

    @staticmethod
    def _plan_parts(m):
        """(scalar factor, per-cell coefficient or None) of one merged term of _plan_fused()"""
        if m is None:
            return 0.0, None
        return (1.0, m["arr"].contiguous()) if m["arr"] is not None else (m["scal"], None)

    def _host_table(self, m, key):
        if m is None:
            return None
        h = host_tables(self.space.mesh.TD, self.space.p, m["q"])[key]
        return h.ctypes.data_as(C.c_void_p)

    def _values_buffer(self, nnz, device, out):
        if out is None:
            return torch.empty(nnz, dtype=torch.float64, device=device)
        if not (isinstance(out, torch.Tensor) and out.dtype == torch.float64 and out.is_contiguous()
                and out.shape == (nnz,) and out.device == torch.device(device)):
            raise ValueError(f"out must be a contiguous float64 tensor of shape ({nnz},) on {device}")
        return out

    def _assemble_fused(self, plan, out=None):
        space, mesh = self.space, self.space.mesh
        sym = symbolic_pattern(space)
        values = self._values_buffer(sym["nnz"], mesh.device, out)
        dm, mm = plan.get("diffusion"), plan.get("mass")
        sd, ad = self._plan_parts(dm)
        sm_, am = self._plan_parts(mm)
        hostp = self._host_table
        kernel = _os.environ.get("FB2_ASM_KERNEL", "auto")
        if kernel == "auto":
            kernel = "v4"
        if kernel == "v4" and (ASM4_TILE + sym["max_row"] >= 4096 or sym["max_row"] > 1024 or sym["L"] > 20):
            kernel = "v2"          # very long rows (high-valence meshes): tile offsets / first-touch bitmap of v4 do not cover them
        NH = mesh.TD * (mesh.TD + 1) // 2 + 1          # reduced geometry record (csrc/assemble.cu A4Geo)
        self.last_kernel = kernel
        if kernel == "v4":
            pl = asm4_plan(space)
            geom = getattr(self, "_asm4_geom", None)      # per-cell geometry scratch, kept with the form (no allocator churn per assembly)
            if geom is None or geom.shape[0] != sym["NC"] or geom.device != mesh.device:
                geom = self._asm4_geom = torch.empty((sym["NC"], (NH + 1) // 2 * 2), dtype=torch.float64, device=mesh.device)
            _lib.call("fb2_assemble_scalar_const_v4", mesh.TD, space.p, sym["NC"], _lib.ptr(mesh.node), _lib.ptr(mesh.cell),
                      _lib.ptr(sym["crow"]), _lib.ptr(pl["blk_row"]), pl["ntile"], pl["tile"], sym["max_row"], _lib.ptr(pl["batch_ptr"]),
                      _lib.ptr(pl["batch_i"]), _lib.ptr(pl["blocks"]), sym["slot_bytes"], hostp(dm, "Ms"), hostp(mm, "Mm"),
                      sd, _lib.ptr(ad), sm_, _lib.ptr(am), _lib.ptr(geom), _lib.ptr(values), _lib.stream())
            return sym["crow"], sym["col"], values
        _lib.call("fb2_assemble_scalar_const", mesh.TD, space.p, sym["NC"], sym["gdof"], _lib.ptr(mesh.node), _lib.ptr(mesh.cell),
                  _lib.ptr(sym["adj_ptr"]), _lib.ptr(sym["adj_pair"]), _lib.ptr(sym["slots"]), sym["slot_bytes"],
                  _lib.ptr(sym["crow"]), sym["max_row"], _lib.ptr(sym["blk_row"]), sym["nblk"], sym["tile"],
                  _lib.ptr(dm["tabs"]["Ms"]) if dm else None, _lib.ptr(mm["tabs"]["Mm"]) if mm else None,
                  sd, _lib.ptr(ad), sm_, _lib.ptr(am), _lib.ptr(values), _lib.stream())
        return sym["crow"], sym["col"], values

    def _summed_ke_scalar(self):
        """sum of several scalar diffusion / mass integrators written ONCE into one (NC, l, l) block: constant and per-cell
        terms are merged per (kind, q) and folded into the first quadrature-loop kernel's final write, every further kernel
        accumulates in place (csrc/elem.cu) -- no per-integrator block, no torch add.  None: not applicable."""
        if self._is_tensor_space():
            return None
        its = self._flat_integrators()
        if len(its) < 2 or not all(hasattr(it, "describe") and hasattr(it, "KIND") for it in its):
            return None
        space, mesh = self.space, self.space.mesh
        descs = [it.describe(space) for it in its]
        const, quad = {"diffusion": {}, "mass": {}}, []
        for d in descs:
            if d["coef_kind"] in ("scalar", "cell"):
                m = const[d["kind"]].setdefault(d["q"], dict(tabs=d["tabs"], scal=0.0, arr=None))
                if d["coef_kind"] == "scalar":
                    if m["arr"] is None:
                        m["scal"] += d["coef"]
                    else:
                        m["arr"] = m["arr"] + d["coef"]
                else:
                    m["arr"] = d["coef"] + (m["scal"] if m["arr"] is None else m["arr"])
                    m["scal"] = 0.0
            else:
                quad.append(d)
        dts, mts = list(const["diffusion"].values()), list(const["mass"].values())
        bundles = [(dts[k] if k < len(dts) else None, mts[k] if k < len(mts) else None) for k in range(max(len(dts), len(mts)))]
        TD, p, NC = mesh.TD, space.p, mesh.number_of_cells()
        L = space.number_of_local_dofs()
        out = torch.empty((NC, L, L), dtype=torch.float64, device=mesh.device)

        def cargs(b):
            dm, mm = b if b is not None else (None, None)
            sd, ad = self._plan_parts(dm)
            sm_, am = self._plan_parts(mm)
            return (_lib.ptr(dm["tabs"]["Ms"]) if dm else None, _lib.ptr(mm["tabs"]["Mm"]) if mm else None, sd, _lib.ptr(ad), sm_, _lib.ptr(am))
        first = True
        for d in quad:
            is_mass = d["kind"] == "mass"
            tabs = d["tabs"]
            fold = bundles.pop(0) if (first and bundles) else None
            _lib.call("fb2_elem_scalar_quad_fused", TD, p, NC, _lib.ptr(mesh.node), _lib.ptr(mesh.cell), int(is_mass), tabs["ws"].shape[0],
                      _lib.ptr(tabs["ws"]), _lib.ptr(tabs["phi"] if is_mass else tabs["R"]), 2 if d["coef_kind"] == "quad" else 3,
                      _lib.ptr(d["coef"]), *cargs(fold), _lib.ptr(out), 0 if first else 1, _lib.stream())
            first = False
        for b in bundles:
            ms, mm, sd, ad, sm_, am = cargs(b)
            _lib.call("fb2_elem_scalar_const_acc", TD, p, NC, _lib.ptr(mesh.node), _lib.ptr(mesh.cell), ms, mm, sd, ad, sm_, am,
                      _lib.ptr(out), 0 if first else 1, _lib.stream())
            first = False
        return out

    def _summed_ke(self):
        ke = self._summed_ke_scalar()
        if ke is not None:
            return ke
        for it in self.integrators.values():
            k = it.assembly(self.space)
            if not isinstance(k, torch.Tensor) or k.ndim != 3:
                raise ValueError("Output of operator integrators should be 3D, "
                                 f"but got shape {tuple(getattr(k, 'shape', ()))}.")
            ke = k if ke is None else ke + k          # out of place: an integrator may return a cached block
        return ke.contiguous()

    def _plan_elasticity_p1(self):
        """one LinearElasticityIntegrator on a P1 tensor space: grad(phi) is cell-constant, the element matrix is a
        closed form of grad(lambda) and the gather computes its entries on the fly (no K_e block in HBM)"""
        its = self._flat_integrators()
        if not self._is_tensor_space() or len(its) != 1 or getattr(its[0], "KIND", None) != "elasticity":
            return None
        space, it = self.space, its[0]
        sspace = space.scalar_space
        mesh = sspace.mesh
        if getattr(sspace, "p", None) != 1 or space.dof_numel != mesh.geo_dimension() or mesh.TD not in (2, 3):
            return None
        q = sspace.p + 3 if it.q is None else it.q
        wsum = float(host_tables(mesh.TD, 1, q)["M4"][0, 0, 0, 0])
        return dict(coef=it.coefficients(space), wsum=wsum)

    def _assemble_elasticity_p1(self, plan, out=None):
        space = self.space
        sspace, mesh = space.scalar_space, space.scalar_space.mesh
        sym, pat = symbolic_pattern(sspace), tensor_pattern(space)
        TD, NC = mesh.TD, sym["NC"]
        d_diag, d_lam, d_shear = plan["coef"]
        geo = getattr(self, "_ep1_geo", None)            # per-cell gradient records, kept with the form (no allocator churn)
        if geo is None or geo.shape[0] != NC or geo.device != mesh.device:
            geo = self._ep1_geo = torch.empty((NC, (TD + 1) * TD + 1), dtype=torch.float64, device=mesh.device)
        values = self._values_buffer(pat["col"].shape[0], mesh.device, out)
        _lib.call("fb2_assemble_elasticity_p1", TD, NC, _lib.ptr(mesh.node), _lib.ptr(mesh.cell), int(space.dof_priority), sym["gdof"],
                  d_diag, d_lam, d_shear, plan["wsum"], _lib.ptr(sym["adj_ptr"]), _lib.ptr(sym["adj_pair"]), _lib.ptr(sym["slots"]),
                  sym["slot_bytes"], _lib.ptr(sym["crow"]), sym["max_row"], _lib.ptr(pat["crow"]), _lib.ptr(pat["blk_row"]), pat["nblk"],
                  pat["tile"], _lib.ptr(geo), _lib.ptr(values), _lib.stream())
        return pat["crow"], pat["col"], values

    def _assemble_gather(self, out=None):
        space = self.space
        ke = self._summed_ke()
        if self._is_tensor_space():
            sspace, nc, prio = space.scalar_space, space.dof_numel, int(space.dof_priority)
            sym = symbolic_pattern(sspace)
            pat = tensor_pattern(space)
            crow, col, tiling = pat["crow"], pat["col"], pat
        else:
            sym = symbolic_pattern(space)
            nc, prio = 1, 0
            crow, col, tiling = sym["crow"], sym["col"], sym
        if ke.shape != (sym["NC"], sym["L"] * nc, sym["L"] * nc):
            raise ValueError(f"entity_to_global.shape[0] != local_tensor.shape[0] or wrong local shape {tuple(ke.shape)}")
        values = self._values_buffer(col.shape[0], ke.device, out)
        _lib.call("fb2_assemble_from_ke", sym["NC"], sym["L"], nc, prio, sym["gdof"], _lib.ptr(ke), _lib.ptr(sym["adj_ptr"]),
                  _lib.ptr(sym["adj_pair"]), _lib.ptr(sym["slots"]), sym["slot_bytes"], _lib.ptr(sym["crow"]), sym["max_row"],
                  _lib.ptr(crow), _lib.ptr(tiling["blk_row"]), tiling["nblk"], tiling["tile"], _lib.ptr(values), _lib.stream())
        return crow, col, values

    def _assemble_coo(self):
        """the literal reference pipeline on the device: COO of every group -> sort -> reduce"""
        lib = _lib.load()
        space = self.space
        gdof = space.number_of_global_dofs()
        blocks = [(it.assembly(space), it.to_global_dof(space).contiguous()) for it in self.integrators.values()]
        dev = blocks[0][0].device
        n = sum(k.numel() for k, _ in blocks)
        cbits = _bits(max(gdof, 2))
        keys = torch.empty(n, dtype=torch.int64, device=dev)
        vals = torch.empty(n, dtype=torch.float64, device=dev)
        off = 0
        for k, c2d in blocks:
            NC, lr, lc = k.shape
            if c2d.shape != (NC, lr) or lr != lc:
                raise ValueError("entity_to_global.shape[0] != local_tensor.shape[0]")
            m = k.numel()
            _lib.call("fb2_coo_keys_from_c2d", _lib.ptr(c2d), _lib.ptr(c2d), NC, lr, lc, cbits,
                      C.c_void_p(keys.data_ptr() + 8 * off), _lib.stream())
            vals[off:off + m] = k.reshape(-1)
            off += m
        perm = torch.empty(n, dtype=torch.int32, device=dev)
        ws = _lib.workspace(lib.fb2_coo_workspace_bytes(n), dev)
        nnz = C.c_int64(0)
        _lib.call("fb2_coo_symbolic", _lib.ptr(keys), _lib.ptr(perm), n, 2 * cbits, _lib.ptr(ws), C.byref(nnz), _lib.stream())
        nnz = nnz.value
        crow = torch.empty(gdof + 1, dtype=torch.int64, device=dev)
        col = torch.empty(nnz, dtype=torch.int32, device=dev)
        seg = torch.empty(nnz + 1, dtype=torch.int64, device=dev)
        _lib.call("fb2_coo_fill", _lib.ptr(keys), n, cbits, gdof, _lib.ptr(ws), _lib.ptr(crow), _lib.ptr(col), 4, _lib.ptr(seg),
                  _lib.stream())
        values = torch.empty(nnz, dtype=torch.float64, device=dev)
        _lib.call("fb2_coo_reduce", _lib.ptr(perm), _lib.ptr(seg), nnz, _lib.ptr(vals), _lib.ptr(values), _lib.stream())
        return crow, col, values

    def assembly(self, *, format="csr", out=None):
        """fem/bilinear_form.py:83-105.  `out` (optional, beyond the reference signature): a float64 tensor of nnz entries
        that receives the CSR values, so that repeated assemblies allocate nothing."""
        if format not in ("csr", "coo"):
            raise ValueError(f"Unsupported format {format}.")
        if not self.integrators:
            raise ValueError("no integrators added")
        flat = self._flat_integrators()
        for it in flat:                      # classification (incl. the evaluation of callable coefficients) happens once
            if hasattr(it, "describe"):      # per assembly(), whichever path ends up using it
                it._describe_memo = (None, None)
        try:
            return self._assembly_impl(format, out)
        finally:
            for it in flat:
                if hasattr(it, "describe"):
                    it._describe_memo = None

    def _assembly_impl(self, format, out):
        path = self.assembly_path
        plan = None
        if path in ("auto", "fused"):
            plan = self._plan_fused()
            if plan is None and path == "fused" and self._plan_elasticity_p1() is None:
                raise NotImplementedError("the fused path needs constant / per-cell scalar coefficients on a scalar space")
        eplan = self._plan_elasticity_p1() if (plan is None and path in ("auto", "fused")) else None
        if plan is not None:
            crow, col, values = self._assemble_fused(plan, out)
            self.last_path = "fused"
        elif eplan is not None:
            crow, col, values = self._assemble_elasticity_p1(eplan, out)
            self.last_path = "fused-elasticity-p1"
        elif path == "coo":
            crow, col, values = self._assemble_coo()
            if out is not None:
                values = self._values_buffer(values.shape[0], values.device, out).copy_(values)
            self.last_path = "coo"
        else:
            crow, col, values = self._assemble_gather(out)
            self.last_path = "gather"
        if self.last_path != "coo":          # the symbolic phase's rank scratch: only the v4 schedule recycles it
            sp = self.space.scalar_space if self._is_tensor_space() else self.space
            getattr(sp, "_b200_symbolic", {}).pop("scratch", None)
        if not self.share_pattern and self.last_path != "coo":
            crow, col = crow.clone(), col.clone()
        M = CSRTensor(crow, col, values, self.shape)
        if self._transposed:
            raise NotImplementedError("transposed forms are not on the accelerated path")
        self._M = M if format == "csr" else M.tocoo()
        return self._M

    def __matmul__(self, u):
        """assembled product if assembly() was called, else matrix-free (fem/bilinear_form.py:126-158)"""
        if self._M is not None:
            return self._M @ u
        if not isinstance(u, torch.Tensor) or u.ndim != 1 or u.dtype != torch.float64:
            raise NotImplementedError("matrix-free products take a 1-D float64 CUDA tensor")
        sym = adjacency(self.space)
        if u.shape[0] != sym["gdof"]:
            raise ValueError("shape mismatch")
        v = torch.empty_like(u)
        plan = self._plan_fused() if (self.assembly_path in ("auto", "fused") and not self._is_tensor_space()) else None
        if plan is not None:
            # constant / per-cell coefficients: neither A nor K_e is formed (csrc/assemble.cu matfree_cell_kernel)
            space, mesh = self.space, self.space.mesh
            dm, mm = plan.get("diffusion"), plan.get("mass")
            sd, ad = self._plan_parts(dm)
            sm_, am = self._plan_parts(mm)
            ws = getattr(self, "_matfree_ws", None)       # (NC, l) per-cell products, kept with the form
            if ws is None or ws.shape != (sym["NC"], sym["L"]) or ws.device != u.device:
                ws = self._matfree_ws = torch.empty((sym["NC"], sym["L"]), dtype=torch.float64, device=u.device)
            # cell products in adjacency order (contiguous per-dof sum: 1.29 against 1.79 ms at tet P2 128^3, faster than the
            # assembled SpMV's 1.35 ms) once the block outgrows L2; a block that stays in L2 is gathered in pair order
            # (config 1: 0.111 against 0.122 ms).  pair_pos = inverse of adj_pair, built once per space.
            order = _os.environ.get("FB2_MATFREE_ORDER", "adj" if sym["NC"] * sym["L"] * 8 > (64 << 20) else "pair")
            pos = sym.get("pair_pos") if order == "adj" else None
            if pos is None and order == "adj":
                pos = torch.empty(sym["NC"] * sym["L"], dtype=torch.int32, device=u.device)
                _lib.call("fb2_pair_positions", sym["NC"] * sym["L"], _lib.ptr(sym["adj_pair"]), _lib.ptr(pos), _lib.stream())
                sym["pair_pos"] = pos
            _lib.call("fb2_matfree_scalar_const", mesh.TD, space.p, sym["NC"], sym["gdof"], _lib.ptr(mesh.node), _lib.ptr(mesh.cell),
                      _lib.ptr(space.cell_to_dof().contiguous()), _lib.ptr(sym["adj_ptr"]), _lib.ptr(sym["adj_pair"]), _lib.ptr(pos),
                      self._host_table(dm, "Ms"), self._host_table(mm, "Mm"), sd, _lib.ptr(ad), sm_, _lib.ptr(am),
                      _lib.ptr(u.contiguous()), _lib.ptr(ws), _lib.ptr(v), _lib.stream())
            self.last_matfree = "fused"
            return v
        ke = self._summed_ke()
        self.last_matfree = "ke"
        _lib.call("fb2_matfree_apply", sym["gdof"], sym["L"], _lib.ptr(sym["adj_ptr"]), _lib.ptr(sym["adj_pair"]),
                  _lib.ptr(self.space.cell_to_dof().contiguous()), _lib.ptr(ke), _lib.ptr(u.contiguous()), _lib.ptr(v), _lib.stream())
        return v
