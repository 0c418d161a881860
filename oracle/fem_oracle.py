"""CPU oracle: a numpy restatement of FEALPy's Lagrange-FEM assembly + CG path.

TEST INFRASTRUCTURE ONLY.  Nothing under fealpy_b200/ may import this module; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do,
and there only as the checker / the timed CPU baseline.

Parity status: PINNED.  tools/gen_golden.py runs the real reference (imported from
/root/reference in the build container) and (a) asserts this restatement reproduces it
(pattern bit-exact, values <= 1e-13 rel) on the whole case ladder, (b) stores the
reference outputs as tests/golden/*.npz, which tests/test_oracle.py re-checks on every
run, together with the reference's own golden vectors (tests/golden/ref_testdata.npz,
extracted from /root/reference/test/**/_data.py by tools/extract_ref_testdata.py).

Every function cites the reference file:line it follows (paths relative to the
reference root).  The arithmetic is restated, not copied: same operation order where
the bits matter (COO order, stable sort, left-to-right duplicate sum, CG recurrence).
"""
from __future__ import annotations

import os
from itertools import combinations_with_replacement
from math import factorial

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_QUAD = None


# --------------------------------------------------------------------------------------
# quadrature (fealpy/quadrature/triangle.py:16-329, tetrahedron.py:7-243; data dumped by
# tools/gen_tables.py from the reference, digits untouched)
# --------------------------------------------------------------------------------------
def quadrature(TD: int, q: int):
    global _QUAD
    if _QUAD is None:
        _QUAD = dict(np.load(os.path.join(_HERE, "quadrature.npz")))
    name = {2: "tri", 3: "tet"}[TD]
    try:
        return _QUAD[f"{name}_q{q}_bcs"], _QUAD[f"{name}_q{q}_ws"]
    except KeyError:
        raise ValueError(f"no {name} quadrature table for q={q}")


# --------------------------------------------------------------------------------------
# reference-element basis  (fealpy/backend/numpy_backend.py:356-365, 423-477)
# --------------------------------------------------------------------------------------
def multi_index_matrix(p: int, TD: int) -> np.ndarray:
    """Rows = exponents (a_0..a_TD), sum p, in descending lexicographic order
    (numpy_backend.py:356-365 builds the same list from combinations_with_replacement)."""
    seps = np.array(tuple(combinations_with_replacement(range(p + 1), TD)), dtype=np.int32)[::-1]
    ext = np.zeros((seps.shape[0], TD + 2), dtype=np.int32)
    ext[:, 1:-1] = seps
    ext[:, -1] = p
    return ext[:, 1:] - ext[:, :-1]


def _A_table(bc: np.ndarray, p: int) -> np.ndarray:
    """A[..., m, b] = prod_{t<m} (p*lam_b - t) / m!   (numpy_backend.py:431-438)."""
    TD1 = bc.shape[-1]
    A = np.ones(bc.shape[:-1] + (p + 1, TD1), dtype=bc.dtype)
    for m in range(1, p + 1):
        A[..., m, :] = A[..., m - 1, :] * (p * bc - (m - 1))
    fact = 1.0
    for m in range(1, p + 1):
        fact *= m
        A[..., m, :] *= 1.0 / fact
    return A


def shape_function(bc: np.ndarray, p: int) -> np.ndarray:
    """phi[..., i] = prod_b A[mi[i,b], b]   (numpy_backend.py:423-440)."""
    if p == 1:
        return bc
    TD = bc.shape[-1] - 1
    mi = multi_index_matrix(p, TD)
    A = _A_table(bc, p)
    phi = np.ones(bc.shape[:-1] + (mi.shape[0],), dtype=bc.dtype)
    for b in range(TD + 1):
        phi = phi * A[..., mi[:, b], b]
    return phi


def grad_shape_function(bc: np.ndarray, p: int) -> np.ndarray:
    """R[..., i, b] = d phi_i / d lam_b   (numpy_backend.py:442-477).

    F[m,b] = d/dlam_b A[m,b] = (1/m!) * sum_{s<m} p * prod_{t<m, t!=s} (p lam_b - t)."""
    TD = bc.shape[-1] - 1
    mi = multi_index_matrix(p, TD)
    A = _A_table(bc, p)
    F = np.zeros_like(A)
    fact = 1.0
    for m in range(1, p + 1):
        fact *= m
        acc = np.zeros_like(bc)
        for s in range(m):
            term = np.full_like(bc, float(p))
            for t in range(m):
                if t != s:
                    term = term * (p * bc - t)
            acc = acc + term
        F[..., m, :] = acc * (1.0 / fact)
    ldof = mi.shape[0]
    R = np.zeros(bc.shape[:-1] + (ldof, TD + 1), dtype=bc.dtype)
    for b in range(TD + 1):
        val = F[..., mi[:, b], b]
        for b2 in range(TD + 1):
            if b2 != b:
                val = val * A[..., mi[:, b2], b2]
        R[..., b] = val
    return R


# --------------------------------------------------------------------------------------
# meshes  (fealpy/mesh/triangle_mesh.py:1386-1435, tetrahedron_mesh.py:1016-1086)
# --------------------------------------------------------------------------------------
def tri_from_box(box, nx, ny):
    x = np.linspace(box[0], box[1], nx + 1)
    y = np.linspace(box[2], box[3], ny + 1)
    X, Y = np.meshgrid(x, y, indexing="ij")
    node = np.stack([X.ravel(), Y.ravel()], axis=1)
    idx = np.arange((nx + 1) * (ny + 1), dtype=np.int32).reshape(nx + 1, ny + 1)
    # lower triangles then upper triangles, each block with j (y) slowest (.T before ravel)
    lo = np.stack([idx[1:, :-1].T.ravel(), idx[1:, 1:].T.ravel(), idx[:-1, :-1].T.ravel()], axis=1)
    up = np.stack([idx[:-1, 1:].T.ravel(), idx[:-1, :-1].T.ravel(), idx[1:, 1:].T.ravel()], axis=1)
    return node, np.concatenate([lo, up], axis=0).astype(np.int32)


_KUHN = np.array([[0, 1, 2, 6], [0, 5, 1, 6], [0, 4, 5, 6],
                  [0, 7, 4, 6], [0, 3, 7, 6], [0, 2, 3, 6]], dtype=np.int32)


def tet_from_box(box, nx, ny, nz):
    x = np.linspace(box[0], box[1], nx + 1)
    y = np.linspace(box[2], box[3], ny + 1)
    z = np.linspace(box[4], box[5], nz + 1)
    X, Y, Z = np.meshgrid(x, y, z, indexing="ij")
    node = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    idx = np.arange((nx + 1) * (ny + 1) * (nz + 1), dtype=np.int32).reshape(nx + 1, ny + 1, nz + 1)
    c0 = idx[:-1, :-1, :-1].ravel()
    nyz = (ny + 1) * (nz + 1)
    # cube corners 0..7: (0,0,0),(1,0,0),(1,1,0),(0,1,0),(0,0,1),(1,0,1),(1,1,1),(0,1,1)
    off = np.array([0, nyz, nyz + nz + 1, nz + 1, 1, nyz + 1, nyz + nz + 2, nz + 2], dtype=np.int32)
    cube = c0[:, None] + off[None, :]
    return node, cube[:, _KUHN].reshape(-1, 4).astype(np.int32)


_LOCAL_EDGE = {2: np.array([(1, 2), (2, 0), (0, 1)]),
               3: np.array([(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)])}
_LOCAL_FACE_TET = np.array([(1, 2, 3), (0, 3, 2), (0, 1, 3), (0, 2, 1)])


def _unique_rows_first_occurrence(total: np.ndarray):
    """(mesh/utils.py:81-110 `flocc`): ids = rank of the sorted tuple in lexicographic order,
    stored row = first occurrence in `total` order.  Returns (entity, ent_of_total)."""
    srt = np.sort(total, axis=1)
    order = np.lexsort(tuple(srt[:, k] for k in range(srt.shape[1] - 1, -1, -1)))
    s = srt[order]
    new = np.ones(len(s), dtype=bool)
    new[1:] = np.any(s[1:] != s[:-1], axis=1)
    gid_sorted = np.cumsum(new) - 1
    first = order[new]                      # lexsort is stable -> first occurrence
    ent_of_total = np.empty(len(s), dtype=np.int64)
    ent_of_total[order] = gid_sorted
    return total[first], ent_of_total


class Mesh:
    """Minimal simplex mesh: node, cell + the topology `MeshDS.construct` builds
    (mesh/mesh_data_structure.py:428-464)."""

    def __init__(self, node, cell):
        self.node = np.asarray(node, dtype=np.float64)
        self.cell = np.asarray(cell)
        self.TD = self.cell.shape[1] - 1
        self.GD = self.node.shape[1]
        NC = self.cell.shape[0]
        le = _LOCAL_EDGE[self.TD]
        total_edge = self.cell[:, le].reshape(-1, 2)
        self.edge, e_of = _unique_rows_first_occurrence(total_edge)
        self.cell2edge = e_of.reshape(NC, len(le))
        if self.TD == 3:
            total_face = self.cell[:, _LOCAL_FACE_TET].reshape(-1, 3)
            self.face, f_of = _unique_rows_first_occurrence(total_face)
            self.cell2face = f_of.reshape(NC, 4)
        else:
            self.face, self.cell2face = self.edge, self.cell2edge

    NN = property(lambda s: s.node.shape[0])
    NC = property(lambda s: s.cell.shape[0])
    NE = property(lambda s: s.edge.shape[0])
    NF = property(lambda s: s.face.shape[0])

    # --- geometry (triangle_mesh.py:45-98,115-129; tetrahedron_mesh.py:144-175,208-219;
    #               numpy_backend.py:413-421,586-598,619-629)
    def cell_measure(self):
        v = self.node[self.cell]
        e = v[:, 1:, :] - v[:, :-1, :]
        return np.linalg.det(e) / factorial(self.TD)

    def grad_lambda(self):
        v = self.node[self.cell]
        if self.TD == 2:
            e0 = v[:, 2] - v[:, 1]
            e1 = v[:, 0] - v[:, 2]
            e2 = v[:, 1] - v[:, 0]
            nv = e0[:, 0] * e1[:, 1] - e0[:, 1] * e1[:, 0]
            D = np.stack([np.stack([-e[:, 1], e[:, 0]], axis=1) for e in (e0, e1, e2)], axis=1)
            return D / nv[:, None, None]
        vol = self.cell_measure()
        D = np.zeros((self.NC, 4, 3))
        for i in range(4):
            j, k, m = _LOCAL_FACE_TET[i]
            vjk = v[:, k] - v[:, j]
            vjm = v[:, m] - v[:, j]
            D[:, i] = np.cross(vjm, vjk) / (6 * vol[:, None])
        return D

    def bc_to_point(self, bcs):
        return np.einsum("cjk,qj->cqk", self.node[self.cell], bcs)

    # --- dof numbering (mesh_base.py:188-210; triangle_mesh.py:218-270;
    #                    tetrahedron_mesh.py:348-441)
    def number_of_global_ipoints(self, p):
        n = self.NN + (p - 1) * self.NE
        if self.TD == 2:
            return n + (p - 1) * (p - 2) // 2 * self.NC
        return n + (p - 1) * (p - 2) // 2 * self.NF + (p - 1) * (p - 2) * (p - 3) // 6 * self.NC

    def cell_to_ipoint(self, p):
        """Global dof of local dof with exponents a: vertex -> node id; two non-zeros -> the
        t-th interior point of the (oriented, first-occurrence) global edge; three non-zeros ->
        interior point of the (oriented) face / triangle cell; four -> cell interior."""
        if p == 1:
            return self.cell
        TD, NC, NN, NE = self.TD, self.NC, self.NN, self.NE
        mi = multi_index_matrix(p, TD)
        ldof = mi.shape[0]
        c2d = np.zeros((NC, ldof), dtype=self.cell.dtype)
        le = _LOCAL_EDGE[TD]
        edge_lookup = {tuple(sorted(map(int, ab))): k for k, ab in enumerate(le)}
        fidof = (p - 1) * (p - 2) // 2
        if fidof:
            mi2 = multi_index_matrix(p, 2)
            interior2 = [tuple(int(v) for v in r) for r in mi2 if np.all(r > 0)]
            pos2 = {r: k for k, r in enumerate(interior2)}
        if TD == 3:
            face_lookup = {tuple(sorted(map(int, f))): k for k, f in enumerate(_LOCAL_FACE_TET)}
            cdof_idx = 0
        ar = np.arange(NC)
        cell_interior = 0
        for i in range(ldof):
            a = mi[i]
            nzv = [int(b) for b in np.nonzero(a)[0]]
            if len(nzv) == 1:
                c2d[:, i] = self.cell[:, nzv[0]]
            elif len(nzv) == 2:
                va, vb = nzv
                ge = self.cell2edge[:, edge_lookup[(va, vb)]]
                first_is_a = self.edge[ge, 0] == self.cell[:, va]
                # interior point t=1..p-1 counted from the stored first vertex e0 carries
                # weight (p-t)/p on e0 (tetrahedron_mesh.py:323-326)
                t = np.where(first_is_a, p - a[va], a[va])
                c2d[:, i] = NN + (p - 1) * ge + (t - 1)
            elif len(nzv) == 3 and TD == 3:
                gf = self.cell2face[:, face_lookup[tuple(nzv)]]
                fv = self.face[gf]                     # (NC,3) stored vertex order
                beta = np.zeros((NC, 3), dtype=np.int64)
                for lv in nzv:
                    gv = self.cell[:, lv]
                    for s in range(3):
                        beta[fv[:, s] == gv, s] = a[lv]
                idx = np.array([pos2[tuple(int(v) for v in r)] for r in beta])
                c2d[:, i] = NN + (p - 1) * NE + fidof * gf + idx
            else:  # cell interior (triangle: 3 non-zeros, tet: 4)
                if TD == 2:
                    c2d[:, i] = NN + (p - 1) * NE + fidof * ar + cell_interior
                else:
                    idof = (p - 1) * (p - 2) * (p - 3) // 6
                    c2d[:, i] = NN + (p - 1) * NE + fidof * self.NF + idof * ar + cell_interior
                cell_interior += 1
        return c2d

    def interpolation_points(self, p):
        """(triangle_mesh.py:181-216, tetrahedron_mesh.py:302-346) via cell-wise evaluation."""
        mi = multi_index_matrix(p, self.TD)
        pts = np.einsum("cjk,ij->cik", self.node[self.cell], mi / p)
        ip = np.zeros((self.number_of_global_ipoints(p), self.GD))
        ip[self.cell_to_ipoint(p).ravel()] = pts.reshape(-1, self.GD)
        ip[:self.NN] = self.node
        return ip

    def boundary_face_flag(self):
        """face adjacent to exactly one cell (mesh_data_structure.py:377-383)."""
        cnt = np.bincount(self.cell2face.ravel(), minlength=self.NF)
        return cnt == 1

    def face_to_ipoint(self, p):
        """dofs lying on each face, as a per-face index list (order irrelevant to callers
        that only scatter flags, functionspace/dofs.py:23-55)."""
        mi = multi_index_matrix(p, self.TD)
        c2d = self.cell_to_ipoint(p)
        out = {}
        nlf = self.TD + 1
        fdofs = [np.nonzero(mi[:, lf] == 0)[0] for lf in range(nlf)]
        # local face lf is opposite vertex lf for both tri (edges (1,2),(2,0),(0,1)) and tet
        res = np.zeros((self.NF, len(fdofs[0])), dtype=c2d.dtype)
        for lf in range(nlf):
            res[self.cell2face[:, lf]] = c2d[:, fdofs[lf]]
        return res


def tensor_cell_to_dof(c2d, gdof, dof_numel, dof_priority):
    """functionspace/utils.py:83-95."""
    if dof_priority:
        out = (np.arange(dof_numel)[None, :, None] * gdof + c2d[:, None, :])
    else:
        out = c2d[:, :, None] * dof_numel + np.arange(dof_numel)[None, None, :]
    return out.reshape(c2d.shape[0], -1).astype(c2d.dtype)


# --------------------------------------------------------------------------------------
# element matrices
# --------------------------------------------------------------------------------------
def grad_basis(mesh: Mesh, p: int, bcs):
    """gphi[c,q,i,m] = sum_b R[q,i,b] Dlam[c,b,m]  (triangle_mesh.py:142-153, mesh_base.py:713-749)."""
    R = grad_shape_function(bcs, p)
    return np.einsum("qib,cbm->cqim", R, mesh.grad_lambda())


def _apply_coef(core_args, subs_core, coef, NC, NQ):
    """functional.py:68-106 coefficient rules."""
    if coef is None:
        return np.einsum(subs_core + "->cij", *core_args, optimize=True)
    if np.isscalar(coef) or (isinstance(coef, np.ndarray) and coef.size == 1):
        return np.einsum(subs_core + "->cij", *core_args, optimize=True) * coef
    coef = np.asarray(coef)
    if coef.ndim == 4:
        return None  # handled by caller (matrix coefficient)
    while coef.ndim < 3:
        coef = coef[..., None]
    return np.einsum(subs_core + ",cqd->cij", *core_args, coef, optimize=True)


def eval_coef(mesh: Mesh, coef, bcs):
    """utils/utils.py:21-53: callables tagged coordtype='cartesian' get physical points,
    untagged callables get (bcs, index=...)."""
    if callable(coef):
        if getattr(coef, "coordtype", "barycentric") == "cartesian":
            return coef(mesh.bc_to_point(bcs))
        return coef(bcs, index=slice(None))
    return coef


def diffusion_element(mesh: Mesh, p: int, q=None, coef=None, method=None):
    """fem/scalar_diffusion_integrator.py:54-79."""
    q = p + 3 if q is None else q
    bcs, ws = quadrature(mesh.TD, q)
    cm = mesh.cell_measure()
    if method == "fast":
        R = grad_shape_function(bcs, p)
        M = np.einsum("q,qik,qjl->ijkl", ws, R, R)
        gl = mesh.grad_lambda()
        return np.einsum("ijkl,ckm,clm,c->cij", M, gl, gl, cm, optimize=True)
    gphi = grad_basis(mesh, p, bcs)
    coef = eval_coef(mesh, coef, bcs)
    if isinstance(coef, np.ndarray) and coef.ndim == 4:
        return np.einsum("q,c,cqid,cqjn,cqdn->cij", ws, cm, gphi, gphi, coef, optimize=True)
    return _apply_coef((ws, cm, gphi, gphi), "q,c,cqid,cqjd", coef, mesh.NC, len(ws))


def mass_element(mesh: Mesh, p: int, q=None, coef=None):
    """fem/scalar_mass_integrator.py:29-54."""
    q = p + 3 if q is None else q
    bcs, ws = quadrature(mesh.TD, q)
    cm = mesh.cell_measure()
    phi = shape_function(bcs, p)[None, :, :, None]          # (1,NQ,ldof,1)
    coef = eval_coef(mesh, coef, bcs)
    return _apply_coef((ws, cm, phi, phi), "q,c,cqid,cqjd", coef, mesh.NC, len(ws))


def lame(E, nu):
    """material/elastic_material.py:77-121."""
    return nu * E / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu))


def elastic_matrix(lam, mu, hypo="3D", E=None, nu=None):
    """material/elastic_material.py:238-257."""
    if hypo == "3D":
        D = np.diag([2 * mu + lam] * 3 + [mu] * 3).astype(np.float64)
        D[:3, :3] += lam * (1 - np.eye(3))
        return D
    if hypo == "plane_strain":
        return np.array([[2 * mu + lam, lam, 0], [lam, 2 * mu + lam, 0], [0, 0, mu]], dtype=np.float64)
    if hypo == "plane_stress":
        return E / (1 - nu ** 2) * np.array([[1, nu, 0], [nu, 1, 0], [0, 0, (1 - nu) / 2]], dtype=np.float64)
    raise NotImplementedError(hypo)


def elasticity_element(mesh: Mesh, p: int, D: np.ndarray, q=None, dof_priority=False):
    """fem/linear_elasticity_integrator.py:60-181 (default variant, simplex meshes)."""
    q = p + 3 if q is None else q
    bcs, ws = quadrature(mesh.TD, q)
    cm = mesh.cell_measure()
    gphi = grad_basis(mesh, p, bcs)
    GD = mesh.GD
    A = [[np.einsum("q,cqi,cqj,c->cij", ws, gphi[..., a], gphi[..., b], cm, optimize=True)
          for b in range(GD)] for a in range(GD)]
    ldof = gphi.shape[2]
    KK = np.zeros((mesh.NC, GD * ldof, GD * ldof))

    def blk(a, b):
        if dof_priority:
            return (slice(None), slice(a * ldof, (a + 1) * ldof), slice(b * ldof, (b + 1) * ldof))
        return (slice(None), slice(a, None, GD), slice(b, None, GD))
    if GD == 2:
        D00, D01, D22 = D[0, 0], D[0, 1], D[2, 2]
        KK[blk(0, 0)] = D00 * A[0][0] + D22 * A[1][1]
        KK[blk(1, 1)] = D00 * A[1][1] + D22 * A[0][0]
        KK[blk(0, 1)] = D01 * A[0][1] + D22 * A[1][0]
        KK[blk(1, 0)] = D01 * A[1][0] + D22 * A[0][1]
    else:
        D00, D01, D55 = D[0, 0], D[0, 1], D[5, 5]
        # the interleaved (dof_priority=False) branch uses 2*D55+D01 on the diagonal (:160-165)
        dd = D00 if dof_priority else (2 * D55 + D01)
        for a in range(3):
            b, c = [x for x in range(3) if x != a]
            KK[blk(a, a)] = dd * A[a][a] + D55 * (A[b][b] + A[c][c])
            for b in range(3):
                if b != a:
                    KK[blk(a, b)] = D01 * A[a][b] + D55 * A[b][a]
    return KK


# --------------------------------------------------------------------------------------
# COO -> CSR  (fem/bilinear_form.py:46-105, sparse/coo_tensor.py:137-157,184-213,329-342)
# --------------------------------------------------------------------------------------
def coo_from_groups(groups):
    """groups: list of (K_e (NC,l,l), cell2dof (NC,l)).  Entry order = group, cell, i, j;
    row = cell2dof[c,i], col = cell2dof[c,j]  (bilinear_form.py:69-72)."""
    rows, cols, vals = [], [], []
    for Ke, c2d in groups:
        shape = Ke.shape
        rows.append(np.broadcast_to(c2d[:, :, None], shape).ravel())
        cols.append(np.broadcast_to(c2d[:, None, :], shape).ravel())
        vals.append(Ke.reshape(-1))
    return np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)


def coalesce(row, col, val):
    """stable (row,col) sort; duplicates summed left to right in original order
    (coo_tensor.py:184-213; index_add -> np.add.at, numpy_backend.py:128-133)."""
    order = np.lexsort((col, row))
    r, c, v = row[order], col[order], val[order]
    new = np.ones(len(r), dtype=bool)
    new[1:] = (r[1:] != r[:-1]) | (c[1:] != c[:-1])
    out = np.zeros(int(new.sum()), dtype=val.dtype)
    np.add.at(out, np.cumsum(new) - 1, v)
    return r[new], c[new], out


def tocsr(row, col, val, nrow):
    """coo_tensor.py:137-157: crow int64, col keeps its dtype."""
    crow = np.concatenate([[0], np.cumsum(np.bincount(row, minlength=nrow))]).astype(np.int64)
    order = np.argsort(row, kind="stable")
    return crow, col[order], val[order]


def assemble(groups, gdof):
    r, c, v = coalesce(*coo_from_groups(groups))
    return tocsr(r, c, v, gdof)


def matfree_apply(groups, gdof, u):
    """BilinearForm.__matmul__ before assembly (fem/bilinear_form.py:126-158): per group
    gv = einsum('cij,cj->ci', K_e, u[cell2dof]) index_added into v (np.add.at = (cell, i) order)."""
    v = np.zeros(gdof)
    for ke, c2d in groups:
        gv = np.einsum("cij,cj->ci", ke, u[c2d])
        np.add.at(v, c2d.reshape(-1), gv.reshape(-1))
    return v


# --------------------------------------------------------------------------------------
# SpMV + CG  (backend/numpy_backend.py:180-199 -> scipy csr_matvec; solver/cg.py:14-123)
# --------------------------------------------------------------------------------------
def csr_matvec(crow, col, val, x):
    try:
        from scipy.sparse import csr_matrix
        n = len(crow) - 1
        return csr_matrix((val, col, crow), shape=(n, n)) @ x
    except ImportError:  # row-sequential dot, same order as csr_matvec
        y = np.zeros(len(crow) - 1)
        np.add.at(y, np.repeat(np.arange(len(crow) - 1), np.diff(crow)), val * x[col])
        return y


def cg(matvec, b, x0=None, Minv=None, atol=1e-12, rtol=1e-8, maxit=10000):
    """solver/cg.py:76-123, including its quirks: zero-rhs early return, the extra SpMV
    before the loop, stop test on sqrt(r.z) with '<', atol before rtol, maxit last."""
    info = {"residual": 0.0, "niter": 0}
    if np.linalg.norm(b) < 1e-15:
        return np.zeros_like(b), info
    x = np.zeros_like(b) if x0 is None else x0
    r = b - matvec(x)
    z = Minv(r) if Minv is not None else r
    p = z
    b_norm = np.linalg.norm(b)
    rTr = np.sum(r * z, axis=0)
    n = 0
    while True:
        Ap = matvec(p)
        alpha = rTr / np.sum(p * Ap, axis=0)
        x = x + alpha * p
        r_new = r - alpha * Ap
        z_new = Minv(r_new) if Minv is not None else r_new
        rTr_new = np.sum(r_new * z_new, axis=0)
        r_norm = np.sqrt(np.sum(rTr_new))
        n += 1
        info["residual"], info["niter"] = float(r_norm), n
        if r_norm < atol or r_norm < rtol * b_norm:
            break
        if maxit is not None and n >= maxit:
            break
        beta = rTr_new / rTr
        p = z_new + beta * p
        r, z, rTr = r_new, z_new, rTr_new
    return x, info


# --------------------------------------------------------------------------------------
# "next" rows: source vector, Dirichlet BC  (fem/scalar_source_integrator.py:13-57,
# fem/linear_form.py:36-86, fem/dirichlet_bc.py:101-235)
# --------------------------------------------------------------------------------------
def source_vector(mesh: Mesh, p: int, f, q=None):
    q = p + 3 if q is None else q
    bcs, ws = quadrature(mesh.TD, q)
    cm = mesh.cell_measure()
    phi = shape_function(bcs, p)                      # (NQ, ldof)
    val = eval_coef(mesh, f, bcs)
    if val is None:
        raise ValueError("source required")
    if np.isscalar(val):
        Fe = val * np.einsum("c,q,qi->ci", cm, ws, phi, optimize=True)
    else:
        val = np.asarray(val)
        if val.ndim == 1:
            val = val[:, None]
        Fe = np.einsum("c,q,qi,cq->ci", cm, ws, phi, np.broadcast_to(val, (mesh.NC, len(ws))), optimize=True)
    F = np.zeros(mesh.number_of_global_ipoints(p))
    np.add.at(F, mesh.cell_to_ipoint(p).ravel(), Fe.ravel())
    return F


def vector_source_vector(mesh: Mesh, p: int, f, ncomp: int, dof_priority: bool, q=None):
    """LinearForm + VectorSourceIntegrator on a TensorFunctionSpace (fem/vector_source_integrator.py:44-58,
    functional.py:57-59): F_e[c, I] = |K| sum_q w_q phi_I(q) . f(c, q) with the tensor basis phi_(i,a) = phi_i e_a
    (functionspace/utils.py generate_tensor_basis), local index i*ncomp + a (dof_priority False) or a*ldof + i (True);
    index_added through the tensor cell2dof."""
    q = p + 3 if q is None else q
    bcs, ws = quadrature(mesh.TD, q)
    cm = mesh.cell_measure()
    phi = shape_function(bcs, p)                      # (NQ, ldof)
    val = eval_coef(mesh, f, bcs)                     # (NC, NQ, ncomp)
    fe = np.einsum("c,q,qi,cqa->cia", cm, ws, phi, val)
    NC, L = fe.shape[0], fe.shape[1]
    fe = fe.transpose(0, 2, 1).reshape(NC, ncomp * L) if dof_priority else fe.reshape(NC, L * ncomp)
    c2d = mesh.cell_to_ipoint(p)
    sg = mesh.number_of_global_ipoints(p)
    c2d_t = tensor_cell_to_dof(c2d, sg, ncomp, dof_priority)
    F = np.zeros(sg * ncomp)
    np.add.at(F, c2d_t.reshape(-1), fe.reshape(-1))
    return F


def boundary_dof_flag(mesh: Mesh, p: int, threshold=None, method=None):
    """functionspace/dofs.py:23-55.  method None / 'centroid': a callable threshold selects boundary FACES by their
    barycentres and every dof of a kept face is flagged; 'interp': it selects boundary dofs by their interpolation points."""
    gdof = mesh.number_of_global_ipoints(p)
    flag = np.zeros(gdof, dtype=bool)
    f2d = mesh.face_to_ipoint(p)
    index = np.nonzero(mesh.boundary_face_flag())[0]
    if method is None or method == "centroid":
        if callable(threshold):
            bc = mesh.node[mesh.face[index]].mean(axis=1)
            index = index[threshold(bc)]
        flag[f2d[index].ravel()] = True
    elif method == "interp":
        dofs = f2d[index].ravel()
        if callable(threshold):
            dofs = dofs[threshold(mesh.interpolation_points(p)[dofs])]
        flag[dofs] = True
    else:
        raise ValueError(f"Unknown method: {method}")
    return flag


def tensor_boundary_dof_flag(mesh: Mesh, p: int, ncomp: int, dof_priority: bool, threshold=None, method=None):
    """functionspace/tensor_space.py:159-188 (one threshold for all components)"""
    f = boundary_dof_flag(mesh, p, threshold, method)
    return np.tile(f, ncomp) if dof_priority else np.repeat(f, ncomp)


def dirichlet_apply(crow, col, val, F, uh, is_bd):
    """dirichlet_bc.py:101-235 in canonical (sorted-column) CSR form:
    F <- F - A uh, F[bd] = uh[bd]; rows/cols of bd dofs zeroed (entries kept as explicit
    zeros are dropped by the reference's COO round trip -- compare as matrices), unit diagonal."""
    n = len(crow) - 1
    F = F - csr_matvec(crow, col, val, uh)
    F[is_bd] = uh[is_bd]
    rows = np.repeat(np.arange(n), np.diff(crow))
    keep = ~(is_bd[rows] | is_bd[col])
    from scipy.sparse import csr_matrix, diags
    A = csr_matrix((val * keep, col, crow), shape=(n, n)) + diags(is_bd.astype(np.float64))
    return A.tocsr(), F
