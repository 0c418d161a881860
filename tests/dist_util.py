"""numpy backend of the distributed CG driver (tests only): the same `ops` interface as
fealpy_b200.parallel.CudaCgOps, computing with the oracle's csr_matvec on CPU tensors."""
import numpy as np
import torch

from oracle import fem_oracle as O
from fealpy_b200.parallel.dist_cg import SC_RTR, SC_PAP, SC_RTR_NEW, SC_BNORM, SC_RNORM, SC_ALPHA, SC_BETA


class NumpyCgOps:
    def __init__(self, crow, col, val, own, minv=None):
        self.crow, self.col, self.val = crow, col, val
        self.n = len(crow) - 1
        self.own = own
        self.mask = np.zeros(self.n, dtype=bool)
        self.mask[own[0]:own[1]] = True
        self.mask[own[2]:own[3]] = True
        self.minv = minv
        self.sc = torch.zeros(32, dtype=torch.float64)
        self.p = torch.zeros(self.n, dtype=torch.float64)
        self.r = torch.zeros(self.n, dtype=torch.float64)
        self.Ap = torch.zeros(self.n, dtype=torch.float64)
        self.niter, self.done = 0, False

    def _mv(self, v):
        return O.csr_matvec(self.crow, self.col, self.val, v.numpy())

    def _z(self, r):
        return r if self.minv is None else self.minv * r

    def dot_owned(self, a, b):
        return float(np.sum(a.numpy()[self.mask] * b.numpy()[self.mask]))

    def init(self, atol, rtol, maxit, bnorm):
        self.atol, self.rtol, self.maxit = atol, rtol, (1 << 30) if maxit is None else maxit
        self.sc[SC_BNORM] = bnorm

    def residual(self, x, b):
        self.r[:] = torch.from_numpy(b.numpy() - self._mv(x))

    def start(self):
        r = self.r.numpy()
        z = self._z(r)
        self.p[:] = torch.from_numpy(z.copy())
        self.sc[SC_RTR] = float(np.sum((r * z)[self.mask]))

    def spmv_dot(self):
        if self.done:
            return
        self.Ap[:] = torch.from_numpy(self._mv(self.p))
        self.sc[SC_PAP] = float(np.sum((self.p.numpy() * self.Ap.numpy())[self.mask]))

    def update_xr(self, x):
        if self.done:
            return
        alpha = float(self.sc[SC_RTR] / self.sc[SC_PAP])
        x += alpha * self.p
        self.r -= alpha * self.Ap
        r = self.r.numpy()
        self.sc[SC_RTR_NEW] = float(np.sum((r * self._z(r))[self.mask]))
        self.sc[SC_ALPHA] = alpha

    def finalize(self):
        if self.done:
            return
        rn = float(np.sqrt(self.sc[SC_RTR_NEW]))
        self.sc[SC_RNORM] = rn
        self.niter += 1
        if rn < self.atol or rn < self.rtol * float(self.sc[SC_BNORM]) or self.niter >= self.maxit:
            self.done = True
        else:
            self.sc[SC_BETA] = float(self.sc[SC_RTR_NEW] / self.sc[SC_RTR])
            self.sc[SC_RTR] = float(self.sc[SC_RTR_NEW])

    def update_p(self):
        if self.done:
            return
        self.p[:] = torch.from_numpy(self._z(self.r.numpy())) + float(self.sc[SC_BETA]) * self.p

    def scalar(self, slot):
        return self.sc[slot:slot + 1]

    def status(self):
        return self.niter, self.done

    def residual_norm(self):
        return float(self.sc[SC_RNORM])


def oracle_slab_problem(part, p):
    """what a rank assembles: the oracle on the slab's cells with window-local dof ids"""
    nx, ny, nz = part.dims
    node, cell = O.tet_from_box([0, 1, 0, 1, 0, 1], nx, ny, nz)
    gm = O.Mesh(node, cell)
    gc2d = gm.cell_to_ipoint(p)
    cells = np.arange(6 * part.cl0 * ny * nz, 6 * part.cl1 * ny * nz)
    # global -> window-local
    g2l = np.full(part.gdof, -1, dtype=np.int64)
    l = np.arange(part.n_local)
    g2l[part.local_to_global(l)] = l
    c2d_loc = g2l[gc2d[cells]]
    assert c2d_loc.min() >= 0
    nloc_nodes = np.arange(part.cl0 * part.nyz, (part.cl1 + 1) * part.nyz)
    lm = O.Mesh(node[nloc_nodes], cell[cells] - part.cl0 * part.nyz)
    Kd, Km = O.diffusion_element(lm, p), O.mass_element(lm, p)
    crow, col, val = O.assemble([(Kd, c2d_loc.astype(np.int32)), (Km, c2d_loc.astype(np.int32))], part.n_local)
    return gm, gc2d, (crow, col, val), c2d_loc
