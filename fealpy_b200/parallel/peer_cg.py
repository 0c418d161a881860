"""Distributed CG whose per-iteration communication is done by the CG kernels themselves over NVLink peer memory
(csrc/peer.cu): the p-update kernel stores the owned boundary slices straight into the neighbours' vectors, the two scalar
reductions are one-warp kernels that exchange partial sums through mailboxes in every rank's symmetric buffer, and the SpMV is
split into interior rows (no halo column, runs while the neighbours' pushes are in flight) and boundary rows (after a flag
wait).  No NCCL call and no host synchronisation inside the iteration; torch.distributed is used for the set-up of a solve only.
Same recurrence, stopping rules and iteration count as the single-GPU solver (fealpy/solver/cg.py:76-123); the reductions
are summed in rank order on every rank, so all ranks take bit-identical decisions.
"""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.distributed as dist

from .. import _lib
from .dist_cg import SC_RTR, SC_PAP, SC_RTR_NEW, SC_RNORM, SC_NITER_I32, SC_DONE_I32, halo_exchange


def _range_plan(A, ranges, tile):
    """tile plan (first row, one-past-last row per tile) covering several row ranges of A"""
    lib = _lib.load()
    dev = A.device
    lo_parts, hi_parts, max_row = [], [], 0
    for lo, hi in ranges:
        if hi <= lo:
            continue
        nnz_r = int(A.crow[hi] - A.crow[lo])
        if nnz_r == 0:
            continue
        nblk = lib.fb2_spmv_plan_blocks(nnz_r, tile)
        blk = torch.empty(nblk + 2, dtype=torch.int32, device=dev)
        mr = C.c_int32(0)
        sub = A.crow[lo:hi + 1]
        _lib.call("fb2_spmv_plan_build", hi - lo, _lib.ptr(sub), tile, _lib.ptr(blk), nnz_r, C.byref(mr), _lib.stream())
        max_row = max(max_row, mr.value)
        lo_parts.append(blk[:nblk] + lo)
        hi_parts.append(blk[1:nblk + 1] + lo)
    if not lo_parts:
        z = torch.zeros(1, dtype=torch.int32, device=dev)
        return z, z, 0, 0
    blo, bhi = torch.cat(lo_parts).contiguous(), torch.cat(hi_parts).contiguous()
    return blo, bhi, int(blo.shape[0]), max_row


class PeerCG:
    mode = "peer"

    def __init__(self, A, part, minv=None, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        if not dist.is_initialized():
            raise RuntimeError("PeerCG needs an initialised torch.distributed process group")
        grp = group if group is not None else dist.group.WORLD
        self.A, self.part, self.group, self.minv = A, part, group, minv
        self.world, self.rank = dist.get_world_size(grp), dist.get_rank(grp)
        lib = _lib.load()
        dev = A.device
        self.n = n = A.sparse_shape[0]
        if self.world > 16 or len(part.exchanges) > 2:
            raise RuntimeError("PeerCG: at most 16 ranks and 2 neighbours (slab partitions)")
        ctrl_bytes = lib.fb2_peer_ctrl_bytes()
        nmax = torch.tensor([n], dtype=torch.int64, device=dev)
        dist.all_reduce(nmax, op=dist.ReduceOp.MAX, group=grp)          # symmetric allocations have one size on all ranks
        self.buf = symm_mem.empty(ctrl_bytes // 8 + int(nmax.item()), dtype=torch.float64, device=dev)
        self.hdl = symm_mem.rendezvous(self.buf, grp)
        self.buf.zero_()
        self.ctrl = self.buf[: ctrl_bytes // 8]
        self.ctrl_i64 = self.ctrl.view(torch.int64)
        self.p = self.buf[ctrl_bytes // 8: ctrl_bytes // 8 + n]
        ptrs = [int(q) for q in self.hdl.buffer_ptrs]
        self.peer_base = torch.tensor(ptrs, dtype=torch.int64, device=dev)
        # where my send slices land: the neighbour's matching recv slice (same order: nodes first, then edges)
        mine = [(ex.peer, lo, hi) for ex in part.exchanges for (lo, hi) in ex.send]
        local = [(ex.peer, lo, hi - lo) for ex in part.exchanges for (lo, hi) in ex.recv]
        allrecv = [None] * self.world
        dist.all_gather_object(allrecv, local, group=grp)
        taken = {}
        lo_a, hi_a, plo_a, pp_a = [], [], [], []
        for (peer, lo, hi) in mine:
            cands = [l for (src, l, sz) in allrecv[peer] if src == self.rank and sz == hi - lo]
            k = taken.get((peer, hi - lo), 0)
            taken[(peer, hi - lo)] = k + 1
            lo_a.append(lo); hi_a.append(hi); plo_a.append(cands[k]); pp_a.append(ptrs[peer] + ctrl_bytes)
        m = len(mine)
        self.nb = sorted({ex.peer for ex in part.exchanges})
        self.push = dict(n=m, lo=(C.c_int64 * max(m, 1))(*lo_a), hi=(C.c_int64 * max(m, 1))(*hi_a), plo=(C.c_int64 * max(m, 1))(*plo_a),
                         pp=(C.c_void_p * max(m, 1))(*pp_a), nnb=len(self.nb), nbc=(C.c_void_p * 2)(*([ptrs[r] for r in self.nb] + [0, 0])[:2]),
                         nbr=(C.c_int32 * 2)(*((self.nb + [0, 0])[:2])))
        # hflag slots live after red (2*2*16 doubles) and rflag (2*16): offsets in 8-byte words
        self.hflag_off = 2 * 2 * 16 + 2 * 16
        self.own = (C.c_int64 * 4)(*[int(v) for v in part.own_ranges])
        interior, boundary = part.row_split()
        tile = A.SPMV_TILE
        # ONE plan: interior tiles first, boundary tiles (which wait for the neighbours' halo pushes inside the kernel) last
        ilo, ihi, ni, mri = _range_plan(A, interior, tile)
        blo_, bhi_, nbd, mrb = _range_plan(A, boundary, tile)
        if ni and nbd:
            self.plan = (torch.cat([ilo, blo_]).contiguous(), torch.cat([ihi, bhi_]).contiguous(), ni + nbd, ni, max(mri, mrb))
        elif nbd:
            self.plan = (blo_, bhi_, nbd, 0, mrb)
        else:
            self.plan = (ilo, ihi, ni, ni, mri)
        self.tile = tile
        self.sc = torch.zeros(32, dtype=torch.float64, device=dev)
        self.sc_i32 = self.sc.view(torch.int32)
        self.r = torch.empty(n, dtype=torch.float64, device=dev)
        self.Ap = torch.zeros(n, dtype=torch.float64, device=dev)
        self.x = torch.empty(n, dtype=torch.float64, device=dev)
        self.pws = _lib.partial_ws(dev)
        self.counter = torch.zeros(4, dtype=torch.int32, device=dev)
        self.epoch = torch.zeros(1, dtype=torch.int64, device=dev)
        self.epoch_host = 0
        self.full_plan = A.spmv_plan()
        self.tmp = torch.zeros(1, dtype=torch.float64, device=dev)
        dist.barrier(group=grp)                          # every rank's buffers are zeroed and mapped before anybody pushes

    # ---- pieces of one solve -------------------------------------------------------------------------------------
    def _dot_owned(self, a, b):
        tot = 0.0
        o = [int(v) for v in self.part.own_ranges]
        for lo, hi in ((o[0], o[1]), (o[2], o[3])):
            if hi > lo:
                _lib.call("fb2_dot", hi - lo, _lib.ptr(a[lo:hi]), _lib.ptr(b[lo:hi]), _lib.ptr(self.tmp), _lib.ptr(self.pws), _lib.stream())
                tot += float(self.tmp.item())
        return tot

    def _iteration(self):
        A, st = self.A, _lib.stream()
        sc, pws = _lib.ptr(self.sc), _lib.ptr(self.pws)
        pAp = C.c_void_p(self.sc.data_ptr() + 8 * SC_PAP)
        rtn = C.c_void_p(self.sc.data_ptr() + 8 * SC_RTR_NEW)
        ep = _lib.ptr(self.epoch)
        blo, bhi, nb, first_bnd, mr = self.plan
        _lib.call("fb2_cg_spmv_dot_ranges", self.n, A.nnz, _lib.ptr(A.crow), _lib.ptr(A.col), _lib.ptr(A.values), _lib.ptr(self.p),
                  _lib.ptr(self.Ap), _lib.ptr(blo), _lib.ptr(bhi), nb, first_bnd, self.tile, mr, pAp, sc, pws, self.own,
                  _lib.ptr(self.ctrl), self.push["nnb"], self.push["nbr"], ep, st)
        _lib.call("fb2_peer_allreduce", _lib.ptr(self.ctrl), _lib.ptr(self.peer_base), self.world, self.rank, 0, pAp, None, pAp, sc, 0, ep, st)
        _lib.call("fb2_cg_update_xr", self.n, _lib.ptr(self.x), _lib.ptr(self.r), _lib.ptr(self.p), _lib.ptr(self.Ap), _lib.ptr(self.minv),
                  sc, pws, 0, self.own, st)
        _lib.call("fb2_peer_allreduce", _lib.ptr(self.ctrl), _lib.ptr(self.peer_base), self.world, self.rank, 1, rtn, None, rtn, sc, 1, ep, st)
        pu = self.push
        _lib.call("fb2_cg_update_p_push", self.own, _lib.ptr(self.p), _lib.ptr(self.r), _lib.ptr(self.minv), sc, pu["n"], pu["lo"], pu["hi"],
                  pu["plo"], pu["pp"], pu["nnb"], pu["nbc"], self.rank, _lib.ptr(self.counter), ep, st)

    def solve(self, b, x0=None, *, atol=1e-12, rtol=1e-8, maxit=10000, check_every=8):
        grp, A = self.group, self.A
        bb = torch.tensor([self._dot_owned(b, b)], dtype=torch.float64, device=b.device)
        dist.all_reduce(bb, op=dist.ReduceOp.SUM, group=grp)
        bnorm = math.sqrt(float(bb.item()))
        if bnorm < 1e-15:
            return torch.zeros_like(b), {"residual": 0.0, "niter": 0}
        x = self.x
        x.zero_() if x0 is None else x.copy_(x0)
        halo_exchange(x, self.part.exchanges, grp)
        _lib.call("fb2_cg_init", _lib.ptr(self.sc), float(atol), float(rtol), -1 if maxit is None else int(maxit), float(bnorm), 0.0,
                  _lib.stream())
        blk, tile, mr = self.full_plan
        _lib.call("fb2_cg_residual", self.n, A.nnz, _lib.ptr(A.crow), _lib.ptr(A.col), _lib.ptr(A.values), _lib.ptr(x), _lib.ptr(b),
                  _lib.ptr(self.r), _lib.ptr(blk), tile, mr, _lib.stream())
        _lib.call("fb2_cg_start", self.n, _lib.ptr(self.r), _lib.ptr(self.minv), _lib.ptr(self.p), _lib.ptr(self.sc), _lib.ptr(self.pws),
                  self.own, _lib.stream())
        dist.all_reduce(self.sc[SC_RTR:SC_RTR + 1], op=dist.ReduceOp.SUM, group=grp)
        halo_exchange(self.p, self.part.exchanges, grp)
        # sequence numbers of this solve: flags of earlier solves compare "older"; the first iteration's halo came by NCCL
        self.epoch_host += 1 << 32
        self.epoch.fill_(self.epoch_host)
        for nbr in self.nb:
            self.ctrl_i64[self.hflag_off + nbr] = self.epoch_host
        torch.cuda.current_stream().synchronize()
        dist.barrier(group=grp)                          # nobody enters the iteration before every rank has armed its flags
        it = 0
        limit = maxit if maxit is not None else 1 << 30
        while True:
            self._iteration()
            it += 1
            if it % check_every == 0 or it >= limit:
                st = self.sc_i32[SC_NITER_I32:SC_DONE_I32 + 1].tolist()
                if st[1]:
                    break
        niter = int(self.sc_i32[SC_NITER_I32].item())
        return x, {"residual": float(self.sc[SC_RNORM].item()), "niter": niter}
