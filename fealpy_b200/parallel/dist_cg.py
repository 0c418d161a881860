"""Distributed CG over a row partition: one process per GPU, halo exchange + two scalar
all-reduces per iteration through torch.distributed (NCCL over NVLink on GPUs; gloo in the CPU
tests).  The recurrence, stopping rules and iteration count are those of the single-GPU solver
(fealpy/solver/cg.py:76-123); only the inner products are summed over ranks.

The driver is written against a small `ops` interface so that the host-side logic (exchange
schedule, reductions, convergence handling) is exercised on CPU with a numpy backend in
tests/, while the product backend (`CudaCgOps`) launches the CUDA kernels of csrc/cg.cu with
all scalars resident on the device.
"""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.distributed as dist

from .. import _lib

# slots of the device-resident scalar block (csrc/cg.cuh CgScalars): doubles first
SC_RTR, SC_PAP, SC_RTR_NEW, SC_BNORM, SC_RNORM, SC_ALPHA, SC_BETA, SC_TMP = range(8)
SC_NITER_I32, SC_DONE_I32 = 16, 17          # int32 view


def _pack(v, idx):
    """contiguous copy of v[idx] (the owned values a neighbour needs): the library's gather kernel on CUDA tensors"""
    if v.is_cuda:
        out = torch.empty(idx.shape[0], dtype=v.dtype, device=v.device)
        _lib.call("fb2_gather_f64", idx.shape[0], _lib.ptr(idx), _lib.ptr(v), _lib.ptr(out), _lib.stream())
        return out
    return v[idx]


def halo_exchange(v, exchanges, group=None):
    """send owned boundary values, receive halo values.  Slab partitions (box_partition.Exchange) send contiguous slices
    of `v`; general partitions (mesh_partition.PackedExchange) send a packed gather v[send_idx]; every receive lands
    directly in the peer's contiguous halo range."""
    if not exchanges:
        return
    ops, keep = [], []
    for ex in exchanges:
        idx = getattr(ex, "send_idx", None)
        if idx is not None:
            if idx.numel():
                buf = _pack(v, idx)
                keep.append(buf)
                ops.append(dist.P2POp(dist.isend, buf, ex.peer, group))
        else:
            for lo, hi in ex.send:
                ops.append(dist.P2POp(dist.isend, v[lo:hi], ex.peer, group))
        for lo, hi in ex.recv:
            if hi > lo:
                ops.append(dist.P2POp(dist.irecv, v[lo:hi], ex.peer, group))
    if not ops:
        return
    for w in dist.batch_isend_irecv(ops):
        w.wait()


class CudaCgOps:
    """CUDA backend of the distributed driver: kernels of csrc/cg.cu on the local window matrix"""

    def __init__(self, A, own, minv=None):
        self.A, self.n = A, A.sparse_shape[0]
        self.own = (C.c_int64 * 4)(*[int(v) for v in own])
        self.own_t = tuple(int(v) for v in own)
        dev = A.device
        self.minv = minv
        self.sc = torch.zeros(32, dtype=torch.float64, device=dev)
        self.sc_i32 = self.sc.view(torch.int32)
        self.r = torch.empty(self.n, dtype=torch.float64, device=dev)
        self.p = torch.empty(self.n, dtype=torch.float64, device=dev)
        self.Ap = torch.empty(self.n, dtype=torch.float64, device=dev)
        self.pws = _lib.partial_ws(dev)
        self.plan = A.spmv_plan()

    def dot_owned(self, a, b):
        tot = 0.0
        out = torch.zeros(1, dtype=torch.float64, device=a.device)
        for lo, hi in ((self.own_t[0], self.own_t[1]), (self.own_t[2], self.own_t[3])):
            if hi > lo:
                _lib.call("fb2_dot", hi - lo, _lib.ptr(a[lo:hi]), _lib.ptr(b[lo:hi]), _lib.ptr(out), _lib.ptr(self.pws), _lib.stream())
                tot += float(out.item())
        return tot

    def init(self, atol, rtol, maxit, bnorm):
        _lib.call("fb2_cg_init", _lib.ptr(self.sc), float(atol), float(rtol), -1 if maxit is None else int(maxit), float(bnorm),
                  0.0, _lib.stream())

    def residual(self, x, b):
        A, (blk, tile, mr) = self.A, self.plan
        _lib.call("fb2_cg_residual", self.n, A.nnz, _lib.ptr(A.crow), _lib.ptr(A.col), _lib.ptr(A.values), _lib.ptr(x), _lib.ptr(b),
                  _lib.ptr(self.r), _lib.ptr(blk), tile, mr, _lib.stream())

    def start(self):
        _lib.call("fb2_cg_start", self.n, _lib.ptr(self.r), _lib.ptr(self.minv), _lib.ptr(self.p), _lib.ptr(self.sc),
                  _lib.ptr(self.pws), self.own, _lib.stream())

    def spmv_dot(self):
        A, (blk, tile, mr) = self.A, self.plan
        _lib.call("fb2_cg_spmv_dot", self.n, A.nnz, _lib.ptr(A.crow), _lib.ptr(A.col), _lib.ptr(A.values), _lib.ptr(self.p),
                  _lib.ptr(self.Ap), _lib.ptr(blk), tile, mr, _lib.ptr(self.sc), _lib.ptr(self.pws), self.own, _lib.stream())

    def update_xr(self, x):
        _lib.call("fb2_cg_update_xr", self.n, _lib.ptr(x), _lib.ptr(self.r), _lib.ptr(self.p), _lib.ptr(self.Ap), _lib.ptr(self.minv),
                  _lib.ptr(self.sc), _lib.ptr(self.pws), 0, self.own, _lib.stream())

    def finalize(self):
        _lib.call("fb2_cg_finalize", _lib.ptr(self.sc), _lib.stream())

    def update_p(self):
        _lib.call("fb2_cg_update_p", self.n, _lib.ptr(self.p), _lib.ptr(self.r), _lib.ptr(self.minv), _lib.ptr(self.sc), _lib.stream())

    def scalar(self, slot):
        return self.sc[slot:slot + 1]

    def status(self):
        st = self.sc_i32[SC_NITER_I32:SC_DONE_I32 + 1].tolist()
        return st[0], bool(st[1])

    def residual_norm(self):
        return float(self.sc[SC_RNORM].item())


def dist_cg(ops, b, x0, exchanges, *, atol=1e-12, rtol=1e-8, maxit=10000, check_every=8, group=None, x_out=None):
    """x (window-local vector; owned entries are the solution), info = {'residual', 'niter'}"""
    world = dist.get_world_size(group) if dist.is_initialized() else 1

    def allreduce(t):
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)

    bb = torch.tensor([ops.dot_owned(b, b)], dtype=torch.float64, device=b.device)
    allreduce(bb)
    bnorm = math.sqrt(float(bb.item()))
    if bnorm < 1e-15:
        return torch.zeros_like(b), {"residual": 0.0, "niter": 0}
    if x_out is None:
        x = torch.zeros_like(b) if x0 is None else x0.clone()
    else:                                    # caller-owned solution buffer (persistent solver: no allocation per solve)
        x = x_out
        x.zero_() if x0 is None else x.copy_(x0)
    halo_exchange(x, exchanges, group)
    ops.init(atol, rtol, maxit, bnorm)
    ops.residual(x, b)                       # r = b - A x on every local row (halo rows are scratch)
    ops.start()                              # p = z = M r, local r.z
    allreduce(ops.scalar(SC_RTR))
    it = 0
    limit = maxit if maxit is not None else 1 << 30
    while True:
        halo_exchange(ops.p, exchanges, group)
        ops.spmv_dot()
        allreduce(ops.scalar(SC_PAP))
        ops.update_xr(x)
        allreduce(ops.scalar(SC_RTR_NEW))
        ops.finalize()
        ops.update_p()
        it += 1
        if it % check_every == 0 or it >= limit:
            niter, done = ops.status()
            if done:
                break
    niter, _ = ops.status()
    return x, {"residual": ops.residual_norm(), "niter": niter}


class DistCG:
    """Persistent distributed CG solver for one matrix: work vectors, the device scalar block and the SpMV plan are
    allocated ONCE (nothing is allocated inside a timed solve), `solve()` can be called repeatedly.

    part: any object with `own_ranges` (lo0, hi0, lo1, hi1 in local ids) and `exchanges` (list of Exchange)."""

    mode = "nccl"

    def __init__(self, A, part, minv=None, group=None):
        self.A, self.part, self.group = A, part, group
        self.ops = CudaCgOps(A, part.own_ranges, minv)
        self.x = torch.empty(A.sparse_shape[0], dtype=torch.float64, device=A.device)

    def solve(self, b, x0=None, *, atol=1e-12, rtol=1e-8, maxit=10000, check_every=8):
        return dist_cg(self.ops, b, x0, self.part.exchanges, atol=atol, rtol=rtol, maxit=maxit, check_every=check_every,
                       group=self.group, x_out=self.x)


def make_dist_solver(A, part, minv=None, group=None, mode="auto"):
    """'nccl': halo exchange + scalar all-reduces through torch.distributed; 'peer': the exchange and the reductions are
    fused into the CG kernels over NVLink peer memory (parallel/peer_cg.py); 'auto' picks 'peer' when it can be set up"""
    if mode in ("auto", "peer"):
        try:
            from .peer_cg import PeerCG
            return PeerCG(A, part, minv=minv, group=group)
        except Exception:
            if mode == "peer":
                raise
    return DistCG(A, part, minv=minv, group=group)
