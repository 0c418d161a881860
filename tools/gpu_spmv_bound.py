"""What bounds the streaming SpMV?  Same CSR structure (crow, values) of config 2, column patterns of decreasing gather cost:
   real | morton (dofs renumbered along a Z-curve of their coordinates, rows kept: only the x-gather locality changes) |
   banded (a row's columns are consecutive) | zero (every gather hits x[0]).
   python tools/gpu_spmv_bound.py [n]"""
import os
import statistics
import sys

sys.path.insert(0, os.getcwd())
import torch
import bench
from fealpy_b200.sparse import CSRTensor
from fealpy_b200.parallel.mesh_partition import morton_codes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda", 0)
prob = bench.Problem(2, n, dev, 1, 0)
A = prob.assemble()
gdof = A.shape[0]
x = torch.rand(gdof, dtype=torch.float64, device=dev)


def timed(M, label):
    y = M @ x
    torch.cuda.synchronize()
    ts = []
    for _ in range(12):
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        y = M @ x
        s1.record()
        torch.cuda.synchronize()
        ts.append(s0.elapsed_time(s1))
    print(f"n {n} {label:28s}: min {min(ts[2:]):.4f} med {statistics.median(ts[2:]):.4f} ms  ({12 * M.nnz / min(ts[2:]) / 1e6:.0f} GB/s of (val, col))", flush=True)


timed(A, "real columns")
rows = torch.repeat_interleave(torch.arange(gdof, device=dev, dtype=torch.int32), (A.crow[1:] - A.crow[:-1]).to(torch.int32))
k = torch.arange(A.nnz, device=dev, dtype=torch.int64) - A.crow[:-1][rows.long()]
banded = torch.clamp(rows.long() + k - 14, 0, gdof - 1).to(torch.int32)
del k
timed(CSRTensor(A.crow, banded, A.values, A.shape), "banded columns")
del banded
timed(CSRTensor(A.crow, torch.zeros_like(A.col), A.values, A.shape), "all columns = 0")
# Morton renumbering of the COLUMN space only (x is gathered through the new numbering; rows stay where they are):
ip = prob.sspace.interpolation_points()
perm = torch.argsort(morton_codes(ip), stable=True)            # new position -> old dof
inv = torch.empty_like(perm)
inv[perm] = torch.arange(gdof, device=dev)
col_m = inv[A.col.long()].to(torch.int32)
timed(CSRTensor(A.crow, col_m, A.values, A.shape), "morton columns (rows kept)")
# full symmetric permutation P A P^T: rows in Morton order too (what an internal reordering would run)
order = perm
lens = (A.crow[1:] - A.crow[:-1])[order]
crow_p = torch.zeros(gdof + 1, dtype=torch.int64, device=dev)
crow_p[1:] = torch.cumsum(lens, 0)
src = torch.repeat_interleave(A.crow[:-1][order], lens) + (torch.arange(A.nnz, device=dev) - torch.repeat_interleave(crow_p[:-1], lens))
col_p = col_m[src]
val_p = A.values[src]
del src, col_m
timed(CSRTensor(crow_p, col_p, val_p, A.shape), "morton rows + columns")
