"""Time CG iterations (fixed count, atol = rtol = 0) of a config with CUDA events; FB2_SPMV_COLZ=0 disables the
compressed column stream.   python tools/gpu_time_cg.py [config] [n] [iters]"""
import os
import statistics
import sys

sys.path.insert(0, os.getcwd())
import torch
import bench
from fealpy_b200.solver import cg

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else bench.CONFIGS[cfg]["n"]
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 100
dev = torch.device("cuda", 0)
prob = bench.Problem(cfg, n, dev, 1, 0)
A = prob.assemble()
A_sys, b, M = prob.system(A)
torch.cuda.synchronize()
ts = []
for k in range(6):
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    x, info = cg(A_sys, b, M=M, atol=0.0, rtol=0.0, maxit=iters, returninfo=True)
    s1.record()
    torch.cuda.synchronize()
    ts.append(s0.elapsed_time(s1) / max(info["niter"], 1))
ts = ts[1:]
gdof = A_sys.shape[0]
bytes_it = 12 * A_sys.nnz + 8 * (gdof + 1) + 104 * gdof
print(f"cfg {cfg} n {n} kernel {os.environ.get('FB2_SPMV_KERNEL', 'tma')}: {min(ts):.4f} ms/it (med {statistics.median(ts):.4f}), "
      f"{1e3 / min(ts):.1f} it/s, algorithmic {bytes_it / min(ts) / 1e6:.0f} GB/s, x checksum {float(x.sum()):.12e}", flush=True)
