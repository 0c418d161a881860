"""Where the cold (first-assembly) time of a config goes: wall clock with a device synchronize after every phase, plus the
allocator's segment count (each new segment is a cudaMalloc).  Run twice in one process: the second pass finds the freed
blocks of the first in torch's caching allocator, so the difference is allocation cost.
   python tools/gpu_cold_breakdown.py [config] [n]"""
import os
import sys
import time

sys.path.insert(0, os.getcwd())
import torch
import bench
from fealpy_b200.fem import bilinear_form as bfm

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else bench.CONFIGS[cfg]["n"]
dev = torch.device("cuda", 0)
torch.zeros(1, device=dev)
torch.cuda.synchronize()


def segs():
    return torch.cuda.memory_stats(dev)["segment.all.allocated"]


def phase(name, fn, log):
    torch.cuda.synchronize()
    s0, t0 = segs(), time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    log.append((name, (time.perf_counter() - t0) * 1e3, segs() - s0))
    return out


for rep in range(2):
    log = []
    prob = phase("mesh.from_box + spaces", lambda: bench.Problem(cfg, n, dev, 1, 0), log)
    phase("edges (topology)", lambda: prob.mesh.edge, log)
    phase("cell_to_dof", lambda: prob.c2d(), log)
    phase("symbolic_pattern", lambda: bfm.symbolic_pattern(prob.sspace), log)
    if prob.bform._plan_fused() is not None:
        phase("asm4_plan (schedule)", lambda: bfm.asm4_plan(prob.sspace), log)
    A = phase("first assembly", lambda: prob.assemble(), log)
    phase("second assembly", lambda: prob.assemble(), log)
    phase("spmv plan + colz", lambda: A.spmv_plan(), log)
    tot = sum(t for _, t, _ in log[:6])
    print(f"cfg {cfg} n {n} pass {rep}: cold total (to first assembly) {tot:.2f} ms, reserved {torch.cuda.memory_reserved(dev) / 2**30:.2f} GiB")
    for name, t, s in log:
        print(f"   {name:28s} {t:8.2f} ms   new segments {s}")
    del prob, A
