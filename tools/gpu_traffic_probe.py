"""One warm assembly + a 4-iteration CG of a config, for an ncu pass that records per-launch DRAM bytes
(tools/make_traffic.py turns the CSV into profiles/r02_traffic.json, which bench.py reports as roofline.traffic).
   python tools/gpu_traffic_probe.py [config]"""
import os
import sys

sys.path.insert(0, os.getcwd())
import torch
import bench
from fealpy_b200.solver import cg

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda", 0)
prob = bench.Problem(cfg, bench.CONFIGS[cfg]["n"], dev, 1, 0)
A = prob.assemble()
A = prob.assemble()
torch.cuda.synchronize()
print("MARK assembly-start", flush=True)
A = prob.assemble()
torch.cuda.synchronize()
A_sys, b, M = prob.system(A)
x, info = cg(A_sys, b, M=M, atol=0.0, rtol=0.0, maxit=4, returninfo=True)
torch.cuda.synchronize()
print("done", info)
