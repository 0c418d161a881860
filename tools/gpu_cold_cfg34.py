"""Where the set-up time of configs 3 and 4 goes (wall clock, synchronised after every phase), two passes per config."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
import bench
from fealpy_b200.fem import bilinear_form as bfm

dev = torch.device("cuda", 0)
torch.zeros(1, device=dev); torch.cuda.synchronize()

def phase(name, fn, log):
    torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize()
    log.append((name, (time.perf_counter() - t0) * 1e3)); return out

for cfg in (4, 3):
    for rep in range(2):
        log = []
        prob = phase("Problem()", lambda: bench.Problem(cfg, bench.CONFIGS[cfg]["n"], dev, 1, 0), log)
        phase("cell_to_dof", lambda: prob.c2d(), log)
        phase("symbolic_pattern(scalar space)", lambda: bfm.symbolic_pattern(prob.sspace), log)
        if cfg == 4:
            phase("tensor_pattern", lambda: bfm.tensor_pattern(prob.bf_space), log)
        A = phase("first assembly", lambda: prob.assemble(), log)
        A = phase("second assembly", lambda: prob.assemble(), log)
        if cfg == 4:
            from fealpy_b200.fem import DirichletBC
            from fealpy_b200.sparse import CSRTensor
            gdof = A.shape[0]
            b = phase("b = A @ v", lambda: A @ (torch.sin(0.37 * torch.arange(gdof, dtype=torch.float64, device=dev)) + 1.5), log)
            ip = phase("interpolation_points", lambda: prob.sspace.interpolation_points(), log)
            flag = (ip[:, 0] < 1e-12).repeat_interleave(3)
            bc = phase("DirichletBC()", lambda: DirichletBC(prob.bf_space, gd=torch.zeros(gdof, dtype=torch.float64, device=dev), threshold=flag), log)
            A2, b2 = phase("bc.apply", lambda: bc.apply(A, b), log)
            d = phase("A2.diags()", lambda: A2.diags(), log)
        print(f"cfg {cfg} pass {rep}: " + " | ".join(f"{n} {t:.1f}" for n, t in log), flush=True)
        del prob, A
