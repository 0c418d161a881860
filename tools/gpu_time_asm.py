"""Time the warm assembly() of a config (CUDA events, best / median of N) and check it against the v4 kernel's values.
   python tools/gpu_time_asm.py [config] [n] ; knobs through the environment (FB2_ASM_KERNEL, FB2_ASM5_CAP, FB2_ASM5_THREADS)"""
import os
import statistics
import sys

sys.path.insert(0, os.getcwd())
import torch
import bench

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else bench.CONFIGS[cfg]["n"]
dev = torch.device("cuda", 0)
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    prob = bench.Problem(cfg, n, dev, 1, 0)
    if os.environ.get("FB2_ASM_PATH"):                 # force 'gather' / 'coo' on a form the fused kernel would take
        prob.bform.assembly_path = os.environ["FB2_ASM_PATH"]
    A = prob.assemble()
    torch.cuda.synchronize()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    ts = []
    for k in range(12):
        s0, s1 = ev(), ev()
        s0.record()
        A = prob.assemble()
        s1.record()
        torch.cuda.synchronize()
        ts.append(s0.elapsed_time(s1))
    ts = ts[2:]
    from fealpy_b200.fem.bilinear_form import symbolic_pattern
    sym = symbolic_pattern(prob.sspace)
    extra = ""
    if sym.get("asm5"):
        pl = sym["asm5"]
        extra = f" ntile {pl['ntile']} nbatch {pl['nbatch']} fill {sym['NC'] * sym['L'] / 32 / pl['nbatch']:.3f} acc_stride {pl['acc_stride']}"
    elif sym.get("asm4"):
        pl = sym["asm4"]
        extra = f" ntile {pl['ntile']} nbatch {pl['nbatch']} fill {sym['NC'] * sym['L'] / 32 / pl['nbatch']:.3f}"
    chk = float(A.values.sum()), float(A.values.abs().sum())
    print(f"cfg {cfg} n {n} kernel {getattr(prob.bform, 'last_kernel', prob.bform.last_path)} cap {os.environ.get('FB2_ASM5_CAP', '-')} "
          f"thr {os.environ.get('FB2_ASM5_THREADS', '-')}: min {min(ts):.3f} med {statistics.median(ts):.3f} ms  nnz {A.nnz}{extra}  "
          f"checksum {chk[0]:.12e} {chk[1]:.12e}", flush=True)
