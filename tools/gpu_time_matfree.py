"""Time the fused matrix-free product `form @ u` (no A, no K_e) against the K_e-based product and the assembled SpMV.
   python tools/gpu_time_matfree.py [config] [n]"""
import os
import statistics
import sys

sys.path.insert(0, os.getcwd())
import torch
import bench
from fealpy_b200.fem import BilinearForm, ScalarDiffusionIntegrator, ScalarMassIntegrator

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else bench.CONFIGS[cfg]["n"]
dev = torch.device("cuda", 0)


def timed(fn, reps=12):
    ev = lambda: torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        s0, s1 = ev(), ev()
        s0.record()
        out = fn()
        s1.record()
        torch.cuda.synchronize()
        ts.append(s0.elapsed_time(s1))
    return min(ts[2:]), statistics.median(ts[2:]), out


prob = bench.Problem(cfg, n, dev, 1, 0)
space = prob.sspace
q = 3 if cfg == 1 else None
form = BilinearForm(space).add_integrator(ScalarDiffusionIntegrator(q=q)).add_integrator(ScalarMassIntegrator(q=q))
gdof = space.number_of_global_dofs()
u = torch.rand(gdof, dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(3))
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
w = form @ u
ev1.record()
torch.cuda.synchronize()
first = ev0.elapsed_time(ev1)
mn, med, w = timed(lambda: form @ u)
NC, L = space.cell_to_dof().shape
TD = space.mesh.TD
b_alg = 4 * NC * (TD + 1) + 4 * NC * L + 8 * TD * space.mesh.node.shape[0] + 16 * gdof      # cell + cell2dof + node + u + v
print(f"cfg {cfg} n {n} matfree {form.last_matfree}: first (adjacency build incl.) {first:.2f} ms, min {mn:.3f} med {med:.3f} ms  "
      f"NC {NC} gdof {gdof}  compulsory {b_alg / 1e9:.3f} GB -> {b_alg / mn / 1e6:.0f} GB/s", flush=True)
if cfg == 2 and n <= 64 or cfg != 2:
    fk = BilinearForm(space, assembly_path="gather").add_integrator(ScalarDiffusionIntegrator(q=q)).add_integrator(ScalarMassIntegrator(q=q))
    mn2, med2, w2 = timed(lambda: fk @ u, 5)
    print(f"   K_e-based product ({fk.last_matfree}): min {mn2:.3f} ms   max|diff| {float((w - w2).abs().max()):.3e}", flush=True)
form2 = BilinearForm(space).add_integrator(ScalarDiffusionIntegrator(q=q)).add_integrator(ScalarMassIntegrator(q=q))
A = form2.assembly()
mn3, med3, z = timed(lambda: A @ u)
print(f"   assembled SpMV: min {mn3:.3f} med {med3:.3f} ms  nnz {A.nnz}   max|matfree - A u| {float((w - z).abs().max()):.3e} (max|A u| {float(z.abs().max()):.3e})",
      flush=True)
