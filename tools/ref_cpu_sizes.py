"""Size sweep of the UNMODIFIED reference (baseline/_ref, numpy backend) on the host cores: shows that its assembled
nnz/s is flat in the mesh size (so a bounded sample extrapolates to the full configs) and that its CG rate is not.
Writes profiles/r02_reference_cpu_sizes.json.   python tools/ref_cpu_sizes.py [config:n,n,...] ..."""
import json
import os
import platform
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

plan = {2: [8, 16, 24, 32, 48], 1: [128, 256, 512], 3: [64, 128, 256], 4: [16, 32, 48]}
if len(sys.argv) > 1:
    plan = {}
    for a in sys.argv[1:]:
        c, ns = a.split(":")
        plan[int(c)] = [int(v) for v in ns.split(",")]
rows = []
for cfg, ns in plan.items():
    for n in ns:
        m = bench.cpu_measure(cfg, n, 1, 0, 5)
        rows.append(dict(config=cfg, n=n, gdof=m["gdof"], nnz_per_s=m["value"], cg_iters_per_s=m["cg_iters_per_s"],
                         s_per_step=m["ms_per_step"] / 1e3, kind=m["kind"]))
        print(rows[-1], flush=True)
out = dict(host=platform.processor() or platform.machine(), nproc=os.cpu_count(), blas_threads=bench.host_threads(),
           where="build container (not the GPU box's host)", rows=rows)
with open(os.path.join(ROOT, "profiles", "r02_reference_cpu_sizes.json"), "w") as f:
    json.dump(out, f, indent=1)
