#!/usr/bin/env python
"""Benchmark of the hot path `BilinearForm(...).assembly() -> CSRTensor -> cg` on synthetic `from_box` meshes.

    python bench.py --gpus 1 --steps K --warmup W [--config 1..5]      # our CUDA path
    python bench.py --impl reference --steps K --warmup W              # the UNMODIFIED reference on the host cores

Configs (BASELINE.json `configs`, sizes of SURVEY.md section 8; default 2 = the one the metric is quoted on):
  1  tri 1024^2  P1      diffusion(q=3) + mass(q=3)                       + CG
  2  tet 128^3   P2      diffusion + mass (q=5)                           + CG      [x-slabs of a (128 N)x128x128 box at N GPUs]
  3  tri 1024^2  P3      variable-coefficient diffusion (q=6) + mass(q=6) + CG
  4  tet 128^3   P1 x 3  linear elasticity (q=4), Dirichlet face x=0      + Jacobi-preconditioned CG
  5  tet 322^3   P2      diffusion + mass, ONE box row-partitioned over the N >= 2 GPUs (config 2's kernels)

One "step" = one `assembly()` of the form (symbolic pattern cached per space = "warm") followed by a fixed number of CG
iterations.  `value` = assembled nnz/s with all inputs resident in HBM (CUDA events on the launching stream, max over
ranks); `cg` reports iterations/s; `cold` the one-time costs a single assembly on a fresh space pays; `e2e` repeats the
step through the public API with HOST (pinned) inputs, the H2D / D2H copies inside the timed region.  At N > 1 every rank
first runs the hardware parity check of fealpy_b200.parallel.verify (`verify`), and the run fails if it does.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

_emit = print
METRIC = "assembled_nnz_per_s"
UNIT = "nnz/s"

CONFIGS = {
    1: dict(mesh="tri", n=1024, p=1, what="tri P1 Poisson: ScalarDiffusionIntegrator(q=3) + ScalarMassIntegrator(q=3)", cpu_rate=6.0e5),
    2: dict(mesh="tet", n=128, p=2, what="tet P2 Poisson: ScalarDiffusionIntegrator + ScalarMassIntegrator (q=5)", cpu_rate=2.6e5),
    3: dict(mesh="tri", n=1024, p=3, what="tri P3: ScalarDiffusionIntegrator(coef=kappa(x), q=6) + ScalarMassIntegrator(q=6)",
            cpu_rate=9.0e5),
    4: dict(mesh="tet", n=128, p=1, what="tet P1x3 LinearElasticityIntegrator(E=1, nu=0.3, q=4), Dirichlet face x=0, Jacobi CG",
            cpu_rate=5.5e5),
    5: dict(mesh="tet", n=322, p=2, what="tet P2 Poisson: ScalarDiffusionIntegrator + ScalarMassIntegrator (q=5)", cpu_rate=2.6e5),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--n", type=int, default=0, help="cells per box edge (0 = the config's size)")
    ap.add_argument("--cg-iters", type=int, default=100)
    ap.add_argument("--cpu-n", type=int, default=0, help="box edge of the CPU sample (0 = sized from --cpu-budget-s)")
    ap.add_argument("--cpu-budget-s", type=float, default=0.0, help="seconds of CPU work (0 = 20 s beside our arm, 200 s for --impl reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--no-existing-gpu", action="store_true")
    ap.add_argument("--dist-mode", default=os.environ.get("FB2_DIST", "auto"), choices=["auto", "nccl", "peer"])
    ap.add_argument("--existing-gpu-child", action="store_true", help=argparse.SUPPRESS)
    return ap.parse_args()


def workload(cfg_id, n, cg_iters, world):
    c = CONFIGS[cfg_id]
    box = f"TriangleMesh.from_box n={n}" if c["mesh"] == "tri" else f"TetrahedronMesh.from_box n={n}"
    if cfg_id == 5:
        box += f" (one {n}^3 box over {world} GPUs)"
    elif world > 1:
        box += " per GPU"
    return f"config {cfg_id}: {c['what']} assembly to CSR + {cg_iters} CG iterations, {box}"


def config_dict(cfg_id, n, cg_iters, world):
    par = "single GPU" if world == 1 else (f"{world} x-slabs of one {n}^3 box" if cfg_id == 5 else
                                           f"{world} x-slabs of one {n * world}x{n}x{n} box") + ": owned-row assembly + halo-exchange CG"
    return {"workload": workload(cfg_id, n, cg_iters, world), "config_id": cfg_id,
            "l2_policy": "inputs larger than L2 (multi-GB arrays per step)" if cfg_id in (2, 4, 5) else
                         "inputs + outputs exceed L2 per step except config 1 (164 MB, flushed by the CG vectors between assemblies)",
            "parallelism": par}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(cfg_id):
    """per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernels from the committed
    `ncu --set full` capture of this workload (profiles/r02_traffic.json), or {}"""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            return json.load(f).get(str(cfg_id), {})
    except Exception:
        return {}


class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed region.  In-process NVML (a 30 ms thread) when
    nvidia_ml_py is importable -- an external `nvidia-smi -lms` loop was seen to stall this process's CUDA calls by
    ~30 ms per step on some boxes -- else nvidia-smi."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.path = index, None, None
        self.thread, self.stop_flag, self.samples, self.skip = None, None, [], 0

    def _nvml_loop(self, nv, h):
        R = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
             "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        while not self.stop_flag.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
                bits = (getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons)(h)
                self.samples.append((float(sm), float(mx), {k for k, v in R.items() if bits & v}))
            except Exception:
                pass
            self.stop_flag.wait(float(os.environ.get("FB2_BENCH_SAMPLER_S", "0.03")))      # 30 ms: short configs (config 1: 8 ms per step) still get samples

    def start(self):
        try:
            import threading
            import pynvml as nv
            import torch
            nv.nvmlInit()
            try:
                h = nv.nvmlDeviceGetHandleByUUID("GPU-" + str(torch.cuda.get_device_properties(self.index).uuid))
            except Exception:
                h = nv.nvmlDeviceGetHandleByIndex(self.index)
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            self.stop_flag = threading.Event()
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "250",
                                          "-i", str(self.index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def mark(self):
        """start of the timed region: forget what was sampled before"""
        self.samples.clear()
        if self.proc is not None and self.path:
            try:
                self.skip = sum(1 for _ in open(self.path))
            except Exception:
                self.skip = 0

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            if self.samples:
                out.update(sm_mhz=statistics.median(x[0] for x in self.samples), sm_max_mhz=max(x[1] for x in self.samples),
                           reasons=sorted(set().union(*(x[2] for x in self.samples))), samples=len(self.samples), source="nvml")
            return out
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for ln, line in enumerate(open(self.path)):
                if ln < self.skip:
                    continue
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# --------------------------------------------------------------------------------------
# CPU arm: the unmodified reference (baseline/_ref) on the host cores; the numpy oracle port only if it is absent
# --------------------------------------------------------------------------------------
def kappa_np(p):
    import numpy as np
    return 1.0 + 0.5 * np.sin(2 * np.pi * p[..., 0]) * np.cos(2 * np.pi * p[..., 1])


def host_threads():
    try:
        import threadpoolctl
        th = [d.get("num_threads", 1) for d in threadpoolctl.threadpool_info() if d.get("user_api") == "blas"]
        return max(th) if th else 1
    except Exception:
        return os.cpu_count() or 1


def est_nnz(cfg_id, n):
    """nnz of the config's matrix at box edge n (closed forms of SURVEY.md section 8 scaled; used only to size samples)"""
    full = {1: (7_346_177, 1024, 2), 2: (484_609_025, 128, 3), 3: (160_462_849, 1024, 2), 4: (286_222_473, 128, 3),
            5: (7_693_153_161, 322, 3)}[cfg_id]
    return full[0] * (n / full[1]) ** full[2]


def sample_size(cfg_id, budget_s, steps):
    """largest box edge whose `steps` reference steps fit the CPU budget"""
    c = CONFIGS[cfg_id]
    lo = 4
    grid = [4, 6, 8, 10, 12, 14, 16, 20, 24, 28, 32, 40, 48, 64, 96, 128, 192, 256, 384, 512, 768, 1024]
    best = lo
    for n in grid:
        if n > c["n"]:
            break
        t = est_nnz(cfg_id, n) / c["cpu_rate"] * 1.25          # + CG / setup share
        if t * max(steps, 1) <= budget_s:
            best = n
    return best


class RefProblem:
    """the config on the reference's own objects (numpy backend, or the reference's pytorch backend on `device`)"""

    def __init__(self, cfg_id, n, backend="numpy", device=None):
        from baseline import ref_loader
        ref_loader.install()
        from fealpy.backend import backend_manager as bm
        bm.set_backend(backend)
        from fealpy.mesh import TriangleMesh, TetrahedronMesh
        from fealpy.functionspace import LagrangeFESpace, TensorFunctionSpace
        from fealpy.fem import BilinearForm, ScalarDiffusionIntegrator, ScalarMassIntegrator, LinearElasticityIntegrator
        from fealpy.decorator import cartesian
        self.bm, self.cfg_id = bm, cfg_id
        c = CONFIGS[cfg_id]
        kw = {} if device is None else {"device": device}
        if c["mesh"] == "tri":
            mesh = TriangleMesh.from_box([0, 1, 0, 1], n, n, **kw)
        else:
            mesh = TetrahedronMesh.from_box([0, 1, 0, 1, 0, 1], n, n, n, **kw)
        self.mesh = mesh
        space = LagrangeFESpace(mesh, c["p"])
        NC = mesh.number_of_cells()
        # No `splitter=`: the reference's chunked element loop (fem/form.py:158-188) yields K_e chunks against the FULL
        # cell2dof in BilinearForm._scalar_assembly (fem/bilinear_form.py:69: "operands could not be broadcast ...
        # (663552,10,1) and (262144,10,10)" at n = 48), so the unmodified reference can only assemble unchunked; the
        # sample size is therefore also bounded by host memory (gphi = NC x NQ x ldof x GD doubles).
        split = None
        self.BilinearForm = BilinearForm
        if cfg_id == 4:
            from fealpy.material.elastic_material import LinearElasticMaterial
            space = TensorFunctionSpace(space, shape=(-1, 3))
            mat = LinearElasticMaterial("m", elastic_modulus=1.0, poisson_ratio=0.3, hypo="3D")
            self.groups = [([LinearElasticityIntegrator(mat, q=4)], split)]
        elif cfg_id == 3:
            if backend == "numpy":
                kap = cartesian(lambda p: kappa_np(p))
            else:
                import torch
                kap = cartesian(lambda p: 1.0 + 0.5 * torch.sin(2 * torch.pi * p[..., 0]) * torch.cos(2 * torch.pi * p[..., 1]))
            self.groups = [([ScalarDiffusionIntegrator(coef=kap, q=6)], split), ([ScalarMassIntegrator(q=6)], None)]
        elif cfg_id == 1:
            self.groups = [([ScalarDiffusionIntegrator(q=3)], split), ([ScalarMassIntegrator(q=3)], None)]
        else:
            self.groups = [([ScalarDiffusionIntegrator()], split), ([ScalarMassIntegrator()], None)]
        self.space = space
        self.NC = NC

    def assemble(self):
        bf = self.BilinearForm(self.space)
        for ints, split in self.groups:
            bf.add_integrator(*ints, splitter=split) if split else bf.add_integrator(*ints)
        return bf.assembly()

    def rhs_and_system(self, A):
        """(A_sys, b, M): config 4 constrains the face x = 0 through the reference's DirichletBC and uses its Jacobi
        preconditioner; the others solve A x = A 1 unpreconditioned"""
        bm = self.bm
        gdof = A.shape[0]
        kw = {} if self.mesh.device in (None, "cpu") else {"device": self.mesh.device}
        if self.cfg_id != 4:
            return A, A @ bm.ones(gdof, dtype=bm.float64, **kw), None
        # elasticity: A 1 = 0 (rigid translation), so the right-hand side comes from a non-rigid vector
        b = A @ (bm.sin(0.37 * bm.arange(gdof, dtype=bm.float64, **kw)) + 1.5)
        from fealpy.fem import DirichletBC
        from fealpy.sparse import CSRTensor
        ip = self.space.interpolation_points()
        sflag = ip[:, 0] < 1e-12
        flag = bm.repeat(sflag, 3) if hasattr(bm, "repeat") else None
        bc = DirichletBC(self.space, gd=bm.zeros(gdof, dtype=bm.float64), threshold=flag)
        A2, b2 = bc.apply(A, b)
        d = A2.diags()
        return A2, b2, CSRTensor(d.crow, d.col, 1.0 / d.values, A2.shape)


def ref_step(prob, cg_iters):
    from fealpy.solver import cg
    t0 = time.perf_counter()
    A = prob.assemble()
    t1 = time.perf_counter()
    A2, b, M = prob.rhs_and_system(A)
    t2 = time.perf_counter()
    x, info = cg(A2, b, M=M, atol=0.0, rtol=0.0, maxit=cg_iters, returninfo=True)
    t3 = time.perf_counter()
    return A.nnz, t1 - t0, info["niter"], t3 - t2, A.shape[0]


def oracle_step(n, p, cg_iters):
    """fallback when baseline/_ref is absent: the numpy restatement of the reference path (config 2 shape only)"""
    import numpy as np
    from oracle import fem_oracle as O
    node, cell = O.tet_from_box([0, 1, 0, 1, 0, 1], n, n, n)
    mesh = O.Mesh(node, cell)
    c2d = mesh.cell_to_ipoint(p)
    gdof = mesh.number_of_global_ipoints(p)
    t0 = time.perf_counter()
    crow, col, val = O.assemble([(O.diffusion_element(mesh, p), c2d), (O.mass_element(mesh, p), c2d)], gdof)
    t1 = time.perf_counter()
    b = O.csr_matvec(crow, col, val, np.ones(gdof))
    t2 = time.perf_counter()
    x, info = O.cg(lambda v: O.csr_matvec(crow, col, val, v), b, atol=0.0, rtol=0.0, maxit=cg_iters)
    t3 = time.perf_counter()
    return len(val), t1 - t0, info["niter"], t3 - t2, gdof


def cpu_measure(cfg_id, n, steps, warmup, cg_iters):
    """-> dict(kind, nnz/s, it/s, sample description) from `steps` timed reference steps at box edge n"""
    from baseline import ref_loader
    if ref_loader.available():
        kind = "reference"
        for _ in range(warmup):                                   # warm-up steps run the same code on an 8-cell-per-edge box
            ref_step(RefProblem(cfg_id, min(8, n)), 3)
        prob = RefProblem(cfg_id, n)
        run = lambda: ref_step(prob, cg_iters)
    else:
        kind = "port"
        if cfg_id not in (2, 5):
            return None
        for _ in range(min(warmup, 1)):
            oracle_step(min(6, n), 2, 3)
        run = lambda: oracle_step(n, 2, cg_iters)
    t_asm = t_cg = 0.0
    nnz = its = gdof = 0
    for _ in range(steps):
        nz, ta, it, tc, gdof = run()
        nnz += nz; t_asm += ta; its += it; t_cg += tc
    c = CONFIGS[cfg_id]
    what = ("unmodified FEALPy 3.4.0 (baseline/_ref), numpy backend" if kind == "reference" else "numpy oracle port of the reference path")
    sample = (f"{what}: {c['what']} on from_box n={n} (gdof {gdof}, nnz {nnz // max(steps, 1)}), {steps} step(s) of "
              f"assembly() [{t_asm / max(steps, 1):.1f} s each] + {cg_iters} cg iterations; the full config is n={c['n']}: the "
              f"reference's nnz/s is flat in n (2.6e5 / 2.3e5 / 3.2e5 / 2.7e5 / 2.9e5 at n = 8 / 16 / 24 / 32 / 40 for config 2, profiles/r02_reference_cpu_sizes.json), "
              f"so the sample rate is the extrapolation to the full size; its CG rate is NOT size-independent (see cg_same_matrix)")
    return dict(kind=kind, value=nnz / t_asm, cg_iters_per_s=its / t_cg if t_cg > 0 else None, n=n, gdof=gdof, sample=sample,
                ms_per_step=1e3 * (t_asm + t_cg) / max(steps, 1), cores=host_threads())


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg_id = args.config
    n_full = args.n or CONFIGS[cfg_id]["n"]
    budget = args.cpu_budget_s or 200.0
    n = args.cpu_n or sample_size(cfg_id, budget, args.steps)
    cg_iters = min(args.cg_iters, 10)
    m = cpu_measure(cfg_id, min(n, n_full), args.steps, args.warmup, cg_iters)
    if m is None:
        _emit(json.dumps({"impl": "reference", "unavailable": "baseline/_ref is not installed and the oracle port covers config 2 only"}))
        return
    line = {
        "impl": "reference", "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(cfg_id, n_full, args.cg_iters, args.gpus),
        "cg": {"iters_per_s": m["cg_iters_per_s"], "iters_per_step": cg_iters, "gdof": m["gdof"],
               "note": "CG on the SAMPLE matrix (cache-resident at small n): compare with cpu_baseline.cg_same_matrix of our arm instead"},
        "cpu_baseline": {"value": m["value"], "unit": UNIT, "cores": m["cores"], "kind": m["kind"], "sample": m["sample"],
                         "cores_note": "threads OpenBLAS may use inside einsum; the path is otherwise single-threaded numpy",
                         "cg_iters_per_s": m["cg_iters_per_s"]},
        "e2e": {"value": m["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(json.dumps(line))


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
class Problem:
    """the config on fealpy_b200 objects; `solver` is persistent (no allocation inside the timed region)"""

    def __init__(self, cfg_id, n, dev, world, rank, dist_mode="auto"):
        import torch
        from fealpy_b200.mesh import TriangleMesh, TetrahedronMesh
        from fealpy_b200.functionspace import LagrangeFESpace, TensorFunctionSpace
        from fealpy_b200.fem import (BilinearForm, ScalarDiffusionIntegrator, ScalarMassIntegrator, LinearElasticityIntegrator)
        self.cfg_id, self.dev, self.world, self.rank = cfg_id, dev, world, rank
        c = CONFIGS[cfg_id]
        self.part = None
        if world > 1:
            if cfg_id not in (2, 5):
                raise SystemExit("multi-GPU runs use config 2 (weak scaling) or 5 (one 322^3 box)")
            from fealpy_b200.parallel import SlabProblem
            nx = n * world if cfg_id == 2 else n
            sp = SlabProblem([0, world if cfg_id == 2 else 1, 0, 1, 0, 1], nx, n, n, c["p"], world, rank, device=dev)
            self.mesh, self.space, self.part = sp.mesh, sp.space, sp.part
        elif c["mesh"] == "tri":
            self.mesh = TriangleMesh.from_box([0, 1, 0, 1], n, n, device=dev)
            self.space = LagrangeFESpace(self.mesh, c["p"])
        else:
            self.mesh = TetrahedronMesh.from_box([0, 1, 0, 1, 0, 1], n, n, n, device=dev)
            self.space = LagrangeFESpace(self.mesh, c["p"])
        self.sspace = self.space
        bf_space = self.space
        if cfg_id == 4:
            from fealpy_b200.material import LinearElasticMaterial
            bf_space = TensorFunctionSpace(self.space, shape=(-1, 3))
            ints = [[LinearElasticityIntegrator(LinearElasticMaterial("m", elastic_modulus=1.0, poisson_ratio=0.3, hypo="3D"), q=4)]]
        elif cfg_id == 3:
            def kappa(p):
                return 1.0 + 0.5 * torch.sin(2 * torch.pi * p[..., 0]) * torch.cos(2 * torch.pi * p[..., 1])
            kappa.coordtype = "cartesian"
            ints = [[ScalarDiffusionIntegrator(coef=kappa, q=6)], [ScalarMassIntegrator(q=6)]]
        elif cfg_id == 1:
            ints = [[ScalarDiffusionIntegrator(q=3)], [ScalarMassIntegrator(q=3)]]
        else:
            ints = [[ScalarDiffusionIntegrator()], [ScalarMassIntegrator()]]
        self.bf_space = bf_space
        self.bform = BilinearForm(bf_space)
        for g in ints:
            self.bform.add_integrator(*g)
        self.values = None
        self.dist_mode = dist_mode

    def c2d(self):
        return self.sspace.cell_to_dof()

    def assemble(self):
        A = self.bform.assembly(out=self.values)
        self.values = A.values                     # every later assembly reuses this buffer
        return A

    def system(self, A):
        """(A_sys, b, M) as solved by CG; built once (the BC / preconditioner setup is reported as `bc_ms`, not part of `value`)"""
        import torch
        gdof = A.shape[0]
        if self.cfg_id != 4:
            return A, A @ torch.ones(gdof, dtype=torch.float64, device=self.dev), None
        # elasticity: A 1 = 0 (rigid translation), so the right-hand side comes from a non-rigid vector
        b = A @ (torch.sin(0.37 * torch.arange(gdof, dtype=torch.float64, device=self.dev)) + 1.5)
        from fealpy_b200.fem import DirichletBC
        from fealpy_b200.sparse import CSRTensor
        sflag = self.sspace.interpolation_points()[:, 0] < 1e-12
        flag = sflag.repeat_interleave(3)
        bc = DirichletBC(self.bf_space, gd=torch.zeros(gdof, dtype=torch.float64, device=self.dev), threshold=flag)
        A2, b2 = bc.apply(A, b)
        d = A2.diags()
        return A2, b2, CSRTensor(d.crow, d.col, 1.0 / d.values, A2.shape)


def alg_bytes(cfg_id, NC, NN, L, nnz, gdof, TD, GD, NQ_coef):
    """algorithmic bytes.  `survey`: cell + cell2dof + node + coef + values + col + crow, what ONE
    cold assembly() must move.  `warm`: what the numeric phase must move when the pattern is cached and shared (the
    default, BilinearForm(share_pattern=True)): cell + node + coef + values -- the figure `frac` uses."""
    coef = 8 * NC * NQ_coef
    survey = 4 * NC * (TD + 1) + 4 * NC * L + 8 * GD * NN + coef + 12 * nnz + 8 * (gdof + 1)
    warm = 4 * NC * (TD + 1) + 8 * GD * NN + coef + 8 * nnz
    it = 12 * nnz + 8 * (gdof + 1) + 104 * gdof
    return survey, warm, it


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from fealpy_b200.fem.bilinear_form import symbolic_pattern
    from fealpy_b200.solver import cg

    cfg_id = args.config
    if world > 1 and cfg_id not in (2, 5):
        raise SystemExit("multi-GPU runs use --config 2 (weak scaling) or --config 5")
    if cfg_id == 5 and world < 2:
        raise SystemExit("config 5 (322^3, 200 M cells) is the multi-GPU configuration: run it with --gpus 2/4/8")
    c = CONFIGS[cfg_id]
    n = args.n or c["n"]
    stream = torch.cuda.Stream(device=dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    verify = None
    with torch.cuda.stream(stream):
        # ---- hardware parity of the multi-GPU data plane, before anything is timed -------------------------------
        if world > 1 and not args.no_verify:
            from fealpy_b200.parallel import verify_slab, make_dist_solver
            verify = verify_slab(world, rank, dev, solver_factory=lambda A_, part_, group=None: make_dist_solver(A_, part_, mode=args.dist_mode))
            if not verify["ok"]:
                if rank == 0:
                    sys.stderr.write("bench: multi-GPU parity check FAILED: " + json.dumps(verify) + "\n")
                dist.destroy_process_group()
                sys.exit(3)
        # ---- setup (not timed as part of the step; reported as cold costs) ----------------------------------------
        e0, e1, e2, e3 = ev(), ev(), ev(), ev()
        e0.record()
        prob = Problem(cfg_id, n, dev, world, rank, args.dist_mode)
        mesh, space, part, bform = prob.mesh, prob.sspace, prob.part, prob.bform
        c2d = prob.c2d()
        e1.record()
        sym = symbolic_pattern(space)
        e2.record()
        A = prob.assemble()
        e3.record()
        torch.cuda.synchronize(dev)
        t_mesh, t_sym, t_first = e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3)
        nnz, gdof, NC, NN = A.nnz, A.shape[0], mesh.number_of_cells(), mesh.number_of_nodes()
        L = c2d.shape[1]
        eb0, eb1 = ev(), ev()
        eb0.record()
        A_sys, b, M = prob.system(A)
        eb1.record()
        torch.cuda.synchronize(dev)
        t_bc = eb0.elapsed_time(eb1)
        own_mask = None
        if part is not None:       # count only owned rows (halo rows are scratch)
            cr = A.crow
            (a0, a1), (b0, b1) = part.own_nodes, part.own_edges
            nnz = int(cr[a1] - cr[a0]) + int(cr[b1] - cr[b0])
            own_mask = torch.zeros(gdof, dtype=torch.bool, device=dev)
            own_mask[a0:a1] = True
            own_mask[b0:b1] = True
            from fealpy_b200.parallel import make_dist_solver
            dsolver = make_dist_solver(A, part, mode=args.dist_mode)

        def solve(A_, rhs, maxit, atol=0.0, rtol=0.0):
            if part is None:
                return cg(A_, rhs, M=M, atol=atol, rtol=rtol, maxit=maxit, returninfo=True)
            return dsolver.solve(rhs, atol=atol, rtol=rtol, maxit=maxit, check_every=min(maxit, 25))

        def step():
            s0, s1, s2 = ev(), ev(), ev()
            s0.record()
            A_ = prob.assemble()
            s1.record()
            x, info = solve(A_sys if cfg_id == 4 else A_, b, args.cg_iters)
            s2.record()
            return s0, s1, s2, info["niter"], x

        # the clock sampler starts BEFORE the warm-up: NVML (or nvidia-smi) initialisation was seen to stall this
        # process's CUDA calls for ~100 ms once, which must not land in the first timed step; samples taken during
        # the warm-up are dropped
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        for _ in range(max(args.warmup, 3)):
            step()
        barrier()
        sampler.mark()
        # a generation-2 pass of Python's cyclic GC takes 10-100 ms in a process that has imported torch; when it
        # fires between s0.record() and the kernel launch it shows up as "assembly time".  Collect now, keep the
        # collector off inside the timed regions.
        import gc
        gc.collect()
        gc.disable()
        t_wall0 = time.perf_counter()
        recs, last_x = [], None
        for _ in range(args.steps):
            r = step()
            recs.append(r[:4])
            last_x = r[4]
            del r
        barrier()
        t_wall = time.perf_counter() - t_wall0
        gc.enable()
        clocks = sampler.stop() if rank == 0 else {}
        t_asm = sum(r[0].elapsed_time(r[1]) for r in recs) * 1e-3
        t_cg = sum(r[1].elapsed_time(r[2]) for r in recs) * 1e-3
        iters = sum(r[3] for r in recs)

        # ---- a CONVERGED solve (reference tolerances), outside the timed region: the error of the returned solution -----
        xc, cinfo = solve(A_sys if cfg_id == 4 else A, b, 20000, atol=1e-12, rtol=1e-8)
        if cfg_id == 4:
            resid = b - A_sys @ xc
            conv = {"niter": int(cinfo["niter"]), "rel_residual": float(resid.norm() / b.norm())}
        else:
            d = (xc - 1.0)
            if own_mask is not None:
                d = d[own_mask]
            e_inf = d.abs().max().reshape(1)
            if world > 1:
                dist.all_reduce(e_inf, op=dist.ReduceOp.MAX)
            conv = {"niter": int(cinfo["niter"]), "x_err_vs_exact": float(e_inf), "residual": float(cinfo["residual"])}
        del xc

        # ---- e2e: host (pinned) inputs, copies inside the timed region ------------------------------
        e2e = None
        if not args.no_e2e:
            h_node, h_cell, h_c2d = (t.cpu().pin_memory() for t in (mesh.node, mesh.cell, c2d))
            h_b = b.cpu().pin_memory()
            h_vals = torch.empty(A.nnz, dtype=torch.float64).pin_memory()
            h_x = torch.empty(gdof, dtype=torch.float64).pin_memory()

            def e2e_step():
                s0, s1, s2 = ev(), ev(), ev()
                s0.record()
                mesh.node.copy_(h_node, non_blocking=True)
                mesh.cell.copy_(h_cell, non_blocking=True)
                c2d.copy_(h_c2d, non_blocking=True)
                A_ = prob.assemble()
                h_vals.copy_(A_.values, non_blocking=True)
                s1.record()
                b.copy_(h_b, non_blocking=True)
                x, info = solve(A_sys if cfg_id == 4 else A_, b, args.cg_iters)
                h_x.copy_(x, non_blocking=True)
                s2.record()
                return s0, s1, s2, info["niter"]
            for _ in range(2):
                e2e_step()
            barrier()
            gc.collect()
            gc.disable()
            er = [e2e_step() for _ in range(max(2, min(args.steps, 5)))]
            barrier()
            gc.enable()
            ta = sum(r[0].elapsed_time(r[1]) for r in er) * 1e-3
            tc = sum(r[1].elapsed_time(r[2]) for r in er) * 1e-3
            e2e = {"value": nnz * len(er) / ta, "unit": UNIT,
                   "h2d_bytes_per_step": int(h_node.nbytes + h_cell.nbytes + h_c2d.nbytes + h_b.nbytes),
                   "d2h_bytes_per_step": int(h_vals.nbytes + h_x.nbytes),
                   "cg_iters_per_s": sum(r[3] for r in er) / tc,
                   "note": "assembly: H2D node+cell+cell2dof -> assembly() -> D2H values; CG: H2D b -> cg -> D2H x"}
            del h_vals, h_x, h_node, h_cell, h_c2d

    # ---- reduce over ranks (max time) ----------------------------------------------------------------
    times = torch.tensor([t_asm, t_cg, t_wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        tot = torch.tensor([nnz, NC, part.n_owned], dtype=torch.int64, device=dev)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        nnz_glob, NC_ghosted, gdof_glob = (int(v) for v in tot)      # every rank also assembles one ghost layer of cells
        NC_glob = 6 * (n * world if cfg_id == 2 else n) * n * n       # cells of the one box the ranks share
        if e2e is not None:
            ev_ = torch.tensor([nnz * 1.0 / e2e["value"], 1.0 / e2e["cg_iters_per_s"]], dtype=torch.float64, device=dev)
            dist.all_reduce(ev_, op=dist.ReduceOp.MAX)
            e2e["value"] = nnz_glob / float(ev_[0])
            e2e["cg_iters_per_s"] = 1.0 / float(ev_[1])
    else:
        nnz_glob, NC_glob, gdof_glob, NC_ghosted = nnz, NC, gdof, NC
    t_asm, t_cg, t_wall = (float(v) for v in times)

    if rank == 0:
        peak, peak_src = peaks()
        traffic = measured_traffic(cfg_id)
        TD = mesh.TD
        GD = TD
        NQ_coef = 21 if cfg_id == 3 else 0
        ncomp = 3 if cfg_id == 4 else 1
        b_survey, b_warm, b_it = alg_bytes(cfg_id, NC, NN, L * ncomp, nnz, A.shape[0], TD, GD, NQ_coef)
        nsys, nnz_sys = (A_sys.shape[0], A_sys.nnz) if cfg_id == 4 else (gdof, A.nnz)
        b_it = 12 * nnz_sys + 8 * (nsys + 1) + (104 + (16 if M is not None else 0)) * nsys
        asm_gbs = b_warm * args.steps / t_asm / 1e9          # per GPU (rank 0's share of the bytes, max-over-ranks time)
        cg_gbs = b_it * iters / t_cg / 1e9
        path = bform.last_path
        asm_kernel = {"fused": "cell_geometry4_kernel + assemble_const_v6_kernel", "gather": "elem_* (K1) + assemble_from_ke_kernel",
                      "fused-elasticity-p1": "cell_gradients_kernel + assemble_from_ke_kernel<ETD=3>"}.get(path, path)
        asm_launches = {"fused": 2, "fused-elasticity-p1": 2}.get(path, 6)
        line = {
            "metric": METRIC, "value": nnz_glob * args.steps / t_asm, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * (t_asm + t_cg) / args.steps, "higher_is_better": True,
            "scaling": "weak" if cfg_id != 5 else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(cfg_id, n, args.cg_iters, world),
            "problem": {"NC": NC_glob, "NC_assembled_incl_ghost_layers": NC_ghosted, "gdof": gdof_glob, "nnz": nnz_glob, "assembly_path": path,
                        "pattern": "warm (symbolic pattern cached per space, shared by the returned matrices: share_pattern=True)"},
            "cg": {"iters_per_s": iters / t_cg, "iters_per_step": args.cg_iters, "ms_per_iter": 1e3 * t_cg / iters,
                   "preconditioner": "jacobi" if M is not None else None, "converged_solve": conv,
                   "dist_mode": getattr(dsolver, "mode", None) if part is not None else None},
            "assembly_ms": 1e3 * t_asm / args.steps,
            "assembly_ms_steps": [round(r[0].elapsed_time(r[1]), 3) for r in recs],      # this rank's per-step times (diagnostic)
            "cold": {"mesh_topology_ms": t_mesh, "symbolic_ms": t_sym, "first_assembly_ms": t_first, "bc_setup_ms": t_bc,
                     "nnz_per_s_incl_symbolic": nnz / ((t_sym + t_first) * 1e-3),
                     "roofline_frac_cold": b_survey / ((t_sym + t_first) * 1e-3) / 1e9 / peak},
            "roofline": {"bound": "hbm", "kernel": "spmv_stream_kernel (one CG iteration = spmv+dot, update_xr, update_p)",
                         "achieved": cg_gbs, "peak": peak, "unit": "GB/s", "frac": cg_gbs / peak,
                         "traffic": traffic.get("cg_iteration_bytes"),
                         "peak_source": peak_src, "algorithmic_bytes_per_iter": b_it,
                         "note": "achieved = algorithmic bytes of one CG iteration x iterations / CUDA-event time of cg()"},
            "roofline_assembly": {"bound": "hbm", "kernel": asm_kernel, "achieved": asm_gbs,
                                  "peak": peak, "unit": "GB/s", "frac": asm_gbs / peak,
                                  "traffic": traffic.get("assembly_bytes"),
                                  "algorithmic_bytes": b_warm, "survey_bytes": b_survey,
                                  "frac_survey_bytes": b_survey * args.steps / t_asm / 1e9 / peak,
                                  "note": "algorithmic_bytes = cell + node + coef + values: what the warm numeric phase must move "
                                          "(col/crow/cell2dof are consumed by the cached symbolic phase, not per assembly); "
                                          "frac_survey_bytes keeps SURVEY 8(d)'s cold-assembly byte count for comparison with round 1"},
            "e2e": e2e, "gpu_launches": args.steps * (asm_launches + 3 * args.cg_iters + 7),
            "clocks": clocks, "wall_s_timed_region": t_wall,
        }
        if verify is not None:
            line["verify"] = verify
        if not args.no_cpu_baseline and world == 1:
            try:
                line["cpu_baseline"] = cpu_baseline_leg(args, cfg_id, A_sys if cfg_id == 4 else A, b, M, t_cg / iters)
            except Exception as e:          # the CPU leg must never cost the GPU line
                line["cpu_baseline"] = {"error": f"{type(e).__name__}: {e}"}
            if not args.no_existing_gpu:
                line["existing_gpu_path"] = existing_gpu_subprocess(cfg_id)
        _emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline_leg(args, cfg_id, A, b, M, gpu_s_per_iter):
    """rank 0, N = 1: the reference on the host cores on a bounded sample (~20 s), plus its CG on the SAME matrix as
    the GPU line (copied to the host once)"""
    budget = args.cpu_budget_s or 20.0
    n = args.cpu_n or sample_size(cfg_id, budget, 1)
    m = cpu_measure(cfg_id, n, 1, 0, 5)
    if m is None:
        return {"error": "no CPU baseline for this config without baseline/_ref"}
    out = {"value": m["value"], "unit": UNIT, "cores": m["cores"], "kind": m["kind"], "sample": m["sample"],
           "cg_iters_per_s_sample": m["cg_iters_per_s"],
           "cores_note": "threads OpenBLAS may use inside einsum; the path is otherwise single-threaded numpy"}
    if m["kind"] == "reference":
        import numpy as np
        from fealpy.backend import backend_manager as bm
        bm.set_backend("numpy")
        from fealpy.sparse import CSRTensor
        from fealpy.solver import cg as ref_cg
        t0 = time.perf_counter()
        Ah = CSRTensor(A.crow.cpu().numpy(), A.col.cpu().numpy(), A.values.cpu().numpy(), spshape=A.sparse_shape)
        Mh = None
        if M is not None:
            Mh = CSRTensor(M.crow.cpu().numpy(), M.col.cpu().numpy(), M.values.cpu().numpy(), spshape=M.sparse_shape)
        bh = b.cpu().numpy()
        t1 = time.perf_counter()
        its = 3
        x, info = ref_cg(Ah, bh, M=Mh, atol=0.0, rtol=0.0, maxit=its, returninfo=True)
        t2 = time.perf_counter()
        rate = info["niter"] / (t2 - t1)
        out["cg_same_matrix"] = {"iters_per_s": rate, "iters": int(info["niter"]), "gdof": int(A.shape[0]), "nnz": int(A.nnz),
                                 "d2h_s": t1 - t0, "gpu_over_cpu": (1.0 / gpu_s_per_iter) / rate,
                                 "note": "fealpy.solver.cg (numpy backend -> scipy csr_matvec) on the GPU line's own matrix, host copy"}
    return out


def existing_gpu_leg(cfg_id, dev):
    """the reference's own pytorch backend on this GPU ('the existing GPU path') at a size it fits"""
    import torch
    from baseline import ref_loader
    if not ref_loader.available():
        return {"unavailable": "baseline/_ref not installed"}
    n = {1: 512, 2: 32, 3: 256, 4: 32, 5: 32}[cfg_id]
    try:
        prob = RefProblem(cfg_id, n, backend="pytorch", device=str(dev))
        from fealpy.solver import cg
        A = prob.assemble()                      # warm-up (cuda context of the reference's ops)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        A = prob.assemble()
        torch.cuda.synchronize(dev)
        t1 = time.perf_counter()
        out = {"kind": "reference pytorch backend, device=cuda", "n": n, "nnz": int(A.nnz), "value": A.nnz / (t1 - t0), "unit": UNIT,
               "assembly_s": t1 - t0}
        # a second, larger size (BASELINE.md section 4 asks for 32^3 / 64^3): assembly only
        n2 = {2: 64, 4: 64}.get(cfg_id)
        if n2:
            try:
                prob2 = RefProblem(cfg_id, n2, backend="pytorch", device=str(dev))
                A2_ = prob2.assemble()
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                A2_ = prob2.assemble()
                torch.cuda.synchronize(dev)
                dt = time.perf_counter() - t0
                out["larger"] = {"n": n2, "nnz": int(A2_.nnz), "value": A2_.nnz / dt, "unit": UNIT, "assembly_s": dt}
                del prob2, A2_
                torch.cuda.empty_cache()
            except Exception as e:
                out["larger"] = {"n": n2, "error": f"{type(e).__name__}: {str(e)[:200]}"}
        _emit(json.dumps(dict(out, cg_error="the reference's torch-cuda cg did not return (CUDA fault)")))     # survives a fault below
        try:
            A2, b, M = prob.rhs_and_system(A)
            torch.cuda.synchronize(dev)
            t2 = time.perf_counter()
            x, info = cg(A2, b, M=M, atol=0.0, rtol=0.0, maxit=20, returninfo=True)
            torch.cuda.synchronize(dev)
            t3 = time.perf_counter()
            out["cg_iters_per_s"] = info["niter"] / (t3 - t2)
        except Exception as e:
            out["cg_error"] = f"{type(e).__name__}: {str(e)[:200]}"
        return out
    finally:
        from fealpy.backend import backend_manager as bm
        bm.set_backend("numpy")


def existing_gpu_subprocess(cfg_id):
    """the reference's own torch-cuda code runs in a CHILD process: a CUDA fault inside it (its cg hit an illegal memory
    access on this stack) must not poison this process's context"""
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--existing-gpu-child", "--config", str(cfg_id)],
                           capture_output=True, text=True, timeout=240)
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"error": f"child rc={r.returncode}: {r.stderr.strip()[-300:]}"}
    except Exception as e:
        return {"error": f"{type(e).__name__}: {str(e)[:300]}"}


class StdoutToStderr:
    """Everything libraries print on fd 1 (e.g. NCCL's version banner) goes to stderr; only the final JSON
    line is written to the real stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, line):
        sys.stdout.flush()
        os.write(self.saved, (line + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


if __name__ == "__main__":
    a = parse()
    with StdoutToStderr() as out:
        _emit = out.emit
        if a.existing_gpu_child:
            import torch
            try:
                res = existing_gpu_leg(a.config, torch.device("cuda", 0))
            except Exception as e:
                res = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
            _emit(json.dumps(res))
            os._exit(0)                       # skip torch's teardown: the context may be in an error state
        elif a.impl == "reference":
            run_reference(a)
        else:
            run_ours(a)
