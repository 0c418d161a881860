#!/usr/bin/env python
"""Benchmark of the hot path: P2 tetrahedral Poisson (diffusion + mass) global assembly to CSR
plus CG, on synthetic `TetrahedronMesh.from_box` meshes (BASELINE.json configs[1]).

    python bench.py --gpus 1 --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --steps K --warmup W    # the CPU arm (oracle port of the reference)

One "step" = one assembly() of the form (pattern cached per space = "warm") followed by a
fixed number of CG iterations on A x = A 1.  `value` = assembled nnz/s with all inputs resident
in HBM; `cg.iters_per_s` is reported beside it; `e2e` repeats the step through the public API
with HOST (pinned) inputs: H2D of node/cell/cell2dof and of b, D2H of the CSR values and of x,
all inside the timed region.  Prints ONE JSON line on rank 0.
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=128, help="cells per box edge (per GPU)")
    ap.add_argument("--p", type=int, default=2)
    ap.add_argument("--cg-iters", type=int, default=100)
    ap.add_argument("--cpu-n", type=int, default=20, help="box edge of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def measured_traffic():
    """per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of the two dominant kernels from the
    committed `ncu --set full` capture of this same workload (profiles/r01_traffic.json), or {}"""
    path = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return {}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed region.  In-process NVML (a 100 ms thread) when
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
            self.stop_flag.wait(0.1)

    def start(self):
        try:
            import threading
            import pynvml as nv
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
# CPU arm: the oracle port of the reference path on the host cores
# --------------------------------------------------------------------------------------
def cpu_problem(n, p):
    from oracle import fem_oracle as O
    node, cell = O.tet_from_box([0, 1, 0, 1, 0, 1], n, n, n)
    mesh = O.Mesh(node, cell)
    c2d = mesh.cell_to_ipoint(p)
    return O, mesh, c2d, mesh.number_of_global_ipoints(p)


def cpu_step(O, mesh, c2d, gdof, p, cg_iters):
    t0 = time.perf_counter()
    groups = [(O.diffusion_element(mesh, p), c2d), (O.mass_element(mesh, p), c2d)]
    crow, col, val = O.assemble(groups, gdof)
    t1 = time.perf_counter()
    b = O.csr_matvec(crow, col, val, __import__("numpy").ones(gdof))
    t2 = time.perf_counter()
    x, info = O.cg(lambda v: O.csr_matvec(crow, col, val, v), b, atol=0.0, rtol=0.0, maxit=cg_iters)
    t3 = time.perf_counter()
    return len(val), t1 - t0, info["niter"], t3 - t2


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    O, mesh, c2d, gdof = cpu_problem(args.cpu_n, args.p)
    for _ in range(min(args.warmup, 1)):
        cpu_step(O, mesh, c2d, gdof, args.p, 5)
    t_asm = t_cg = 0.0
    nnz = its = 0
    for _ in range(args.steps):
        nz, ta, it, tc = cpu_step(O, mesh, c2d, gdof, args.p, args.cg_iters)
        nnz += nz; t_asm += ta; its += it; t_cg += tc
    val = nnz / t_asm
    sample = f"tet P{args.p} from_box n={args.cpu_n} ({mesh.NC} cells, gdof {gdof}), diffusion+mass q={args.p + 3}, numpy oracle port"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * (t_asm + t_cg) / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"tet P{args.p} Poisson diffusion+mass assembly + CG, from_box n=128 per GPU (CPU arm: bounded sample)"},
        "cg": {"iters_per_s": its / t_cg, "iters_per_step": args.cg_iters},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                         "cg_iters_per_s": its / t_cg},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(json.dumps(line))


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
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

    from fealpy_b200.mesh import TetrahedronMesh
    from fealpy_b200.functionspace import LagrangeFESpace
    from fealpy_b200.fem import BilinearForm, ScalarDiffusionIntegrator, ScalarMassIntegrator
    from fealpy_b200.fem.bilinear_form import symbolic_pattern
    from fealpy_b200.solver import cg

    n, p = args.n, args.p
    stream = torch.cuda.Stream(device=dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    with torch.cuda.stream(stream):
        # ---- setup (not timed as part of the step; reported as cold costs) --------------------
        e0, e1, e2, e3 = ev(), ev(), ev(), ev()
        e0.record()
        part = None
        if world == 1:
            mesh = TetrahedronMesh.from_box([0, 1, 0, 1, 0, 1], n, n, n, device=dev)
            space = LagrangeFESpace(mesh, p)
        else:
            # weak scaling: ONE global box of (n*world) x n x n cubes, row-partitioned into x-slabs;
            # every rank assembles its owned rows (+1 ghost cell layer) and CG swaps halo planes over NCCL
            from fealpy_b200.parallel import SlabProblem, CudaCgOps, dist_cg
            sp = SlabProblem([0, world, 0, 1, 0, 1], n * world, n, n, p, world, rank, device=dev)
            mesh, space, part = sp.mesh, sp.space, sp.part
        c2d = space.cell_to_dof()
        e1.record()
        sym = symbolic_pattern(space)
        e2.record()
        bform = BilinearForm(space)
        bform.add_integrator(ScalarDiffusionIntegrator())
        bform.add_integrator(ScalarMassIntegrator())
        A = bform.assembly()
        e3.record()
        torch.cuda.synchronize(dev)
        t_mesh, t_sym, t_first = e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3)
        nnz, gdof, NC, NN = A.nnz, A.shape[0], mesh.number_of_cells(), mesh.number_of_nodes()
        L = c2d.shape[1]
        ones = torch.ones(gdof, dtype=torch.float64, device=dev)
        b = A @ ones
        if part is not None:       # count only owned rows (halo rows are scratch)
            cr = A.crow
            (a0, a1), (b0, b1) = part.own_nodes, part.own_edges
            nnz = int(cr[a1] - cr[a0]) + int(cr[b1] - cr[b0])
            own_mask = torch.zeros(gdof, dtype=torch.bool, device=dev)
            own_mask[a0:a1] = True
            own_mask[b0:b1] = True

        def solve(A_, rhs):
            if part is None:
                return cg(A_, rhs, atol=0.0, rtol=0.0, maxit=args.cg_iters, returninfo=True)
            return dist_cg(CudaCgOps(A_, part.own_ranges), rhs, torch.zeros_like(rhs), part.exchanges, atol=0.0, rtol=0.0,
                           maxit=args.cg_iters, check_every=args.cg_iters)

        def step():
            s0, s1, s2 = ev(), ev(), ev()
            s0.record()
            A_ = bform.assembly()
            s1.record()
            x, info = solve(A_, b)
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
        # fires between s0.record() and the kernel launch it shows up as "assembly time" (seen: one step of 15 ms
        # among four of 4.3 ms).  Collect now, keep the collector off inside the timed regions.
        import gc
        gc.collect()
        gc.disable()
        t_wall0 = time.perf_counter()
        # only the LAST solution is kept: holding every step's x (136 MB each) makes torch's caching allocator carve
        # them out of the freed 3.9 GB `values` blocks, and every fifth assembly then pays a 150 ms cudaMalloc
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
        xerr = float(((last_x - 1.0)[own_mask] if part is not None else (last_x - 1.0)).abs().max())

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
                A_ = bform.assembly()
                h_vals.copy_(A_.values, non_blocking=True)
                s1.record()
                b.copy_(h_b, non_blocking=True)
                x, info = solve(A_, b)
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

    # ---- reduce over ranks (max time) ----------------------------------------------------------------
    times = torch.tensor([t_asm, t_cg, t_wall], dtype=torch.float64, device=dev)
    nnz_local = nnz
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        tot = torch.tensor([nnz, NC, part.n_owned], dtype=torch.int64, device=dev)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        nnz_glob, NC_glob, gdof_glob = (int(v) for v in tot)
        if e2e is not None:
            ev_ = torch.tensor([nnz * 1.0 / e2e["value"], 1.0 / e2e["cg_iters_per_s"]], dtype=torch.float64, device=dev)
            dist.all_reduce(ev_, op=dist.ReduceOp.MAX)
            e2e["value"] = nnz_glob / float(ev_[0])
            e2e["cg_iters_per_s"] = 1.0 / float(ev_[1])
    t_asm, t_cg, t_wall = (float(v) for v in times)

    if world == 1:
        nnz_glob, NC_glob, gdof_glob = nnz, NC, gdof
    if rank == 0:
        peak, peak_src = peaks()
        traffic = measured_traffic()
        # algorithmic bytes, per launch == per step for both kernels
        b_asm = 4 * NC * 4 + 4 * NC * L + 8 * 3 * NN + 12 * nnz + 8 * (gdof + 1)
        b_it = 12 * nnz + 8 * (gdof + 1) + 104 * gdof
        asm_gbs = b_asm * args.steps / t_asm / 1e9          # per GPU (rank 0's share of the bytes, max-over-ranks time)
        cg_gbs = b_it * iters / t_cg / 1e9
        line = {
            "metric": METRIC, "value": nnz_glob * args.steps / t_asm, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * (t_asm + t_cg) / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"tet P{p} Poisson: ScalarDiffusionIntegrator + ScalarMassIntegrator (q={p + 3}) assembly to CSR "
                                   f"+ {args.cg_iters} CG iterations, TetrahedronMesh.from_box n={n} per GPU",
                       "NC": NC_glob, "gdof": gdof_glob, "nnz": nnz_glob, "l2_policy": "inputs larger than L2 (multi-GB arrays per step)",
                       "assembly_path": bform.last_path, "pattern": "warm (symbolic cached per space)",
                       "parallelism": (f"{world} x-slabs of one {n * world}x{n}x{n} box: owned-row assembly + NCCL halo CG"
                                       if world > 1 else "single GPU")},
            "cg": {"iters_per_s": iters / t_cg, "iters_per_step": args.cg_iters, "ms_per_iter": 1e3 * t_cg / iters,
                   "x_err_vs_exact": xerr},
            "assembly_ms": 1e3 * t_asm / args.steps,
            "assembly_ms_steps": [round(r[0].elapsed_time(r[1]), 3) for r in recs],      # this rank's per-step times (diagnostic)
            "cold": {"mesh_topology_ms": t_mesh, "symbolic_ms": t_sym, "first_assembly_ms": t_first,
                     "nnz_per_s_incl_symbolic": nnz / ((t_sym + t_first) * 1e-3)},
            "roofline": {"bound": "hbm", "kernel": "spmv_stream_kernel<1> (one CG iteration = spmv+dot, update_xr, update_p)",
                         "achieved": cg_gbs, "peak": peak, "unit": "GB/s", "frac": cg_gbs / peak,
                         "traffic": traffic.get("cg_iteration_bytes") if (n == 128 and p == 2) else None,
                         "peak_source": peak_src, "algorithmic_bytes_per_iter": b_it,
                         "note": "achieved = algorithmic bytes of one CG iteration x iterations / CUDA-event time of cg()"},
            "roofline_assembly": {"bound": "hbm", "kernel": "cell_geometry4_kernel + assemble_const_v4_kernel", "achieved": asm_gbs,
                                  "peak": peak, "unit": "GB/s", "frac": asm_gbs / peak,
                                  "traffic": traffic.get("assembly_bytes") if (n == 128 and p == 2) else None,
                                  "algorithmic_bytes": b_asm,
                                  "note": "warm pattern: col/crow are cached and not re-written; limiter is the L1/LSU data pipe (DESIGN.md)"},
            "e2e": e2e, "gpu_launches": args.steps * (2 + 3 * args.cg_iters + 7),
            "clocks": clocks, "wall_s_timed_region": t_wall,
        }
        if not args.no_cpu_baseline and world == 1:
            O, om, oc2d, ogdof = cpu_problem(args.cpu_n, p)
            nz, ta, it, tc = cpu_step(O, om, oc2d, ogdof, p, min(args.cg_iters, 50))
            line["cpu_baseline"] = {"value": nz / ta, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": f"tet P{p} from_box n={args.cpu_n} ({om.NC} cells, nnz {nz}), numpy oracle port of the "
                                              f"reference path, {ta:.1f}s assembly",
                                    "cg_iters_per_s": it / tc}
        _emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


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
        if a.impl == "reference":
            run_reference(a)
        else:
            run_ours(a)
