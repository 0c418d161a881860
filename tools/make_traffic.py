"""profiles/r02_traffic.json from the ncu CSVs of tools/gpu_traffic_probe.py (metrics dram__bytes_read.sum,
dram__bytes_write.sum, gpu__time_duration.sum per launch): per config the DRAM bytes of ONE warm assembly (its last
occurrence in the capture) and of ONE CG iteration (SpMV+dot, x/r update, p update)."""
import csv
import json
import sys

ASM = ("cell_geometry4_kernel", "assemble_const_v6_kernel", "assemble_const_kernel", "elem_quad_kernel", "elem_const_kernel",
       "assemble_from_ke_kernel", "cell_gradients_kernel", "bc_to_points_kernel")
CG = ("spmv_stream_kernel", "cg_update_xr_kernel", "cg_update_p_kernel")


def parse(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            h, st = r, i
            break
    ki, mi, vi, ui, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("ID")
    launches = {}
    for r in rows[st + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1)      # us
        d = launches.setdefault(int(r[ii]), {"name": r[ki]})
        d[r[mi]] = v * mult
    return [launches[k] for k in sorted(launches)]


def base(name):
    return name.split("<")[0].split("(")[0].replace("void ", "").replace("fb2::", "").strip()


out = {}
for arg in sys.argv[1:]:
    cfg, path = arg.split("=")
    seq = parse(path)
    # a representative launch per kernel: the median (by duration) of its full-length launches -- the CG graph replays
    # a chunk of iterations and the launches after the stopping iteration exit at once, those must not be picked
    groups = {}
    for d in seq:
        groups.setdefault(base(d["name"]), []).append(d)
    last = {}
    for k, lst in groups.items():
        tmax = max(x["gpu__time_duration.sum"] for x in lst)
        full = sorted((x for x in lst if x["gpu__time_duration.sum"] >= 0.5 * tmax), key=lambda x: x["gpu__time_duration.sum"])
        last[k] = full[len(full) // 2]
    ent = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none "
                     "python tools/gpu_traffic_probe.py %s (per launch; median full-length launch of each kernel)" % cfg,
           "assembly_kernels": {}, "cg_kernels": {}}
    for k in ASM:
        if k in last:
            ent["assembly_kernels"][k] = {"dram_bytes": int(last[k]["dram__bytes_read.sum"] + last[k]["dram__bytes_write.sum"]),
                                         "us": round(last[k]["gpu__time_duration.sum"], 1)}
    for k in CG:
        if k in last:
            ent["cg_kernels"][k] = {"dram_bytes": int(last[k]["dram__bytes_read.sum"] + last[k]["dram__bytes_write.sum"]),
                                   "us": round(last[k]["gpu__time_duration.sum"], 1)}
    ent["assembly_bytes"] = sum(v["dram_bytes"] for v in ent["assembly_kernels"].values())
    ent["cg_iteration_bytes"] = sum(v["dram_bytes"] for v in ent["cg_kernels"].values())
    out[cfg] = ent
json.dump(out, open("profiles/r02_traffic.json", "w"), indent=1)
print(json.dumps({k: (v["assembly_bytes"], v["cg_iteration_bytes"]) for k, v in out.items()}))
