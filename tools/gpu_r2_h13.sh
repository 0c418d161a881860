#!/bin/bash
# round 2, call h13: the default bench line twice on a fresh box (the final-evidence run measured its CG after ten minutes of profiling: 550 it/s)
mkdir -p gpurun_out
for k in a b; do
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/h13_bench_cfg2_$k.json 2> gpurun_out/h13_bench_cfg2_$k.err; echo "bench rc=$?"
python - <<PY
import json
for l in open("gpurun_out/h13_bench_cfg2_$k.json"):
    if l.startswith("{"):
        d=json.loads(l); print(f"{d['value']:.4e}", d["assembly_ms"], d["cg"]["iters_per_s"], d["roofline"]["frac"], d["clocks"], d["cold"]["symbolic_ms"])
PY
done
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active --format=csv
