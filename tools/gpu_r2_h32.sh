#!/bin/bash
# round 2, call h32: bench default + config 1 with the 30 ms clock sampler
mkdir -p gpurun_out
for c in 2 1; do
timeout 900 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/h32_bench_cfg$c.json 2> gpurun_out/h32_bench_cfg$c.err; echo "bench cfg$c rc=$?"
python - <<PY
import json
for l in open("gpurun_out/h32_bench_cfg$c.json"):
    if l.startswith("{"):
        d=json.loads(l); print($c, f"{d['value']:.4e}", d["assembly_ms"], d["cg"]["iters_per_s"], d["roofline"]["frac"], d["clocks"], d["cold"]["symbolic_ms"], d.get("existing_gpu_path",{}).get("larger"))
PY
done
