#!/bin/bash
# round 2, call h17: smoke(), bench lines of configs 1, 3, 4 on the final tree (traffic fields from the corrected r02_traffic.json)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
for c in 1 3 4; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/h17_bench_cfg$c.json 2> gpurun_out/h17_bench_cfg$c.err; echo "bench cfg$c rc=$?"
done
python - <<'PY'
import json
for c in (1,3,4):
    for l in open(f"gpurun_out/h17_bench_cfg{c}.json"):
        if l.startswith("{"):
            d=json.loads(l); print(c, f"{d['value']:.3e}", round(d["assembly_ms"],3), round(d["cg"]["iters_per_s"],1), round(d["roofline"]["frac"],3), d["roofline"]["traffic"], round(d["roofline_assembly"]["frac"],3), d["roofline_assembly"]["traffic"], {k:round(v,1) for k,v in d["cold"].items() if k.endswith("_ms")}, f"{d['e2e']['value']:.3e}", d["clocks"])
PY
