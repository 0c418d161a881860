#!/bin/bash
# round 2, call h23: last check of the final tree: GPU suite, smoke, the default bench line
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/h23_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h23_pytest.log
tail -3 gpurun_out/h23_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/h23_bench_cfg2.json 2> gpurun_out/h23_bench_cfg2.err; echo "bench rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/h23_bench_cfg2.json"):
    if l.startswith("{"):
        d=json.loads(l); print(f"{d['value']:.4e}", d["assembly_ms"], d["cg"]["iters_per_s"], d["roofline"]["frac"], d["roofline"]["traffic"], d["roofline_assembly"]["frac"], d["roofline_assembly"]["traffic"], d["cold"], f"{d['e2e']['value']:.3e}", d["clocks"])
PY
