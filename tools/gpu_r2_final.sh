#!/bin/bash
# round 2, final evidence on the final tree: GPU suite, per-launch DRAM traffic of every config (-> profiles/r02_traffic.json),
# launch list of the bench command, bench lines of configs 1-4, the reference arm, sanitizer passes
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -rs > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
tail -6 gpurun_out/f_pytest.log
for c in 2 1 3 4; do
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/f_traffic_cfg$c.csv python tools/gpu_traffic_probe.py $c > gpurun_out/f_traffic_cfg$c.log 2>&1
done
python tools/make_traffic.py 2=gpurun_out/f_traffic_cfg2.csv 1=gpurun_out/f_traffic_cfg1.csv 3=gpurun_out/f_traffic_cfg3.csv 4=gpurun_out/f_traffic_cfg4.csv
cp profiles/r02_traffic.json gpurun_out/f_r02_traffic.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-existing-gpu > gpurun_out/f_bench_under_ncu.log 2>&1
for c in 2 1 3 4; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/f_bench_cfg$c.json 2> gpurun_out/f_bench_cfg$c.err; echo "bench cfg$c rc=$?"
done
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 --cpu-budget-s 60 > gpurun_out/f_bench_reference.json 2> gpurun_out/f_bench_reference.err; echo "bench reference rc=$?"
python - <<'PY'
import json
for c in (2,1,3,4):
    for l in open(f"gpurun_out/f_bench_cfg{c}.json"):
        if l.startswith("{"):
            d=json.loads(l); print(c, f"{d['value']:.3e}", round(d["assembly_ms"],3), round(d["cg"]["iters_per_s"],1), round(d["roofline"]["frac"],3), d["roofline"]["traffic"], round(d["roofline_assembly"]["frac"],3), d["roofline_assembly"]["traffic"], d["cold"]["symbolic_ms"], f"{d['e2e']['value']:.3e}", d["clocks"])
for l in open("gpurun_out/f_bench_reference.json"):
    if l.startswith("{"): print(l[:600])
PY
SEL='assembly_csr or spmv_and_cg or poisson_source or batched or slab or sort_and_scan or reference or schedule or dirichlet or elasticity or matfree or summed'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "($SEL) and not config2_size" > gpurun_out/f_memcheck.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/f_memcheck.txt
tail -4 gpurun_out/f_memcheck.txt
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "($SEL) and not config2_size" > gpurun_out/f_racecheck.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/f_racecheck.txt
tail -4 gpurun_out/f_racecheck.txt
