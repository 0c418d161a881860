#!/bin/bash
# round 2, first GPU call: parity tests, default bench line (with the reference CPU leg + the reference's own torch-cuda path), other configs
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2_gpu.txt
free -g > gpurun_out/r2_host.txt; nproc >> gpurun_out/r2_host.txt
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
tail -15 gpurun_out/r2_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_cfg2.json 2> gpurun_out/r2_bench_cfg2.err; echo "bench2 rc=$?"
for c in 1 3 4; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2_bench_cfg$c.json 2> gpurun_out/r2_bench_cfg$c.err; echo "bench$c rc=$?"
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 --cpu-budget-s 40 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"
tail -c 1500 gpurun_out/r2_bench_cfg2.json
tail -5 gpurun_out/r2_bench_cfg2.err
