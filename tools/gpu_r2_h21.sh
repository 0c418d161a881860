#!/bin/bash
# round 2, call h21: symbolic phase with rank-by-counting instead of the bitonic network: cold breakdown + full GPU suite
mkdir -p gpurun_out
python tools/gpu_cold_breakdown.py 2 2>&1 | grep -v Warning | tee gpurun_out/h21_cold.txt
for c in 1 3 4; do python tools/gpu_cold_breakdown.py $c 2>&1 | grep -v Warning | grep "pass 1\|symbolic" | tail -2; done | tee -a gpurun_out/h21_cold.txt
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/h21_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h21_pytest.log
tail -4 gpurun_out/h21_pytest.log
