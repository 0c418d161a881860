#!/bin/bash
# round 2, call h22: flat replay of the symbolic stash (thread per adjacency position) against the warp-per-row replay
mkdir -p gpurun_out
{ python tools/gpu_cold_breakdown.py 2 | grep "pass 1\|symbolic"; FB2_SYM_REPLAY=r python tools/gpu_cold_breakdown.py 2 | grep "pass 1\|symbolic"; } 2>&1 | grep -v Warning | tee gpurun_out/h22_cold.txt
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/h22_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h22_pytest.log
tail -3 gpurun_out/h22_pytest.log
