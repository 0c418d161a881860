#!/bin/bash
# round 2, call h9: full GPU suite after the batched-CG / vector-source / gather-tile changes; config 3 timing
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/h9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h9_pytest.log
tail -4 gpurun_out/h9_pytest.log
python tools/gpu_time_asm.py 3 2>&1 | grep -v Warning | tee gpurun_out/h9_time_cfg3.txt
