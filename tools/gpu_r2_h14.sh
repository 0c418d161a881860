#!/bin/bash
# round 2, call h14: set-up (cold) phases of configs 4 and 3
mkdir -p gpurun_out
python tools/gpu_cold_cfg34.py 2>&1 | grep -v Warning | tee gpurun_out/h14_cold_cfg34.txt
