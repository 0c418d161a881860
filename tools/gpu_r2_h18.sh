#!/bin/bash
# round 2, call h18: cost of the Morton partition at size (64^3 and 128^3 tet P2, 8 ranks)
mkdir -p gpurun_out
{ python tools/gpu_time_partition.py 64 8; python tools/gpu_time_partition.py 128 8; } 2>&1 | grep -v Warning | tee gpurun_out/h18_partition.txt
