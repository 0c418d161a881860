#!/bin/bash
# round 2, call h4: what bounds the SpMV -- column patterns of decreasing gather cost on config 2's structure
mkdir -p gpurun_out
python tools/gpu_spmv_bound.py 128 2>&1 | grep -v Warning | tee gpurun_out/h4_spmv_bound.txt
