#!/bin/bash
# round 2, call g8: SpMV tile->CTA map at full occupancy
mkdir -p gpurun_out
{
python tools/gpu_time_cg.py 2
FB2_SPMV_MAP=strided python tools/gpu_time_cg.py 2
python tools/gpu_time_cg.py 3
FB2_SPMV_MAP=strided python tools/gpu_time_cg.py 3
python tools/gpu_time_cg.py 4
FB2_SPMV_MAP=strided python tools/gpu_time_cg.py 4
python tools/gpu_time_cg.py 1
FB2_SPMV_MAP=strided python tools/gpu_time_cg.py 1
} 2>&1 | grep -v Warning | tee gpurun_out/g8_tune_cg.txt
