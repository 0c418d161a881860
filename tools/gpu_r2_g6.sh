#!/bin/bash
# round 2, call g6: bulk-copy SpMV with more gathers in flight (512 / 1024 threads per CTA)
mkdir -p gpurun_out
V=fealpy_b200/csrc/build/variants
{
FB2_LIB_PATH=$V/t512.so python tools/gpu_time_cg.py 2
FB2_LIB_PATH=$V/t512.so FB2_SPMV_TILE=2048 FB2_SPMV_PERSM=4 python tools/gpu_time_cg.py 2
FB2_LIB_PATH=$V/t512u3.so FB2_SPMV_TILE=2048 FB2_SPMV_PERSM=4 python tools/gpu_time_cg.py 2
FB2_LIB_PATH=$V/t512u3.so FB2_SPMV_TILE=1536 FB2_SPMV_PERSM=4 python tools/gpu_time_cg.py 2
FB2_LIB_PATH=$V/t1024.so FB2_SPMV_TILE=3072 FB2_SPMV_PERSM=2 python tools/gpu_time_cg.py 2
FB2_LIB_PATH=$V/t1024.so FB2_SPMV_TILE=4096 FB2_SPMV_PERSM=2 python tools/gpu_time_cg.py 2
FB2_SPMV_KERNEL=stream python tools/gpu_time_cg.py 2
} 2>&1 | grep -v Warning | tee gpurun_out/g6_tune_cg.txt
