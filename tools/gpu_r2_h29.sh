#!/bin/bash
# round 2, call h29: config 1 (L2-resident matrix) CG with and without the evict_last gather policy
mkdir -p gpurun_out
V=fealpy_b200/csrc/build/variants
{
for k in 1 2 3; do
FB2_SPMV_KERNEL=el python tools/gpu_time_cg.py 1
FB2_SPMV_KERNEL=default-policy FB2_LIB_PATH=$V/noel.so python tools/gpu_time_cg.py 1
done
} 2>&1 | grep -v Warning | tee gpurun_out/h29_tune_cg_cfg1.txt
