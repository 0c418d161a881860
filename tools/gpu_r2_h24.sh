#!/bin/bash
# round 2, call h24: SpMV with L2 cache policies: (val, col) stream evict_first and / or x gather evict_last, against the default
mkdir -p gpurun_out
V=fealpy_b200/csrc/build/variants
{
for c in 2 3 4; do
  FB2_SPMV_KERNEL=base python tools/gpu_time_cg.py $c
  for v in ef efel el; do FB2_SPMV_KERNEL=$v FB2_LIB_PATH=$V/$v.so python tools/gpu_time_cg.py $c; done
done
} 2>&1 | grep -v Warning | tee gpurun_out/h24_tune_cg.txt
