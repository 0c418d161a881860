#!/bin/bash
# round 2, call h28: streaming SpMV: unroll 5 / 2 (tile 2560 = 256 x 10), 320 threads per CTA (2560 = 320 x 8), against the default
mkdir -p gpurun_out
V=fealpy_b200/csrc/build/variants
{
FB2_SPMV_KERNEL=base python tools/gpu_time_cg.py 2
for v in stu5 stu2 stt320 stt320u2; do FB2_SPMV_KERNEL=$v FB2_LIB_PATH=$V/$v.so python tools/gpu_time_cg.py 2; done
FB2_SPMV_KERNEL=base python tools/gpu_time_cg.py 4
for v in stu5 stt320; do FB2_SPMV_KERNEL=$v FB2_LIB_PATH=$V/$v.so python tools/gpu_time_cg.py 4; done
} 2>&1 | grep -v Warning | tee gpurun_out/h28_tune_cg.txt
