#!/bin/bash
# round 2, call h26: K_e / elasticity gather with streaming stores of the values; v6 default with the L2 policies (check)
mkdir -p gpurun_out
V=fealpy_b200/csrc/build/variants
{
python tools/gpu_time_asm.py 4; FB2_LIB_PATH=$V/gcs.so python tools/gpu_time_asm.py 4
python tools/gpu_time_asm.py 3; FB2_LIB_PATH=$V/gcs.so python tools/gpu_time_asm.py 3
python tools/gpu_time_asm.py 2
} 2>&1 | grep -v Warning | tee gpurun_out/h26_tune_gather_cs.txt
