#!/bin/bash
# round 2, call h25: assembly v6 with L2 cache policies: geometry records evict_last (hk), + values st.global.cs (hkcs),
# + schedule blocks evict_first (all), st.cs alone (cs), against the default
mkdir -p gpurun_out
V=fealpy_b200/csrc/build/variants
{
python tools/gpu_time_asm.py 2
for v in a6hk a6hkcs a6all a6cs; do echo "variant $v"; FB2_LIB_PATH=$V/$v.so python tools/gpu_time_asm.py 2; done
python tools/gpu_time_asm.py 1
FB2_LIB_PATH=$V/a6all.so python tools/gpu_time_asm.py 1
} 2>&1 | grep -v Warning | tee gpurun_out/h25_tune_asm.txt
