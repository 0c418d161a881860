#!/bin/bash
# round 2, call g1: bulk-copy assembly kernel (v6) against the cp.async-ring kernel, compile-time variants, parity, ncu
mkdir -p gpurun_out
V=fealpy_b200/csrc/build/variants
{
python tools/gpu_time_asm.py 2
FB2_ASM4_VARIANT=ring python tools/gpu_time_asm.py 2
FB2_LIB_PATH=$V/w3.so python tools/gpu_time_asm.py 2
FB2_LIB_PATH=$V/jc10.so python tools/gpu_time_asm.py 2
FB2_LIB_PATH=$V/jc2.so python tools/gpu_time_asm.py 2
FB2_ASM4_TILE=2048 python tools/gpu_time_asm.py 2
FB2_ASM4_TILE=3072 python tools/gpu_time_asm.py 2
python tools/gpu_time_asm.py 1
FB2_ASM4_VARIANT=ring python tools/gpu_time_asm.py 1
} 2>&1 | grep -v Warning | tee gpurun_out/g1_tune_asm.txt
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/g1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g1_pytest.log
tail -8 gpurun_out/g1_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:assemble_const_v6 -s 2 -c 1 -o gpurun_out/g1_ncu_v6 python tools/gpu_time_asm.py 2 > gpurun_out/g1_ncu_v6.log 2>&1
ncu -i gpurun_out/g1_ncu_v6.ncu-rep --page raw --csv > gpurun_out/g1_ncu_v6_raw.csv 2>/dev/null
ncu -i gpurun_out/g1_ncu_v6.ncu-rep --page source --csv > gpurun_out/g1_ncu_v6_source.csv 2>/dev/null
rm -f gpurun_out/g1_ncu_v6.ncu-rep
