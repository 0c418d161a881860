#!/bin/bash
# round 2, call g3: assembly v6 occupancy variants (JC=10), compressed-column SpMV against the 32-bit stream, parity
mkdir -p gpurun_out
V=fealpy_b200/csrc/build/variants
{
python tools/gpu_time_asm.py 2
FB2_LIB_PATH=$V/w3.so python tools/gpu_time_asm.py 2
FB2_LIB_PATH=$V/w2r4.so python tools/gpu_time_asm.py 2
FB2_ASM4_TILE=2432 FB2_LIB_PATH=$V/w2r3.so python tools/gpu_time_asm.py 2
FB2_ASM4_TILE=2304 FB2_LIB_PATH=$V/w2r3.so python tools/gpu_time_asm.py 2
} 2>&1 | grep -v Warning | tee gpurun_out/g3_tune_asm.txt
{
python tools/gpu_time_cg.py 2
FB2_SPMV_COLZ=0 python tools/gpu_time_cg.py 2
python tools/gpu_time_cg.py 1
FB2_SPMV_COLZ=0 python tools/gpu_time_cg.py 1
python tools/gpu_time_cg.py 3
FB2_SPMV_COLZ=0 python tools/gpu_time_cg.py 3
python tools/gpu_time_cg.py 4
FB2_SPMV_COLZ=0 python tools/gpu_time_cg.py 4
} 2>&1 | grep -v Warning | tee gpurun_out/g3_tune_cg.txt
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/g3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g3_pytest.log
tail -8 gpurun_out/g3_pytest.log
