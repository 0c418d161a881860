#!/bin/bash
# round 2, call h27: v6 with the geometry records fetched two batches ahead (three register buffers, loop unrolled by three)
mkdir -p gpurun_out
V=fealpy_b200/csrc/build/variants
{
python tools/gpu_time_asm.py 2; FB2_LIB_PATH=$V/a6pf2.so python tools/gpu_time_asm.py 2
python tools/gpu_time_asm.py 1; FB2_LIB_PATH=$V/a6pf2.so python tools/gpu_time_asm.py 1
FB2_LIB_PATH=$V/a6pf2.so python tools/gpu_time_asm.py 2 64
} 2>&1 | grep -v Warning | tee gpurun_out/h27_tune_asm.txt
FB2_LIB_PATH=$V/a6pf2.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or assembly or paths or schedule or full_values" > gpurun_out/h27_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h27_pytest.log
tail -3 gpurun_out/h27_pytest.log
