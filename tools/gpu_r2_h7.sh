#!/bin/bash
# round 2, call h7: K_e gather chunk size (independent element-row loads in flight per lane group) against the valence
mkdir -p gpurun_out
{
for u in 2 4 8; do FB2_GATHER_U=$u python tools/gpu_time_asm.py 3; done
python tools/gpu_time_asm.py 3
for u in 2 4 8; do FB2_GATHER_U=$u FB2_ASM_PATH=gather python tools/gpu_time_asm.py 2 64; done
} 2>&1 | grep -v Warning | tee gpurun_out/h7_tune_gather.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/h7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h7_pytest.log
tail -3 gpurun_out/h7_pytest.log
