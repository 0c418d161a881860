#!/bin/bash
# round 2, call h8: batched CG with the device-side stopping test + 8-column SpMM; gather tile sweep on config 3
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/h8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h8_pytest.log
tail -4 gpurun_out/h8_pytest.log
{
for t in 4096 3072 2048 1536; do FB2_ASM_TILE=$t python tools/gpu_time_asm.py 3; done
for t in 4096 2048; do FB2_ASM_TILE=$t python tools/gpu_time_asm.py 4; done
} 2>&1 | grep -v Warning | tee gpurun_out/h8_tune_gather_tile.txt
