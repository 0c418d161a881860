#!/bin/bash
# round 2, call g4: staged-x SpMV against the direct gather; full parity suite; cold-path breakdown; v6 (3 warps) timing
mkdir -p gpurun_out
{
python tools/gpu_time_cg.py 2
FB2_SPMV_COLZ=0 python tools/gpu_time_cg.py 2
python tools/gpu_time_cg.py 1
python tools/gpu_time_cg.py 3
FB2_SPMV_COLZ=0 python tools/gpu_time_cg.py 3
python tools/gpu_time_cg.py 4
} 2>&1 | grep -v Warning | tee gpurun_out/g4_tune_cg.txt
{
python tools/gpu_time_asm.py 2
python tools/gpu_time_asm.py 1
} 2>&1 | grep -v Warning | tee gpurun_out/g4_tune_asm.txt
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/g4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g4_pytest.log
tail -8 gpurun_out/g4_pytest.log
python tools/gpu_cold_breakdown.py 2 2>&1 | grep -v Warning | tee gpurun_out/g4_cold.txt
