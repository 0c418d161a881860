#!/bin/bash
# round 2, call g5: bulk-copy pipelined SpMV against the streaming kernel; parity; cold breakdown
mkdir -p gpurun_out
{
python tools/gpu_time_cg.py 2
FB2_SPMV_KERNEL=stream python tools/gpu_time_cg.py 2
FB2_SPMV_PERSM=2 python tools/gpu_time_cg.py 2
FB2_SPMV_TILE=2048 FB2_SPMV_PERSM=4 python tools/gpu_time_cg.py 2
FB2_SPMV_TILE=3072 python tools/gpu_time_cg.py 2
python tools/gpu_time_cg.py 1
FB2_SPMV_KERNEL=stream python tools/gpu_time_cg.py 1
python tools/gpu_time_cg.py 3
FB2_SPMV_KERNEL=stream python tools/gpu_time_cg.py 3
python tools/gpu_time_cg.py 4
FB2_SPMV_KERNEL=stream python tools/gpu_time_cg.py 4
} 2>&1 | grep -v Warning | tee gpurun_out/g5_tune_cg.txt
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/g5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g5_pytest.log
tail -8 gpurun_out/g5_pytest.log
