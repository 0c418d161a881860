#!/bin/bash
# round 2, call g7: SpMV tile->CTA map (SM-chunked vs strided); closed-form box numbering test; cold breakdown
mkdir -p gpurun_out
{
python tools/gpu_time_cg.py 2
FB2_SPMV_MAP=strided python tools/gpu_time_cg.py 2
python tools/gpu_time_cg.py 1
FB2_SPMV_MAP=strided python tools/gpu_time_cg.py 1
python tools/gpu_time_cg.py 3
FB2_SPMV_MAP=strided python tools/gpu_time_cg.py 3
python tools/gpu_time_cg.py 4
FB2_SPMV_MAP=strided python tools/gpu_time_cg.py 4
} 2>&1 | grep -v Warning | tee gpurun_out/g7_tune_cg.txt
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/g7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g7_pytest.log
tail -5 gpurun_out/g7_pytest.log
python tools/gpu_cold_breakdown.py 2 2>&1 | grep -v Warning | tee gpurun_out/g7_cold.txt
