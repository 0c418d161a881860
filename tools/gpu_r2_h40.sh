#!/bin/bash
# round 2, call h40: matrix-free product with the cell products written in adjacency order (contiguous per-dof sum) against pair order + gather
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "matfree or matrix_free or operator or bilinear_form_matmul" > gpurun_out/h40_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h40_pytest.log
tail -3 gpurun_out/h40_pytest.log
{ python tools/gpu_time_matfree.py 2 | head -2; FB2_MATFREE_ORDER=pair python tools/gpu_time_matfree.py 2 | head -1; python tools/gpu_time_matfree.py 1 | head -1; FB2_MATFREE_ORDER=pair python tools/gpu_time_matfree.py 1 | head -1; } 2>&1 | grep -v "Warning\|Broken\|Traceback\|File\|print" | tee gpurun_out/h40_matfree.txt
