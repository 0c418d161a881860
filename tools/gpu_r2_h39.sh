#!/bin/bash
# round 2, call h39: matrix-free tests + timing on the final library (rolled row loop in the cell kernel)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "matfree or matrix_free or operator or bilinear_form_matmul" > gpurun_out/h39_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h39_pytest.log
tail -3 gpurun_out/h39_pytest.log
{ python tools/gpu_time_matfree.py 2; python tools/gpu_time_matfree.py 1; python tools/gpu_time_matfree.py 2 64; } 2>&1 | grep -v Warning | tee gpurun_out/h39_matfree.txt
