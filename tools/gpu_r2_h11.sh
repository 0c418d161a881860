#!/bin/bash
# round 2, call h11: fused P1 elasticity gather with vertex-quarter records (two 256-bit loads per pair)
mkdir -p gpurun_out
python tools/gpu_time_asm.py 4 2>&1 | grep -v Warning | tee gpurun_out/h11_time_cfg4.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "elasticity or golden or tensor or paths" > gpurun_out/h11_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h11_pytest.log
tail -3 gpurun_out/h11_pytest.log
