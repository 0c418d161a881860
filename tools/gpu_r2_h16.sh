#!/bin/bash
# round 2, call h16: Dirichlet matrix kernels with 8 lanes per row; P1 interpolation points; config 4 set-up phases
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "dirichlet or bc or elasticity or poisson or operator" > gpurun_out/h16_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h16_pytest.log
tail -3 gpurun_out/h16_pytest.log
python tools/gpu_cold_cfg34.py 2>&1 | grep -v Warning | tee gpurun_out/h16_cold_cfg34.txt
