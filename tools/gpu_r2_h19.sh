#!/bin/bash
# round 2, call h19: interpolation points restricted to the requested dofs; Dirichlet set-up cost at config-2 size
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_integration_plugin.py -x -q -m gpu -k "dirichlet or bc or elasticity or poisson or operator or plugin or boundary or geometry or wide or relabelled" > gpurun_out/h19_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h19_pytest.log
tail -3 gpurun_out/h19_pytest.log
python tools/gpu_time_bc.py 128 2>&1 | grep -v Warning | tee gpurun_out/h19_bc_setup.txt
