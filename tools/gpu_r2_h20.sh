#!/bin/bash
# round 2, call h20: full GPU suite on the tree with the wide face sort / restricted interpolation points
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/h20_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h20_pytest.log
tail -4 gpurun_out/h20_pytest.log
