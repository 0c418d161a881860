#!/bin/bash
# round 2, call h34: matrix-free product pinned to the reference run (all golden cases, scalar + tensor spaces); full GPU suite
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/h34_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h34_pytest.log
tail -4 gpurun_out/h34_pytest.log
