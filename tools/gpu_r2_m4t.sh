#!/bin/bash
# round 2, 4-GPU call: multi-GPU parity tests (2- and 4-rank slab NCCL / peer, Morton partitions with up to three neighbours)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/m4t_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/m4t_pytest.log
tail -5 gpurun_out/m4t_pytest.log
