#!/bin/bash
# round 2, second GPU call: full parity suite (incl. the v5 schedule invariants), gather micro-benchmark, v5 tile / warp sweep
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -12 gpurun_out/r2b_pytest.log
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/gather_bench tools/micro/gather_bench.cu && /tmp/gather_bench | tee gpurun_out/r2b_gather_bench.txt
{
FB2_ASM_KERNEL=v4 python tools/gpu_time_asm.py 2
for cap in 2560 3072 3584; do for thr in 96 128; do
  timeout 300 python tools/gpu_time_asm.py 2
done; done
FB2_ASM5_CAP=4096 FB2_ASM5_THREADS=64 timeout 300 python tools/gpu_time_asm.py 2
FB2_ASM5_CAP=2048 FB2_ASM5_THREADS=128 timeout 300 python tools/gpu_time_asm.py 2
FB2_ASM_KERNEL=v4 python tools/gpu_time_asm.py 1
FB2_ASM5_CAP=3072 FB2_ASM5_THREADS=96 timeout 300 python tools/gpu_time_asm.py 1
} 2>&1 | grep -v Warning | tee gpurun_out/r2b_tune_asm5.txt
