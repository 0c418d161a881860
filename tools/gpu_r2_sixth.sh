#!/bin/bash
# round 2, sixth GPU call: v4 with the deeper entry prefetch; full parity suite; default bench + cfg 4
mkdir -p gpurun_out
{
python tools/gpu_time_asm.py 2
python tools/gpu_time_asm.py 1
FB2_ASM_KERNEL=v4 python tools/gpu_time_asm.py 1
python tools/gpu_time_asm.py 4 64
} 2>&1 | grep -v Warning | tee gpurun_out/r2f_tune_asm.txt
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -8 gpurun_out/r2f_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_bench_cfg2.json 2> gpurun_out/r2f_bench_cfg2.err; echo "bench2 rc=$?"
timeout 600 python bench.py --config 4 --steps 5 --warmup 3 > gpurun_out/r2f_bench_cfg4.json 2> gpurun_out/r2f_bench_cfg4.err; echo "bench4 rc=$?"
tail -c 600 gpurun_out/r2f_bench_cfg4.json
