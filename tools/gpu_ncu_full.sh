#!/bin/bash
# full ncu capture of the two dominant kernels (one GPU); reports come back in gpurun_out/
mkdir -p gpurun_out
B="python bench.py --gpus 1 --steps 1 --warmup 3 --cg-iters 4 --no-e2e --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:assemble_const -s 2 -c 1 -f -o gpurun_out/prof_assemble $B > gpurun_out/ncu_asm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmv_stream -s 6 -c 1 -f -o gpurun_out/prof_spmv $B > gpurun_out/ncu_spmv.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/ncu_asm.log gpurun_out/ncu_spmv.log
