#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --gpus 1 --steps 1 --warmup 3 --cg-iters 2 --no-e2e --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:"assemble_const|cell_geometry" -s 4 -c 2 -f -o gpurun_out/prof_assemble $B > gpurun_out/ncu_asm.log 2>&1
tail -2 gpurun_out/ncu_asm.log
