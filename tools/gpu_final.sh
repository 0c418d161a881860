#!/bin/bash
# end-of-round capture on one GPU: tests, bench line, ncu launch list, full captures of the dominant kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; tail -c 400 gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv \
    python bench.py --gpus 1 --steps 2 --warmup 3 --cg-iters 10 --no-e2e --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
B="python bench.py --gpus 1 --steps 1 --warmup 3 --cg-iters 4 --no-e2e --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:"assemble_const_v4|cell_geometry4" -s 4 -c 2 -f -o gpurun_out/prof_assemble $B > gpurun_out/ncu_asm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"spmv_stream|cg_update" -s 12 -c 3 -f -o gpurun_out/prof_cg $B > gpurun_out/ncu_cg.log 2>&1
ls -la gpurun_out/*.ncu-rep
