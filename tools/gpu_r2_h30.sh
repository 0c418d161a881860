#!/bin/bash
# round 2, call h30: ncu --set full of the FINAL assembly kernel (v6 with L2 policies), geometry kernel and SpMV kernel
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"assemble_const_v6|cell_geometry4" -s 4 -c 2 -o gpurun_out/h30_ncu_asm python tools/gpu_time_asm.py 2 > gpurun_out/h30_ncu_asm.log 2>&1
ncu -i gpurun_out/h30_ncu_asm.ncu-rep --page raw --csv > gpurun_out/h30_ncu_asm_raw.csv 2>/dev/null
rm -f gpurun_out/h30_ncu_asm.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spmv_stream_kernel|cg_update" -s 30 -c 3 -o gpurun_out/h30_ncu_cg python tools/gpu_time_cg.py 2 > gpurun_out/h30_ncu_cg.log 2>&1
ncu -i gpurun_out/h30_ncu_cg.ncu-rep --page raw --csv > gpurun_out/h30_ncu_cg_raw.csv 2>/dev/null
rm -f gpurun_out/h30_ncu_cg.ncu-rep
tail -2 gpurun_out/h30_ncu_asm.log gpurun_out/h30_ncu_cg.log
