#!/bin/bash
# round 2, call h15: ncu --set full of the symbolic ranking kernel and the schedule kernel (cold path)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sym_rows_kernel|asm4_schedule_kernel" -c 3 -o gpurun_out/h15_ncu_sym python tools/gpu_cold_breakdown.py 2 > gpurun_out/h15_ncu_sym.log 2>&1
ncu -i gpurun_out/h15_ncu_sym.ncu-rep --page raw --csv > gpurun_out/h15_ncu_sym_raw.csv 2>/dev/null
ncu -i gpurun_out/h15_ncu_sym.ncu-rep --page source --csv > gpurun_out/h15_ncu_sym_source.csv 2>/dev/null
rm -f gpurun_out/h15_ncu_sym.ncu-rep
tail -2 gpurun_out/h15_ncu_sym.log
