#!/bin/bash
# round 2, call h5: ncu --set full of the quadrature-loop element kernel (config 3: tri P3, NQ = 12, variable coefficient)
# -- the evidence the north-star asks for before deciding on DMMA for p >= 3 -- and of the tet P3 variant (NQ = 24)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:elem_quad_kernel -s 1 -c 1 -o gpurun_out/h5_ncu_quad_tri python tools/gpu_time_asm.py 3 > gpurun_out/h5_ncu_quad_tri.log 2>&1
ncu -i gpurun_out/h5_ncu_quad_tri.ncu-rep --page raw --csv > gpurun_out/h5_ncu_quad_tri_raw.csv 2>/dev/null
ncu -i gpurun_out/h5_ncu_quad_tri.ncu-rep --page source --csv > gpurun_out/h5_ncu_quad_tri_source.csv 2>/dev/null
rm -f gpurun_out/h5_ncu_quad_tri.ncu-rep
tail -2 gpurun_out/h5_ncu_quad_tri.log
python tools/gpu_time_asm.py 3 2>&1 | grep -v Warning | tee gpurun_out/h5_time_cfg3.txt
