#!/bin/bash
# round 2, third GPU call: v5 (fixed write-back) timings at 6 / 8 warps; ncu --set full with source of the v4 and v5 numeric kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "asm5 or elasticity_dirichlet or operator_cg or geometry_vectors or user_integrator or against_oracle or share_pattern or source_dirichlet or slab" > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -8 gpurun_out/r2c_pytest.log
{
FB2_ASM_KERNEL=v4 python tools/gpu_time_asm.py 2
FB2_ASM5_CAP=2304 FB2_ASM5_THREADS=128 timeout 300 python tools/gpu_time_asm.py 2
FB2_ASM5_CAP=2432 FB2_ASM5_THREADS=128 timeout 300 python tools/gpu_time_asm.py 2
FB2_ASM5_CAP=3072 FB2_ASM5_THREADS=96 timeout 300 python tools/gpu_time_asm.py 2
FB2_ASM5_CAP=3584 FB2_ASM5_THREADS=96 timeout 300 python tools/gpu_time_asm.py 2
} 2>&1 | grep -v Warning | tee gpurun_out/r2c_tune_asm5.txt
for k in v4 v5; do
  FB2_ASM_KERNEL=$k FB2_ASM5_CAP=2304 FB2_ASM5_THREADS=128 timeout 600 ncu --set full --clock-control none --import-source on -k regex:assemble_const_v -s 2 -c 1 \
     -o gpurun_out/r2c_ncu_$k python tools/gpu_time_asm.py 2 > gpurun_out/r2c_ncu_$k.log 2>&1
  ncu -i gpurun_out/r2c_ncu_$k.ncu-rep --page raw --csv > gpurun_out/r2c_ncu_${k}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r2c_ncu_$k.ncu-rep --page source --csv > gpurun_out/r2c_ncu_${k}_source.csv 2>/dev/null
  ls -la gpurun_out/r2c_ncu_$k.ncu-rep
  sz=$(stat -c %s gpurun_out/r2c_ncu_$k.ncu-rep); if [ "$sz" -gt 25000000 ]; then rm gpurun_out/r2c_ncu_$k.ncu-rep; fi
done
