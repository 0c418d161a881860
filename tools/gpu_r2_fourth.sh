#!/bin/bash
# round 2, fourth GPU call: v5 on v4's loop (transposed tile) timings + ncu; cfg 1/3/4 launch lists; full parity suite
mkdir -p gpurun_out
{
FB2_ASM_KERNEL=v4 python tools/gpu_time_asm.py 2
python tools/gpu_time_asm.py 2
FB2_ASM5_CAP=2304 python tools/gpu_time_asm.py 2
FB2_ASM5_CAP=2048 python tools/gpu_time_asm.py 2
FB2_ASM_KERNEL=v4 python tools/gpu_time_asm.py 1
python tools/gpu_time_asm.py 1
python tools/gpu_time_asm.py 3
python tools/gpu_time_asm.py 4
} 2>&1 | grep -v Warning | tee gpurun_out/r2d_tune_asm5.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:assemble_const_v5 -s 2 -c 1 -o gpurun_out/r2d_ncu_v5 python tools/gpu_time_asm.py 2 > gpurun_out/r2d_ncu_v5.log 2>&1
ncu -i gpurun_out/r2d_ncu_v5.ncu-rep --page raw --csv > gpurun_out/r2d_ncu_v5_raw.csv 2>/dev/null
ncu -i gpurun_out/r2d_ncu_v5.ncu-rep --page source --csv > gpurun_out/r2d_ncu_v5_source.csv 2>/dev/null
rm -f gpurun_out/r2d_ncu_v5.ncu-rep
for c in 3 4 1; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2d_launches_cfg$c.csv python tools/gpu_time_asm.py $c > /dev/null 2>&1
done
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -8 gpurun_out/r2d_pytest.log
