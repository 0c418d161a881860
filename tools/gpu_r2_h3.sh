#!/bin/bash
# round 2, call h3: SpMV v3 (next header, row pointers, x[r] issued one phase early by plain loads) against v1; CG + SpMV tests; ncu of v2
mkdir -p gpurun_out
{
for c in 2 1 3 4; do
python tools/gpu_time_cg.py $c
FB2_SPMV_KERNEL=1 python tools/gpu_time_cg.py $c
done
} 2>&1 | grep -v Warning | tee gpurun_out/h3_tune_cg.txt
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/h3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h3_pytest.log
tail -3 gpurun_out/h3_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_stream3 -s 20 -c 1 -o gpurun_out/h3_ncu_spmv2 python tools/gpu_time_cg.py 2 > gpurun_out/h3_ncu_spmv2.log 2>&1
ncu -i gpurun_out/h3_ncu_spmv2.ncu-rep --page raw --csv > gpurun_out/h3_ncu_spmv2_raw.csv 2>/dev/null
ncu -i gpurun_out/h3_ncu_spmv2.ncu-rep --page source --csv > gpurun_out/h3_ncu_spmv2_source.csv 2>/dev/null
rm -f gpurun_out/h3_ncu_spmv2.ncu-rep
tail -2 gpurun_out/h3_ncu_spmv2.log
