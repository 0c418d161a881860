#!/bin/bash
# round 2, call h1: fused matrix-free product (tests + timing), full GPU test suite, ncu of the quadrature-loop element kernel (config 3)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "matfree or bilinear_form_matmul or operator" > gpurun_out/h1_pytest_matfree.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h1_pytest_matfree.log
tail -4 gpurun_out/h1_pytest_matfree.log
{
python tools/gpu_time_matfree.py 2
python tools/gpu_time_matfree.py 2 64
python tools/gpu_time_matfree.py 1
} 2>&1 | grep -v Warning | tee gpurun_out/h1_matfree.txt
timeout 1800 python -m pytest tests -x -q -m gpu -rs > gpurun_out/h1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h1_pytest.log
tail -5 gpurun_out/h1_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"elem_scalar_quad|assemble_from_ke" -s 2 -c 2 -o gpurun_out/h1_ncu_cfg3 python tools/gpu_time_asm.py 3 > gpurun_out/h1_ncu_cfg3.log 2>&1
ncu -i gpurun_out/h1_ncu_cfg3.ncu-rep --page raw --csv > gpurun_out/h1_ncu_cfg3_raw.csv 2>/dev/null
ncu -i gpurun_out/h1_ncu_cfg3.ncu-rep --page source --csv > gpurun_out/h1_ncu_cfg3_source.csv 2>/dev/null
rm -f gpurun_out/h1_ncu_cfg3.ncu-rep
tail -3 gpurun_out/h1_ncu_cfg3.log
