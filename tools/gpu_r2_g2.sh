#!/bin/bash
# round 2, call g2: v6 with 256-bit geometry loads; sanitizer passes (memcheck + racecheck) on the final kernels
mkdir -p gpurun_out
V=fealpy_b200/csrc/build/variants
{
python tools/gpu_time_asm.py 2
FB2_LIB_PATH=$V/w3.so python tools/gpu_time_asm.py 2
FB2_LIB_PATH=$V/jc10.so python tools/gpu_time_asm.py 2
python tools/gpu_time_asm.py 1
python tools/gpu_time_asm.py 2 64
} 2>&1 | grep -v Warning | tee gpurun_out/g2_tune_asm.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:assemble_const_v6 -s 2 -c 1 -o gpurun_out/g2_ncu_v6 python tools/gpu_time_asm.py 2 > gpurun_out/g2_ncu_v6.log 2>&1
ncu -i gpurun_out/g2_ncu_v6.ncu-rep --page raw --csv > gpurun_out/g2_ncu_v6_raw.csv 2>/dev/null
ncu -i gpurun_out/g2_ncu_v6.ncu-rep --page source --csv > gpurun_out/g2_ncu_v6_source.csv 2>/dev/null
rm -f gpurun_out/g2_ncu_v6.ncu-rep
SEL='assembly_csr or spmv_and_cg or poisson_source or batched or slab or sort_and_scan or reference or schedule or dirichlet or elasticity'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "$SEL" > gpurun_out/g2_memcheck.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/g2_memcheck.txt
tail -5 gpurun_out/g2_memcheck.txt
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "$SEL" > gpurun_out/g2_racecheck.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/g2_racecheck.txt
tail -5 gpurun_out/g2_racecheck.txt
