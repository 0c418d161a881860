#!/bin/bash
# round 2, call h38: matrix-free cell kernel: unroll of the row loop (register count / occupancy)
mkdir -p gpurun_out
V=fealpy_b200/csrc/build/variants
{
python tools/gpu_time_matfree.py 2 | head -1
for v in mfu1 mfu2 mfu5; do echo "variant $v"; FB2_LIB_PATH=$V/$v.so python tools/gpu_time_matfree.py 2 | head -1; done
} 2>&1 | grep -v Warning | tee gpurun_out/h38_matfree_unroll.txt
