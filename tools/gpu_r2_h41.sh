#!/bin/bash
# round 2, call h41: CG on the unassembled form (fused matrix-free product) against CG on the assembled matrix, config 2
mkdir -p gpurun_out
python tools/gpu_time_matfree_cg.py 128 104 2>&1 | grep -v Warning | tee gpurun_out/h41_matfree_cg.txt
