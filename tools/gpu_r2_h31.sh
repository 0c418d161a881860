#!/bin/bash
# round 2, call h31: the reference's own pytorch-CUDA backend ("existing GPU path") at 32^3 and 64^3
mkdir -p gpurun_out
timeout 600 python bench.py --existing-gpu-child --config 2 > gpurun_out/h31_existing_gpu_cfg2.json 2> gpurun_out/h31_existing_gpu_cfg2.err; echo "rc=$?"
tail -2 gpurun_out/h31_existing_gpu_cfg2.json | cut -c1-600
tail -3 gpurun_out/h31_existing_gpu_cfg2.err | cut -c1-300
