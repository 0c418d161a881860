#!/bin/bash
# A/B tuning of the fused assembly kernel on the GPU box (rebuilds the library per variant)
mkdir -p gpurun_out
run() {  # $1 = minblocks, $2 = tile
  make -C fealpy_b200/csrc clean >/dev/null; make -C fealpy_b200/csrc -j16 EXTRA="-DFB2_ASM3_MINBLOCKS=$1" >/dev/null 2>&1
  FB2_ASM_TILE=$2 python bench.py --gpus 1 --steps 5 --warmup 3 --cg-iters 2 --no-e2e --no-cpu-baseline 2>gpurun_out/tune_err.txt | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('minblocks $1 tile $2 asm_ms %.3f nnz/s %.3e' % (d['assembly_ms'], d['value']))"
}
for mb in 4 5 6; do for tile in 2048 3072 4096 6144; do run $mb $tile; done; done 2>&1 | tee gpurun_out/tune_asm.txt
make -C fealpy_b200/csrc clean >/dev/null; make -C fealpy_b200/csrc -j16 >/dev/null 2>&1
