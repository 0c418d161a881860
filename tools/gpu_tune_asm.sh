#!/bin/bash
# A/B tuning of the v4 fused assembly kernel on the GPU box (rebuilds the library per variant)
mkdir -p gpurun_out
run() {  # $1 = warps, $2 = minblocks, $3 = tile
  make -C fealpy_b200/csrc clean >/dev/null; make -C fealpy_b200/csrc -j16 EXTRA="-DFB2_ASM4_WARPS=$1 -DFB2_ASM4_MINBLOCKS=$2" >/dev/null 2>&1
  FB2_ASM4_TILE=$3 python bench.py --gpus 1 --steps 5 --warmup 3 --cg-iters 2 --no-e2e --no-cpu-baseline 2>gpurun_out/tune_err.txt | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('loop warps $1 minblocks $2 tile $3 asm_ms %.3f nnz/s %.3e' % (d['assembly_ms'], d['value']))"
}
for cfg in "4 4" "4 3" "8 2" "4 5"; do for tile in 1536 2560; do run $cfg $tile; done; done 2>&1 | tee gpurun_out/tune_asm.txt
make -C fealpy_b200/csrc clean >/dev/null; make -C fealpy_b200/csrc -j16 >/dev/null 2>&1
