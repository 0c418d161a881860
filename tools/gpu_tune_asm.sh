#!/bin/bash
# A/B tuning of the v4 fused assembly kernel on the GPU box: pipeline depth (cp.async ring look-ahead)
# and warp-tile size are run-time knobs; warps/CTA, min blocks and column chunk need a rebuild
mkdir -p gpurun_out
bench() {  # $1 = label, $2 = depth, $3 = tile
  FB2_ASM4_DEPTH=$2 FB2_ASM4_TILE=$3 python bench.py --gpus 1 --steps 5 --warmup 3 --cg-iters 2 --no-e2e --no-cpu-baseline 2>gpurun_out/tune_err.txt | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1 depth $2 tile $3 asm_ms %.3f nnz/s %.3e' % (d['assembly_ms'], d['value']))" || tail -3 gpurun_out/tune_err.txt
}
{
for t in 2048 2176; do bench "jc 5" 2 $t; done
make -C fealpy_b200/csrc clean >/dev/null; make -C fealpy_b200/csrc -j16 EXTRA="-DFB2_ASM4_JC=10" >/dev/null 2>&1
bench "jc 10" 1 2560
bench "jc 10" 2 2176
make -C fealpy_b200/csrc clean >/dev/null; make -C fealpy_b200/csrc -j16 EXTRA="-DFB2_ASM4_JC=2" >/dev/null 2>&1
bench "jc 2" 1 2560
} 2>&1 | tee gpurun_out/tune_asm8.txt
make -C fealpy_b200/csrc clean >/dev/null; make -C fealpy_b200/csrc -j16 >/dev/null 2>&1
