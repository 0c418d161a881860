#!/bin/bash
# A/B tuning of the v4 fused assembly kernel on the GPU box: pipeline depth (0 = register prefetch,
# D >= 1 = cp.async rings) and warp-tile size are run-time knobs; warps/CTA and min blocks need a rebuild
mkdir -p gpurun_out
bench() {  # $1 = label, $2 = depth, $3 = tile
  FB2_ASM4_DEPTH=$2 FB2_ASM4_TILE=$3 python bench.py --gpus 1 --steps 5 --warmup 3 --cg-iters 2 --no-e2e --no-cpu-baseline 2>gpurun_out/tune_err.txt | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1 depth $2 tile $3 asm_ms %.3f nnz/s %.3e' % (d['assembly_ms'], d['value']))" || tail -3 gpurun_out/tune_err.txt
}
{

for t in 2048 2304 2432; do bench "warps 4 minblocks 2" 1 $t; done
for t in 1536 1792; do bench "warps 4 minblocks 2" 2 $t; done
} 2>&1 | tee gpurun_out/tune_asm6.txt
