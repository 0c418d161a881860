#!/bin/bash
# A/B tuning of the v4 fused assembly kernel on the GPU box: warps per tile (column split), pipeline
# depth (cp.async ring look-ahead) and tile size are run-time knobs; tiles/CTA and min blocks need a rebuild
mkdir -p gpurun_out
bench() {  # $1 = split, $2 = depth, $3 = tile
  FB2_ASM4_SPLIT=$1 FB2_ASM4_DEPTH=$2 FB2_ASM4_TILE=$3 python bench.py --gpus 1 --steps 5 --warmup 3 --cg-iters 2 --no-e2e --no-cpu-baseline 2>gpurun_out/tune_err.txt | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$LABEL warps/tile $1 depth $2 tile $3 asm_ms %.3f nnz/s %.3e' % (d['assembly_ms'], d['value']))" || tail -3 gpurun_out/tune_err.txt
}
{
LABEL="tiles/cta 4 minblocks 2"
bench 1 1 2560
bench 2 1 2560
bench 2 1 2304
bench 2 2 2176
} 2>&1 | tee gpurun_out/tune_asm9.txt
