#!/bin/bash
# SpMV knobs that need no rebuild: resident CTAs per SM of the persistent grid, tile size
mkdir -p gpurun_out
run() {  # $1 = CTAs per SM, $2 = tile
  FB2_SPMV_PERSM=$1 FB2_SPMV_TILE=$2 python bench.py --gpus 1 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/tune_err.txt | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ctas/sm $1 tile $2 cg ms/it %.3f it/s %.1f frac %.3f' % (d['cg']['ms_per_iter'], d['cg']['iters_per_s'], d['roofline']['frac']))"
}
for c in 8 7 6 5 4; do run $c 2560; done 2>&1 | tee gpurun_out/tune_spmv5.txt
run 8 2304 | tee -a gpurun_out/tune_spmv5.txt
