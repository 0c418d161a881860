#!/bin/bash
mkdir -p gpurun_out
run() {  # $1 = extra flags, $2 = tile
  make -C fealpy_b200/csrc clean >/dev/null; make -C fealpy_b200/csrc -j16 EXTRA="$1" >/dev/null 2>&1
  FB2_SPMV_TILE=$2 python bench.py --gpus 1 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/tune_err.txt | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('flags [$1] tile $2 cg ms/it %.3f it/s %.1f frac %.3f' % (d['cg']['ms_per_iter'], d['cg']['iters_per_s'], d['roofline']['frac']))"
}
for t in 2560 3072; do run "" $t; done
for t in 1280 1536 2048 2560; do run "-DFB2_ST_THREADS=128" $t; done
make -C fealpy_b200/csrc clean >/dev/null; make -C fealpy_b200/csrc -j16 >/dev/null 2>&1
