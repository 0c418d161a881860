#!/bin/bash
# round 2, call h33: does the clock sampler's period perturb the timed CG?  same box, 30 ms / 100 ms / 30 ms / 100 ms, then the plain CG timer
mkdir -p gpurun_out
for p in 0.03 0.1 0.03 0.1; do
FB2_BENCH_SAMPLER_S=$p timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-existing-gpu 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('sampler $p', f\"{d['value']:.4e}\", round(d['assembly_ms'],3), round(d['cg']['iters_per_s'],1), d['clocks'])"
done | tee gpurun_out/h33_sampler.txt
python tools/gpu_time_cg.py 2 2>&1 | grep -v Warning | tee -a gpurun_out/h33_sampler.txt
