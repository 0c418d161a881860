#!/bin/bash
# one GPU box session: bench line + ncu launch list (shares) for the same command
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --gpus 1 --steps 1 --warmup 3 --cg-iters 10 --no-e2e --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/launches.csv') if l.startswith('"')))
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[vi].replace(',', ''))
    if r[ui] == 'ns': v /= 1e3
    elif r[ui] == 'ms': v *= 1e3
    a = agg[r[ki][:90]]; a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{t/1e3:10.3f} ms {100*t/tot:5.1f}% x{n:4d}  {k}")
PY
