#!/bin/bash
# round 2, call h6: fused element sum (constant terms folded into the quadrature kernel, accumulate in place) + symmetric
# FMA-chain quadrature kernel: tests, config 3 timing, launch list
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/h6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h6_pytest.log
tail -4 gpurun_out/h6_pytest.log
python tools/gpu_time_asm.py 3 2>&1 | grep -v Warning | tee gpurun_out/h6_time_cfg3.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/h6_launches_cfg3.csv python tools/gpu_time_asm.py 3 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/h6_launches_cfg3.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=r; st=i; break
ki=h.index('Kernel Name'); vi=h.index('Metric Value'); ui=h.index('Metric Unit')
seq=[(r[ki][:60], float(r[vi].replace(',',''))*(1e-3 if r[ui]=='ns' else 1.0)) for r in rows[st+1:] if len(r)>vi]
for name,t in seq[-22:]: print(f"{t:9.1f} us {name}")
PY
