#!/bin/bash
# round 2, call h12: launch list of a COLD config-2 assembly (mesh -> symbolic -> schedule -> first assembly) on the current tree
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/h12_launches_cold.csv python tools/gpu_cold_breakdown.py 2 > gpurun_out/h12_cold_under_ncu.txt 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/h12_launches_cold.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=r; st=i; break
ki=h.index('Kernel Name'); vi=h.index('Metric Value'); ui=h.index('Metric Unit')
seq=[(r[ki][:90], float(r[vi].replace(',',''))*(1e-3 if r[ui]=='ns' else 1.0)) for r in rows[st+1:] if len(r)>vi]
tot=0
for name,t in seq[:80]:
    tot+=t
    if t>=30: print(f"{t:9.1f} us {name}")
print("sum of first 80 launches", tot)
PY
python tools/gpu_cold_breakdown.py 2 2>&1 | grep -v Warning | tee gpurun_out/h12_cold.txt
