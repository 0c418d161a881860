#!/bin/bash
# round 2, call h10: ncu --set full of config 4's fused P1-elasticity gather and of config 1's kernels
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:assemble_from_ke_kernel -s 2 -c 1 -o gpurun_out/h10_ncu_cfg4 python tools/gpu_time_asm.py 4 > gpurun_out/h10_ncu_cfg4.log 2>&1
ncu -i gpurun_out/h10_ncu_cfg4.ncu-rep --page raw --csv > gpurun_out/h10_ncu_cfg4_raw.csv 2>/dev/null
ncu -i gpurun_out/h10_ncu_cfg4.ncu-rep --page source --csv > gpurun_out/h10_ncu_cfg4_source.csv 2>/dev/null
rm -f gpurun_out/h10_ncu_cfg4.ncu-rep
tail -2 gpurun_out/h10_ncu_cfg4.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/h10_launches_cfg1.csv python tools/gpu_time_asm.py 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/h10_launches_cfg1.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=r; st=i; break
ki=h.index('Kernel Name'); vi=h.index('Metric Value'); ui=h.index('Metric Unit')
seq=[(r[ki][:70], float(r[vi].replace(',',''))*(1e-3 if r[ui]=='ns' else 1.0)) for r in rows[st+1:] if len(r)>vi]
for name,t in seq[-12:]: print(f"{t:9.1f} us {name}")
PY
