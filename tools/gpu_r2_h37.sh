#!/bin/bash
# round 2, call h37: kernel split of the fused matrix-free product (cell kernel vs row-owner gather), config 2
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread --clock-control none -k regex:"matfree_cell_kernel|gather_vector_kernel" -c 4 --csv --log-file gpurun_out/h37_matfree_kernels.csv python tools/gpu_time_matfree.py 2 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/h37_matfree_kernels.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=r; st=i; break
ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); ui=h.index('Metric Unit'); ii=h.index('ID')
for r in rows[st+1:]:
    if len(r)>vi and int(r[ii])>=2: print(r[ii], r[ki][:40], r[mi], r[vi], r[ui])
PY
