#!/bin/bash
# round 2, 8-GPU call: BASELINE config 5 (tet P2 322^3, 200 M cells, ONE box in 8 x-slabs) with the converged-solve verify
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/m8_topo.txt 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --config 5 --steps 3 --warmup 3 > gpurun_out/m8_bench_cfg5.json 2> gpurun_out/m8_bench_cfg5.err; echo "bench cfg5 rc=$?"
tail -3 gpurun_out/m8_bench_cfg5.err
python - <<PY
import json
for l in open("gpurun_out/m8_bench_cfg5.json"):
    if l.startswith("{"):
        d=json.loads(l); print(d["value"], d["assembly_ms"], d["problem"], d["cg"], d.get("verify"), d["roofline"]["frac"], d["e2e"])
PY
