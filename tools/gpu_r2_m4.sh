#!/bin/bash
# round 2, N-GPU call (N from the environment): config-2 weak-scaling line (one (128*N) x 128 x 128 box in N x-slabs), default dist mode
N=${FB2_NGPU:-4}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/m${N}_bench_cfg2.json 2> gpurun_out/m${N}_bench_cfg2.err; echo "bench rc=$?"
tail -2 gpurun_out/m${N}_bench_cfg2.err
python - <<PY
import json
for l in open("gpurun_out/m${N}_bench_cfg2.json"):
    if l.startswith("{"):
        d=json.loads(l); print(d["n_gpus"], d["value"], d["assembly_ms"], d["assembly_ms_steps"], d["cg"]["iters_per_s"], d["roofline"]["frac"], d.get("verify",{}).get("ok"), d["e2e"]["value"], d["e2e"]["cg_iters_per_s"])
PY
