#!/bin/bash
# round 2, 2-GPU call: NCCL + peer-memory CG parity (slab), Morton-partition parity on relabelled meshes, 2-GPU bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/m2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/m2_pytest.log
tail -6 gpurun_out/m2_pytest.log
for mode in peer nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --dist-mode $mode --no-e2e > gpurun_out/m2_bench2_$mode.json 2> gpurun_out/m2_bench2_$mode.err; echo "bench $mode rc=$?"
  python - <<PY
import json
for l in open("gpurun_out/m2_bench2_$mode.json"):
    if l.startswith("{"):
        d=json.loads(l); print("$mode", d["value"], d["cg"], d.get("verify"))
PY
  tail -2 gpurun_out/m2_bench2_$mode.err
done
