#!/bin/bash
# round 2, 2-GPU call: NCCL + peer-memory distributed CG parity (verify_slab) and 2-GPU bench lines
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2m_topo.txt 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log
tail -15 gpurun_out/r2m_pytest.log
for mode in nccl peer; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --dist-mode $mode --no-e2e > gpurun_out/r2m_bench2_$mode.json 2> gpurun_out/r2m_bench2_$mode.err; echo "bench $mode rc=$?"
  tail -c 900 gpurun_out/r2m_bench2_$mode.json; echo
  tail -3 gpurun_out/r2m_bench2_$mode.err
done
