#!/bin/bash
# 2-GPU validation: multi-GPU parity tests (fused exchange, all-rank overflow retry, single-process master) and
# the bench with its parity / target-config sections
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_n2.txt
timeout 900 python -m pytest tests/test_multigpu_gpu.py -v -m gpu --timeout=600 -p no:cacheprovider > gpurun_out/test_multigpu.log 2>&1; echo "exit=$? test_multigpu"; tail -8 gpurun_out/test_multigpu.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "exit=$? bench n2"; tail -c 1500 gpurun_out/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "exit=$? ref n2"
