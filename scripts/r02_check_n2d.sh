#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_multigpu_gpu.py -v -m gpu --timeout=400 -p no:cacheprovider > gpurun_out/test_multigpu.log 2>&1; echo "exit=$? test_multigpu"; tail -12 gpurun_out/test_multigpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-target --large-steps 3 > gpurun_out/bench_n2e.json 2> gpurun_out/bench_n2e.err; echo "exit=$? bench n2"; tail -c 300 gpurun_out/bench_n2e.err
