#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_multigpu_gpu.py -v -m gpu --timeout=600 -p no:cacheprovider > gpurun_out/test_multigpu.log 2>&1; echo "exit=$? test_multigpu"; tail -5 gpurun_out/test_multigpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-large --target-rows 20000000 > gpurun_out/bench_n2b.json 2> gpurun_out/bench_n2b.err; echo "exit=$? bench n2"; tail -c 600 gpurun_out/bench_n2b.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 2 --steps 10 --warmup 3 --exchange nccl --no-target --no-large > gpurun_out/bench_n2_nccl.json 2> gpurun_out/bench_n2_nccl.err; echo "exit=$? bench n2 nccl"
