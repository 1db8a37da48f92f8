#!/bin/bash
# 8-GPU validation: multi-GPU parity tests (4 ranks), bench at N=8 (parity + 100M-row target sections) and N=4
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_n8.txt
timeout 900 python -m pytest tests/test_multigpu_gpu.py -v -m gpu --timeout=600 -p no:cacheprovider > gpurun_out/test_multigpu_n8.log 2>&1; echo "exit=$? test_multigpu"; tail -5 gpurun_out/test_multigpu_n8.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "exit=$? bench n8"; tail -c 800 gpurun_out/bench_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 5 --no-target > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "exit=$? bench n4"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --steps 20 --warmup 5 --exchange nccl --no-target --no-large > gpurun_out/bench_n8_nccl.json 2> gpurun_out/bench_n8_nccl.err; echo "exit=$? bench n8 nccl"
