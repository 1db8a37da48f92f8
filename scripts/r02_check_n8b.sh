#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8b.json 2> gpurun_out/bench_n8b.err; echo "exit=$? bench n8"; tail -c 600 gpurun_out/bench_n8b.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --steps 20 --warmup 5 --exchange nccl --no-target --no-large > gpurun_out/bench_n8_nccl.json 2> gpurun_out/bench_n8_nccl.err; echo "exit=$? bench n8 nccl"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/bench_ref_n8.json 2> gpurun_out/bench_ref_n8.err; echo "exit=$? ref n8"
