#!/bin/bash
# What the driver runs at round end, on one GPU: smoke, the gpu-marked tests in one pytest process, both bench arms.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -1 gpurun_out/smoke.log
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/t_driver.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/t_driver.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit=$?"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; tail -c 300 gpurun_out/bench.err
