#!/bin/bash
# round-2 first GPU pass (1 GPU): host facts, smoke, the driver's pytest invocation, bench + reference arm
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
{ nproc; free -g | head -2; nvidia-smi --query-gpu=name,memory.total --format=csv; } > gpurun_out/host.txt 2>&1
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/t_driver.log 2>&1; echo "pytest exit=$?"; tail -5 gpurun_out/t_driver.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; tail -c 1500 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit=$?"
