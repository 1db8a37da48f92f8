#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/t_driver.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/t_driver.log
timeout 600 python scripts/r02_probe_modes.py > gpurun_out/probe_modes3.jsonl 2> gpurun_out/probe_modes3.err; echo "modes exit=$?"; cat gpurun_out/probe_modes3.jsonl
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; tail -c 300 gpurun_out/bench.err
