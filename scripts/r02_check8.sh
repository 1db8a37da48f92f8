#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/t_driver.log 2>&1; echo "pytest exit=$?"; tail -5 gpurun_out/t_driver.log
timeout 900 python scripts/r02_sweep_schedule2.py > gpurun_out/sweep_schedule2.jsonl 2> gpurun_out/sweep_schedule2.err; echo "sched exit=$?"
timeout 900 python scripts/sweep_batch.py > gpurun_out/batch_sweep4.jsonl 2> gpurun_out/batch_sweep4.err; echo "sweep exit=$?"
timeout 600 python scripts/r02_probe_modes.py > gpurun_out/probe_modes2.jsonl 2> gpurun_out/probe_modes2.err; echo "modes exit=$?"
