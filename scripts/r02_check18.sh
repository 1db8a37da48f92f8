#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/t_driver.log 2>&1; echo "pytest exit=$?"; tail -6 gpurun_out/t_driver.log
timeout 900 python scripts/sweep_batch.py > gpurun_out/batch_sweep_wide.jsonl 2> gpurun_out/batch_sweep_wide.err; echo "sweep wide exit=$?"
VODB_WIDE=0 timeout 900 python scripts/sweep_batch.py > gpurun_out/batch_sweep_nowide.jsonl 2> gpurun_out/batch_sweep_nowide.err; echo "sweep nowide exit=$?"
