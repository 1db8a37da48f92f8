#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/t_driver.log 2>&1; echo "pytest exit=$?"; tail -5 gpurun_out/t_driver.log
timeout 900 python scripts/sweep_batch.py > gpurun_out/batch_sweep.jsonl 2> gpurun_out/batch_sweep.err; echo "sweep exit=$?"
VODB_RASTER=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-config4 --no-config1 > gpurun_out/bench_noraster.json 2> gpurun_out/bench_noraster.err; echo "bench noraster exit=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-config4 --no-config1 > gpurun_out/bench_raster.json 2> gpurun_out/bench_raster.err; echo "bench raster exit=$?"
