#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests/test_search_gpu.py tests/test_fullsize_gpu.py tests/test_stateful_fuzz_gpu.py -x -q -m gpu -p no:cacheprovider > gpurun_out/t_search.log 2>&1; echo "pytest exit=$?"; tail -5 gpurun_out/t_search.log
timeout 900 python scripts/sweep_batch.py > gpurun_out/batch_sweep3.jsonl 2> gpurun_out/batch_sweep3.err; echo "sweep exit=$?"
VODB_RESIDENT=0 timeout 900 python scripts/sweep_batch.py > gpurun_out/batch_sweep3_nores.jsonl 2> gpurun_out/batch_sweep3_nores.err; echo "sweep nores exit=$?"
