#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python scripts/r02_probe_k1000.py > gpurun_out/probe_k1000.jsonl 2> gpurun_out/probe_k1000.err; echo "k1000 exit=$?"; cat gpurun_out/probe_k1000.jsonl; tail -3 gpurun_out/probe_k1000.err
timeout 1500 python -m pytest tests/test_search_gpu.py tests/test_merge_gpu.py tests/test_fullsize_gpu.py -x -q -m gpu -p no:cacheprovider > gpurun_out/t_search.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/t_search.log
