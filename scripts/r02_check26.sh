#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/t_fastsel2.log 2>&1; echo "pytest exit=$?"; tail -5 gpurun_out/t_fastsel2.log
timeout 400 python scripts/r02_probe_k1000.py > gpurun_out/k1000_probe2.jsonl 2>&1; cat gpurun_out/k1000_probe2.jsonl
