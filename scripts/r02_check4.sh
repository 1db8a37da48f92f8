#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/t_driver.log 2>&1; echo "pytest exit=$?"; tail -15 gpurun_out/t_driver.log
timeout 600 python scripts/r02_probe_modes.py > gpurun_out/probe_modes.jsonl 2> gpurun_out/probe_modes.err; echo "modes exit=$?"; cat gpurun_out/probe_modes.jsonl
timeout 600 python scripts/probe_terms.py > gpurun_out/probe_terms.json 2> gpurun_out/probe_terms.err; echo "terms exit=$?"; tail -5 gpurun_out/probe_terms.json
