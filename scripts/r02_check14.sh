#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/t_driver.log 2>&1; echo "pytest exit=$?"; tail -4 gpurun_out/t_driver.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitizer_probe.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck exit=$?"; tail -4 gpurun_out/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitizer_probe.py small > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck exit=$?"; tail -4 gpurun_out/sanitizer_racecheck.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; tail -c 400 gpurun_out/bench.err
