#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitizer_probe.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck exit=$?"; tail -3 gpurun_out/sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitizer_probe.py small > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck exit=$?"; tail -3 gpurun_out/sanitizer_racecheck.log
