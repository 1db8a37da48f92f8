#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:score_tc_kernel -s 3 -c 3 -f -o gpurun_out/prof_small_shard python scripts/r02_ncu_small.py 1250000 64 100 tensor > gpurun_out/ncu_small.log 2>&1; echo "ncu small exit=$?"; tail -3 gpurun_out/ncu_small.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 60 --csv --log-file gpurun_out/launches_small.csv python scripts/r02_ncu_small.py 1250000 64 100 tensor > /dev/null 2>&1; echo "launches exit=$?"
