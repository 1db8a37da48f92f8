#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:score_tc2_kernel -s 7 -c 7 -f -o gpurun_out/prof_q384 python scripts/r02_ncu_small.py 10000000 384 100 tensor > gpurun_out/ncu_q384.log 2>&1; echo "ncu q384 exit=$?"; tail -2 gpurun_out/ncu_q384.log
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:score_tc2_kernel -s 3 -c 3 -f -o gpurun_out/prof_q256 python scripts/r02_ncu_small.py 10000000 256 100 tensor > gpurun_out/ncu_q256.log 2>&1; echo "ncu q256 exit=$?"; tail -2 gpurun_out/ncu_q256.log
