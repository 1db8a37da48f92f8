#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:score_tc2_kernel -s 1 -c 3 -f -o gpurun_out/prof_t3_q64 python scripts/r02_ncu_small.py 10000000 64 100 tensor3 > gpurun_out/ncu_t3.log 2>&1; echo "ncu t3 exit=$?"; tail -2 gpurun_out/ncu_t3.log
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:score_tc2_kernel -s 9 -c 5 -f -o gpurun_out/prof_pair_q8192 python scripts/r02_ncu_small.py 10000000 8192 100 tensor > gpurun_out/ncu_8192.log 2>&1; echo "ncu 8192 exit=$?"; tail -2 gpurun_out/ncu_8192.log
