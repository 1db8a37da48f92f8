#!/bin/bash
# ncu source-level capture of the select kernels of a 1.25M-row, 64-query search (first select: 16384-entry dump list)
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 400 ncu --set full --import-source on --clock-control none --kernel-name regex:select_kernel --launch-skip 4 --launch-count 2 \
  -o gpurun_out/select_fast -f python scripts/r02_ncu_small.py > gpurun_out/select_fast_ncu.log 2>&1; echo "ncu exit=$?"; tail -3 gpurun_out/select_fast_ncu.log
ncu -i gpurun_out/select_fast.ncu-rep --page raw --csv > gpurun_out/select_fast_raw.csv 2>/dev/null
ncu -i gpurun_out/select_fast.ncu-rep --page source --csv > gpurun_out/select_fast_source.csv 2>/dev/null
ls -la gpurun_out | grep select_fast
