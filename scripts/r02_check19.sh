#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for i in 1 2; do
  timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/t_driver_$i.log 2>&1; echo "pytest run $i exit=$?"; tail -2 gpurun_out/t_driver_$i.log
done
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"
