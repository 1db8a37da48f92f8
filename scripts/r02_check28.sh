#!/bin/bash
# resident three-term queries (score_tc2_kernel<64,3,resident>) against the streaming variant, then the gpu tests
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for rt in 0 1; do
  echo "VODB_RESIDENT_TERMS=$rt"
  VODB_RESIDENT_TERMS=$rt timeout 300 python scripts/r02_probe_modes.py 2> gpurun_out/modes_$rt.err | tee gpurun_out/modes_rt$rt.jsonl
done
timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/t_resterms.log 2>&1; echo "pytest exit=$?"; tail -4 gpurun_out/t_resterms.log
