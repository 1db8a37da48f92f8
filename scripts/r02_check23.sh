#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python scripts/r02_probe_idle2.py > gpurun_out/probe_idle2.json 2> gpurun_out/probe_idle2.err; echo "idle2 exit=$?"; cat gpurun_out/probe_idle2.json; tail -2 gpurun_out/probe_idle2.err
