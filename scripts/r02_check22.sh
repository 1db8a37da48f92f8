#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python scripts/r02_probe_idle.py > gpurun_out/probe_idle.json 2> gpurun_out/probe_idle.err; echo "idle exit=$?"; cat gpurun_out/probe_idle.json; tail -2 gpurun_out/probe_idle.err
