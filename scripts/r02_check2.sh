#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python scripts/r02_probe_modes.py > gpurun_out/probe_modes.jsonl 2> gpurun_out/probe_modes.err; echo "modes exit=$?"; cat gpurun_out/probe_modes.jsonl
timeout 2400 python scripts/r02_sweep_schedule.py > gpurun_out/sweep_schedule.jsonl 2> gpurun_out/sweep_schedule.err; echo "sweep exit=$?"
