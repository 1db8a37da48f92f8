#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python scripts/r02_probe_multigpu_store.py > gpurun_out/probe_multigpu_store.json 2> gpurun_out/probe_multigpu_store.err; echo "exit=$?"; cat gpurun_out/probe_multigpu_store.json; tail -3 gpurun_out/probe_multigpu_store.err
