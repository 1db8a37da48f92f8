#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python scripts/r02_sweep_select.py > gpurun_out/sweep_select.jsonl 2> gpurun_out/sweep_select.err; echo "select sweep exit=$?"; cat gpurun_out/sweep_select.jsonl
timeout 900 python scripts/r02_probe_ingest.py > gpurun_out/probe_ingest.json 2> gpurun_out/probe_ingest.err; echo "ingest exit=$?"; cat gpurun_out/probe_ingest.json; tail -3 gpurun_out/probe_ingest.err
