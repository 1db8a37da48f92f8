#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:score_tc2_kernel -s 3 -c 3 -f -o gpurun_out/prof_q64_res python scripts/r02_ncu_small.py 10000000 64 100 tensor > gpurun_out/ncu_q64.log 2>&1; echo "ncu q64 exit=$?"; tail -2 gpurun_out/ncu_q64.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-config1 --no-config4 --large-steps 1 > gpurun_out/b_under_ncu.log 2>&1; echo "launch list exit=$?"
timeout 600 python -m pytest tests/test_search_gpu.py -x -q -m gpu -p no:cacheprovider -k "ingest or zarr or readback or drop_in" > gpurun_out/t_ingest.log 2>&1; echo "ingest tests exit=$?"; tail -3 gpurun_out/t_ingest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; tail -c 600 gpurun_out/bench.err
