#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:score_exact|score_tc_kernel|select_kernel|merge_kernel|sample_kernel|match_labels|gather_picks|split_planes' -c 60 -f -o gpurun_out/prof_aux python scripts/ncu_aux_probe.py > gpurun_out/ncu_aux.log 2>&1; echo "ncu aux exit=$?"; tail -2 gpurun_out/ncu_aux.log
timeout 900 python -m pytest tests/test_search_gpu.py tests/test_merge_gpu.py tests/test_sampling_gpu.py -x -q -m gpu -p no:cacheprovider > gpurun_out/t_search.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/t_search.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 40 --csv --log-file gpurun_out/launches_small2.csv python scripts/r02_ncu_small.py 1250000 64 100 tensor > /dev/null 2>&1; echo "launches exit=$?"
