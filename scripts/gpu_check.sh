#!/bin/bash
# Runs on the GPU box through gpurun: smoke, the GPU parity tests (one pytest process per file so that a trapped
# kernel cannot poison the rest), a short bench, then the ncu passes of B200_PROFILING.md. Logs -> gpurun_out/.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
STEP=${1:-all}
run() { # name timeout cmd...
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "exit=$? ($name)" | tee -a gpurun_out/summary.txt
  tail -n 15 "gpurun_out/$name.log" | tee -a gpurun_out/summary.txt
}
echo "##### $(date) step=$STEP" >> gpurun_out/summary.txt
if [[ $STEP == all || $STEP == tests ]]; then
  run smoke 300 python __graft_entry__.py --smoke
  run t_search 900 python -m pytest tests/test_search_gpu.py -q -m gpu --timeout=300 -p no:cacheprovider
  run t_sampling 900 python -m pytest tests/test_sampling_gpu.py -q -m gpu --timeout=300 -p no:cacheprovider
  run t_merge 600 python -m pytest tests/test_merge_gpu.py -q -m gpu --timeout=300 -p no:cacheprovider
  run t_fullsize 1200 python -m pytest tests/test_fullsize_gpu.py -q -m gpu --timeout=600 -p no:cacheprovider
fi
if [[ $STEP == all || $STEP == driver ]]; then
  # exactly what the driver runs at round end: one pytest process over the whole gpu suite
  run t_driver 1800 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider
fi
if [[ $STEP == all || $STEP == bench ]]; then
  run bench 900 python bench.py --steps 20 --warmup 5
  run bench_ref 600 python bench.py --impl reference --steps 5 --warmup 2
fi
if [[ $STEP == ncu256 ]]; then
  run ncu_full256 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:score_tc2_kernel -s 7 -c 7 -o gpurun_out/prof_score_tc256 python bench.py --steps 1 --warmup 1 --no-cpu --no-config4 --large-steps 1
fi
if [[ $STEP == all || $STEP == ncu ]]; then
  run ncu_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --large-steps 1
  run ncu_full 900 ncu --set full --clock-control none --import-source on -k regex:score_tc_kernel -s 4 -c 4 -o gpurun_out/prof_score_tc python bench.py --steps 2 --warmup 1 --no-cpu --no-large
  run ncu_full256 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:score_tc2_kernel -s 7 -c 7 -o gpurun_out/prof_score_tc256 python bench.py --steps 1 --warmup 1 --no-cpu --no-config4 --large-steps 1
fi
if [[ $STEP == all || $STEP == ncu_aux ]]; then
  run ncu_aux 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:score_exact|select_kernel|merge|sample_kernel|match_labels|gather_picks' -c 60 -o gpurun_out/prof_aux python scripts/ncu_aux_probe.py
fi
if [[ $STEP == all || $STEP == sanitizer ]]; then
  run sanitizer_memcheck 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitizer_probe.py
  run sanitizer_racecheck 600 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitizer_probe.py small
fi
echo "=== done" | tee -a gpurun_out/summary.txt
