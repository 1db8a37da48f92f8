#!/bin/bash
# ncu over rank 0 of a two-rank search through the fused exchange (2-GPU box); rank 1 runs unprofiled beside it.
#   pass "launches": launch list of rank 0's searches (gpu__time_duration of every kernel)
#   pass "nvlink":   NVLink byte counters of rank 0's select / merge kernels
#   pass "full":     --set full of the same kernels
# usage: r02_ncu_exchange.sh [pass ...]   (default: all three)
set -u
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1 WORLD_SIZE=2
run_pair() {  # $1 = master port, rest = rank 0's command prefix
  local port=$1; shift
  MASTER_PORT=$port RANK=1 timeout 150 python scripts/r02_ncu_exchange_probe.py > gpurun_out/xchg_rank1_$port.log 2>&1 &
  local peer=$!
  MASTER_PORT=$port RANK=0 timeout 150 "$@" python scripts/r02_ncu_exchange_probe.py
  local rc=$?
  if [ $rc -ne 0 ]; then sleep 2; kill $peer 2>/dev/null; fi   # rank 1 would wait for a peer that is gone
  wait $peer; echo "pair $port: rank0 rc=$rc rank1 rc=$?"
  tail -2 gpurun_out/xchg_rank1_$port.log | cut -c1-200
}
FILTER='regex:select_kernel|merge_exchange_kernel'
for pass in "${@:-launches nvlink full}"; do for p in $pass; do
  case $p in
    launches)
      run_pair 29612 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/xchg_launches_rank0.csv \
        > gpurun_out/xchg_launches.log 2>&1; tail -3 gpurun_out/xchg_launches.log ;;
    nvlink)
      run_pair 29613 ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum \
        --clock-control none --cache-control none --kernel-name "$FILTER" --launch-skip 3 --launch-count 12 --csv \
        --log-file gpurun_out/xchg_nvlink_rank0.csv > gpurun_out/xchg_nvlink.log 2>&1; tail -5 gpurun_out/xchg_nvlink.log ;;
    full)
      run_pair 29614 ncu --set full --clock-control none --cache-control none --kernel-name "$FILTER" --launch-skip 3 --launch-count 12 \
        -o gpurun_out/xchg_select_merge -f > gpurun_out/xchg_full.log 2>&1; tail -5 gpurun_out/xchg_full.log
      ncu -i gpurun_out/xchg_select_merge.ncu-rep --page raw --csv > gpurun_out/xchg_select_merge_raw.csv 2>/dev/null ;;
  esac
done; done
ls -la gpurun_out | grep xchg
