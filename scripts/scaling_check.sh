#!/bin/bash
# Rehearsal of the driver's SCALE procedure on one 8-GPU box: both arms at N = 1, 2, 4, 8 back to back (default flags),
# then the multi-GPU parity tests. Logs -> gpurun_out/.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { # n port
  local n=$1 port=$2
  if [ "$n" = 1 ]; then
    timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/scale_ref_n1.json 2> gpurun_out/scale_ref_n1.err
    timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port "$port" \
      bench.py --impl reference --gpus "$n" --steps 20 --warmup 5 > "gpurun_out/scale_ref_n$n.json" 2> "gpurun_out/scale_ref_n$n.err"
    timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $((port + 2)) \
      bench.py --gpus "$n" --steps 20 --warmup 5 > "gpurun_out/scale_n$n.json" 2> "gpurun_out/scale_n$n.err"
  fi
  echo "exit=$? n=$n"; tail -c 300 "gpurun_out/scale_n$n.json"; echo
}
run 1 29500
run 2 29510
run 4 29520
run 8 29530
timeout 900 python -m pytest tests/test_multigpu_gpu.py -v -m gpu --timeout=600 -p no:cacheprovider > gpurun_out/test_multigpu.log 2>&1
echo "exit=$? test_multigpu"; tail -4 gpurun_out/test_multigpu.log
