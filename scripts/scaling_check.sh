#!/bin/bash
# Runs on an 8-GPU box through `gpurun --gpus 8`: strong-scaling bench at N=2/4/8 with the fused peer-memory exchange,
# N=8 again with the NCCL all-gather exchange, then the multi-GPU parity tests. Logs -> gpurun_out/.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
launch() { # n exchange port
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$1" --master-addr 127.0.0.1 --master-port "$3" \
    bench.py --gpus "$1" --steps 20 --warmup 5 --exchange "$2" > "gpurun_out/bench_n$1_$2.json" 2> "gpurun_out/bench_n$1_$2.err"
  echo "exit=$? n=$1 $2"; tail -c 600 "gpurun_out/bench_n$1_$2.json"; echo
}
launch 2 p2p 29512
launch 4 p2p 29514
launch 8 p2p 29518
launch 8 nccl 29528
timeout 900 python -m pytest tests/test_multigpu_gpu.py -v -m gpu --timeout=600 -p no:cacheprovider > gpurun_out/test_multigpu.log 2>&1
echo "exit=$? test_multigpu"; tail -8 gpurun_out/test_multigpu.log
