#!/bin/bash
set -u
mkdir -p gpurun_out
for cfg in "8 weak" "8 strong" "4 strong"; do
  set -- $cfg
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $1 --steps 20 --warmup 5 --scaling $2 --no-alt \
    > gpurun_out/r02_bench_$1gpu_$2.json 2> gpurun_out/r02_bench_$1gpu_$2.err
  echo "$1 $2 rc=$?"; tail -c 300 gpurun_out/r02_bench_$1gpu_$2.json | head -c 300; echo
done
