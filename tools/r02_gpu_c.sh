#!/bin/bash
# 2-GPU call: library-owned NCCL all-reduce numerics, weak + strong bench lines, clean exit
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -s > gpurun_out/r02_multi.log 2>&1; echo "multi rc=$?" >> gpurun_out/r02_multi.log
for mode in weak strong; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --scaling $mode > gpurun_out/r02_bench_2gpu_$mode.json 2> gpurun_out/r02_bench_2gpu_$mode.err
  echo "$mode rc=$?"
done
tail -15 gpurun_out/r02_multi.log; tail -c 600 gpurun_out/r02_bench_2gpu_weak.json; tail -3 gpurun_out/r02_bench_2gpu_weak.err
