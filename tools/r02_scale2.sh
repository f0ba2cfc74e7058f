#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -s > gpurun_out/r02_multi_gpu_nccl_parity.log 2>&1; tail -3 gpurun_out/r02_multi_gpu_nccl_parity.log
for sc in weak strong; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 --scaling $sc --no-alt \
    > gpurun_out/r02_bench_2gpu_$sc.json 2> gpurun_out/r02_bench_2gpu_$sc.err
  echo "2 $sc rc=$?"; tail -c 300 gpurun_out/r02_bench_2gpu_$sc.json; echo
done
