#!/bin/bash
# round 2, GPU call A: new parity tests, per-config bench lines, main.py at p=50, launch list of one step
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_regime.py tests/test_gpu_train.py tests/test_gpu_train_step.py -q -s > gpurun_out/r02_regime.log 2>&1
echo "regime rc=$?" >> gpurun_out/r02_regime.log
for c in ar_default fhn sv lv_fix_theta; do
  timeout 300 python bench.py --config $c > gpurun_out/r02_bench_$c.json 2> gpurun_out/r02_bench_$c.err; echo "$c rc=$?"
  timeout 300 python bench.py --config $c --impl reference --steps 10 > gpurun_out/r02_bench_${c}_reference.json 2>> gpurun_out/r02_bench_$c.err
done
( time NMA_MAX_STEPS=2000 timeout 200 python main.py hyperparameters.txt ) > gpurun_out/r02_main_p50.log 2>&1
timeout 300 python tools/bench_streaming.py > gpurun_out/r02_streaming.log 2>&1
timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu --no-alt --no-graph --rows 2048 > gpurun_out/r02_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_step.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-alt --no-graph --rows 2048 > gpurun_out/r02_ncu_launches.log 2>&1
tail -40 gpurun_out/r02_regime.log
