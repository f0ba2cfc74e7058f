#!/bin/bash
# Round-2 evidence on ONE B200: tests, launch lists, ncu capture of the roofline kernel, bench lines.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rs > gpurun_out/r02_gpu_tests.log 2>&1; echo "suite rc=$?" >> gpurun_out/r02_gpu_tests.log
# launch list of exactly one step (cudaProfilerStart/Stop around it): the reference's p = 50 shape and 2048 rows
for rows in 50 2048; do
  T=1000000; [ $rows = 50 ] && T=5000
  timeout 100 python tools/profile_step.py --rows $rows --T $T --tc 7 > gpurun_out/r02_step_${rows}_plain.log 2>&1 && \
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_${rows}.csv \
    python tools/profile_step.py --rows $rows --T $T --tc 7 > gpurun_out/r02_ncu_launches_${rows}.log 2>&1
done
# full capture of the roofline kernel (weight gradient of flow 0 = third launch of the backward pass)
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_conv_wgrad_ts -s 2 -c 1 -o gpurun_out/r02c_full_wgrad0 \
  python tools/profile_step.py --rows 2048 --T 1000000 --tc 7 > gpurun_out/r02c_ncu_full_wgrad0.log 2>&1
# script-shape launch lists
bash tools/r02_small_shapes.sh > /dev/null 2>&1
# bench lines
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_ar_1e8.json 2> gpurun_out/r02_bench_ar_1e8.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_ar_1e8_reference.json 2>> gpurun_out/r02_bench_ar_1e8.err
for c in ar_default fhn sv lv_fix_theta lv_batch; do
  timeout 300 python bench.py --config $c > gpurun_out/r02_bench_$c.json 2> gpurun_out/r02_bench_$c.err
  timeout 300 python bench.py --config $c --impl reference --steps 10 > gpurun_out/r02_bench_${c}_reference.json 2>> gpurun_out/r02_bench_$c.err
done
timeout 300 python tools/bench_streaming.py > gpurun_out/r02_streaming.log 2>&1
( time NMA_MAX_STEPS=2000 timeout 200 python main.py hyperparameters.txt ) > gpurun_out/r02_main_p50.log 2>&1
tail -4 gpurun_out/r02_gpu_tests.log; tail -c 400 gpurun_out/r02_bench_ar_1e8.json
