#!/bin/bash
# Run on the GPU box (under gpurun): default bench, then the ncu launch list and one full capture of the conv kernels.
set -x
CMD="python bench.py --rows 2048 --steps 2 --warmup 3 --no-cpu"
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 72 -c 120 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_conv -s 9 -c 3 -o gpurun_out/prof_conv $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
tail -3 gpurun_out/ncu_full.log
cat gpurun_out/bench_default.json
