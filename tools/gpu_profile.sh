#!/bin/bash
# Run on the GPU box (under gpurun): default bench, then the ncu launch list of ONE step and a full capture of the
# tcgen05 conv kernels of that step (2048 rows).  Every ncu pass follows a plain run of the same command.
CMD="python tools/profile_step.py --rows 2048 --steps 4"
timeout 280 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
timeout 100 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 250 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
timeout 100 $CMD > gpurun_out/plain2.log 2>&1 &&
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'k_conv_(fwd|dgrad|wgrad)_tc' -c 9 -o gpurun_out/prof_conv_tc $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
tail -2 gpurun_out/ncu_full.log
cat gpurun_out/bench_default.json
