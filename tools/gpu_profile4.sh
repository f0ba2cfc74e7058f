#!/bin/bash
# Run on the GPU box (under gpurun): default bench (with the CPU baseline leg) and the reference arm, then the ncu
# launch list of ONE eager step and a full capture of this library's main kernels in that step (2048 rows).
# Every ncu pass follows a plain run of the same command that exited 0.
CMD="python tools/profile_step.py --rows 2048 --steps 3"
timeout 280 python bench.py > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_r01_final.err; echo "bench rc=$?"
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r01_reference.json 2> gpurun_out/bench_r01_reference.err; echo "reference rc=$?"
timeout 100 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 250 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
timeout 100 $CMD > gpurun_out/plain2.log 2>&1 &&
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'k_conv_.*tc|k_feat_.*_tc|k_epi_bwd_tc' -c 20 -f -o gpurun_out/prof_final $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
tail -2 gpurun_out/ncu_full.log
cat gpurun_out/bench_r01_final.json
