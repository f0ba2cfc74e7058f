#!/bin/bash
# bf16-split conv path: ncu launch list of ONE eager step and a full capture of the conv kernels (2048 rows).
# Every ncu pass follows a plain run of the same command that exited 0.
CMD="python tools/profile_step.py --rows 2048 --steps 3 --tc 7"
timeout 100 $CMD > gpurun_out/plain_bf.log 2>&1 &&
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bf16.csv $CMD > gpurun_out/ncu_launches_bf.log 2>&1
echo "launch list rc=$?"
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'k_conv_' -c 9 -f -o gpurun_out/prof_bf16 $CMD > gpurun_out/ncu_full_bf.log 2>&1
echo "full rc=$?"
tail -2 gpurun_out/ncu_full_bf.log
