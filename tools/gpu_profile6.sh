#!/bin/bash
# final state of round 1 (bf16 split, wide dgrad, per-row wgrad, fused dtb): ncu launch list of ONE eager step, 2048 rows
CMD="python tools/profile_step.py --rows 2048 --steps 3 --tc 7"
timeout 100 $CMD > gpurun_out/plain_final_bf.log 2>&1 &&
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bf16_final.csv $CMD > gpurun_out/ncu_launches_final_bf.log 2>&1
echo "launch list rc=$?"
