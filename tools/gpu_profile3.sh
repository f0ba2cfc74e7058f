#!/bin/bash
CMD="python tools/profile_step.py --rows 2048 --steps 3"
$CMD > gpurun_out/plain_f.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_feat_fwd|k_feat_bwd|k_epi_bwd|k_conv_wgrad_tc' -s 12 -c 8 -o gpurun_out/prof_feat $CMD > gpurun_out/ncu_feat.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_feat.log
