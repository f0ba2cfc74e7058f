#!/bin/bash
# ncu --set full captures (2048 rows, flow 0) of the feature / head kernels and the weight gradient
set -u
timeout 100 python tools/profile_step.py --rows 2048 --T 1000000 --tc 7 > gpurun_out/r02_step_2048_plain.log 2>&1 || exit 1
for k in k_feat_fwd_tc k_feat_bwd_tc k_epi_bwd_tc k_conv_wgrad_ts; do
  skip=0; [ $k = k_feat_bwd_tc ] && skip=2; [ $k = k_epi_bwd_tc ] && skip=2; [ $k = k_conv_wgrad_ts ] && skip=2
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o gpurun_out/r02b_full_$k \
    python tools/profile_step.py --rows 2048 --T 1000000 --tc 7 > gpurun_out/r02b_ncu_$k.log 2>&1
done
ls -la gpurun_out/r02b_full_*
