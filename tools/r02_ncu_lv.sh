#!/bin/bash
for k in k_conv_fwd k_epi_bwd; do
NMA_FACADE_GRAPH=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:^$k\$ -s 6 -c 1 -o gpurun_out/r02_lv_$k python bench.py --config lv_fix_theta --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_lv_$k.log 2>&1
done
ls -la gpurun_out/r02_lv_*
