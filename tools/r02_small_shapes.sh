#!/bin/bash
# ncu launch lists (cold-cache, serialised) of the script-shape workloads; run under gpurun
for c in lv_fix_theta lv_batch fhn sv; do
  NMA_FACADE_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file gpurun_out/r02_launches_$c.csv python bench.py --config $c --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_$c.log 2>&1
  python tools/launch_table.py gpurun_out/r02_launches_$c.csv > gpurun_out/r02_launch_table_$c.txt 2>&1
done
tail -n 45 gpurun_out/r02_launch_table_*.txt
