#!/bin/bash
timeout 300 ncu --metrics dram__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_scan_lookback|k_time_till|k_adamax|k_sumsq' -c 12 --csv --log-file gpurun_out/r02_ncu_streaming.csv python tools/bench_streaming.py > gpurun_out/r02_ncu_streaming.log 2>&1
tail -30 gpurun_out/r02_ncu_streaming.csv | cut -c1-220
