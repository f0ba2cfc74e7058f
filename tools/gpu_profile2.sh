#!/bin/bash
# ncu full capture of the tcgen05 conv kernels at 2048 rows (one step), after a plain run of the same command
CMD="python tools/profile_step.py --rows 2048 --steps 3"
$CMD > gpurun_out/plain_tc.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_conv_(fwd|dgrad)_tc' -s 12 -c 2 -o gpurun_out/prof_tc $CMD > gpurun_out/ncu_tc.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_tc.log
python bench.py --no-cpu > gpurun_out/bench_tc4.json 2>gpurun_out/bench_tc4.err; tail -2 gpurun_out/bench_tc4.err
python -c "
import json; b=json.loads(open('gpurun_out/bench_tc4.json').read()); print(b['value'], b['ms_per_step'], b['e2e']['value']); print(b['stage_ms'])"
