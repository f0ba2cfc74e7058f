#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r02_t3.log 2>&1; echo "suite rc=$?" >> gpurun_out/r02_t3.log
timeout 200 python tools/bench_streaming.py 2>&1 | head -2 > gpurun_out/r02_streaming2.log
timeout 100 python tools/profile_step.py --rows 50 --T 5000 --tc 7 > gpurun_out/r02_p50_plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_p50.csv \
  python tools/profile_step.py --rows 50 --T 5000 --tc 7 > gpurun_out/r02_ncu_p50.log 2>&1
timeout 100 python tools/profile_step.py --rows 2048 --T 1000000 --tc 7 > gpurun_out/r02_2048_plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_2048.csv \
  python tools/profile_step.py --rows 2048 --T 1000000 --tc 7 > gpurun_out/r02_ncu_2048.log 2>&1
timeout 200 python bench.py --config ar_default --no-cpu > gpurun_out/r02_bench_ar_default_bf16.json 2>gpurun_out/r02_bench_ar_default_bf16.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_smoke.log
tail -5 gpurun_out/r02_t3.log; cat gpurun_out/r02_streaming2.log; tail -3 gpurun_out/r02_smoke.log
