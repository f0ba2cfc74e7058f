#!/bin/bash
# First GPU call of the next session: run what was written after round 1's GPU budget was spent, then re-confirm the
# verified state.  Everything lands in gpurun_out/first_*.log.   gpurun --timeout 900 -- 'bash tools/gpu_first_steps.sh'
set -u
mkdir -p gpurun_out
# 1. code that has never run on hardware (skipped without the variable, refused by the library without it)
NMA_UNVERIFIED=1 timeout 300 python -m pytest tests/test_gpu_unverified.py -q -s 2>&1 | tail -30 > gpurun_out/first_unverified.log
# 2. the odd-kernel_len tap pairs (zero-kernel pairing of the last tap): step parity at kernel_len = 7 with the pairs forced on
NMA_TAP_PAIRS=1 timeout 120 python - > gpurun_out/first_odd_pairs.log 2>&1 <<'PY'
import sys; sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from viforssms_b200.config import ar_config
from test_gpu_parity import _check_step
for shape in (dict(p=5, K=7, B=5, F=2, H=1, feat_window=2), dict(p=9, K=11, B=6, F=3, H=1, feat_window=3)):
    print(shape, "worst", _check_step(ar_config(T=400, **shape), 400, seed=3, tc=7))
PY
# 3. the verified suite and the bench, both operand formats, and the device theta posterior in the stepper
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/first_suite.log
timeout 150 python bench.py > gpurun_out/first_bench_bf16.json 2> gpurun_out/first_bench_bf16.err
timeout 150 python bench.py --no-cpu --conv-split tf32 > gpurun_out/first_bench_tf32.json 2> gpurun_out/first_bench_tf32.err
timeout 150 python bench.py --no-cpu --device-theta > gpurun_out/first_bench_device_theta.json 2> gpurun_out/first_bench_device_theta.err
# the reference's own shape (p = 50 rows), where the theta flow's ~400 launches matter most
timeout 150 python bench.py --no-cpu --rows 50 --steps 200 > gpurun_out/first_bench_p50.json 2> gpurun_out/first_bench_p50.err
timeout 150 python bench.py --no-cpu --rows 50 --steps 200 --device-theta > gpurun_out/first_bench_p50_device_theta.json 2> gpurun_out/first_bench_p50_device_theta.err
# 4. the drop-in script with and without the captured iteration (p = 50: launch-bound)
( time NMA_MAX_STEPS=600 timeout 200 python main.py hyperparameters.txt ) > gpurun_out/first_main_eager.log 2>&1
( time NMA_FACADE_GRAPH=1 NMA_MAX_STEPS=600 timeout 200 python main.py hyperparameters.txt ) > gpurun_out/first_main_graph.log 2>&1
tail -3 gpurun_out/first_unverified.log gpurun_out/first_odd_pairs.log gpurun_out/first_suite.log
