timeout 300 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_parity.py tests/test_gpu_regime.py -m gpu -q -x -k "weight_gradient or bf16 or wgrad or bench_scale" 2>&1 | tail -2
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu --no-alt > gpurun_out/b_ar.json 2>gpurun_out/b_ar.err; python -c "
import json
d=json.loads(open('gpurun_out/b_ar.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['clocks']); print(d['stage_ms'])"
ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_conv_wgrad_ts -s 2 -c 1 python tools/profile_step.py --rows 2048 --T 1000000 --tc 7 2>&1 | grep -E "dram__|gpu__time"
