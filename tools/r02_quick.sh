timeout 300 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_parity.py -m gpu -q -x -k "weight_gradient or bf16 or wgrad" 2>&1 | tail -3
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu --no-alt > gpurun_out/b_ar.json 2>gpurun_out/b_ar.err; python -c "
import json
d=json.loads(open('gpurun_out/b_ar.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['clocks']); print(d['stage_ms'])"
