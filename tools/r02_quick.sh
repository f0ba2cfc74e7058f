timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
timeout 300 python -m pytest tests/test_gpu_regime.py tests/test_gpu_bf16.py -m gpu -q -s -k "bench_scale or ar_default or at_1e8" 2>&1 | grep -E "rows,|worst|step parity" | head -12
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu --no-alt > gpurun_out/b_ar.json 2>gpurun_out/b_ar.err; python -c "
import json
d=json.loads(open('gpurun_out/b_ar.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['clocks']); print(d['stage_ms'])"
