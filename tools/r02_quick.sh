timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for i in 1 2; do timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu --no-alt > gpurun_out/b_ar.json 2>gpurun_out/b_ar.err; python -c "
import json
d=json.loads(open('gpurun_out/b_ar.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['clocks']); print(d['stage_ms'])"; done
