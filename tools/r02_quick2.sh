for f in ${FLUSHES:-32}; do
NMA_WS_FLUSH=$f timeout 400 python bench.py --steps 4 --warmup 3 --no-cpu --no-alt > gpurun_out/b_ar.json 2>gpurun_out/b_ar.err; python -c "
import json
d=json.loads(open('gpurun_out/b_ar.json').read().strip().splitlines()[-1])
print('flush $f', d['ms_per_step'], d['clocks']['sm_mhz'], {k:v for k,v in d['stage_ms'].items() if 'wgrad' in k})"
done
