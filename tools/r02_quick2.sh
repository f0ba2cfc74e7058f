for f in 0 16 32 64 128 240; do
NMA_DIAG_I_KNOW_RESULTS_ARE_INVALID=1 NMA_DIAG=$f timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu --no-alt > gpurun_out/b_ar.json 2>gpurun_out/b_ar.err; python -c "
import json
d=json.loads(open('gpurun_out/b_ar.json').read().strip().splitlines()[-1])
print('diag $f', d['ms_per_step'], d['clocks']['sm_mhz'], {k:v for k,v in d['stage_ms'].items() if 'feat_fwd' in k})"
done
