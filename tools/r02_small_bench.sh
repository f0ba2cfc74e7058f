#!/bin/bash
for c in ar_default fhn sv lv_fix_theta lv_batch; do
  timeout 300 python bench.py --config $c > gpurun_out/r02_bench_$c.json 2> gpurun_out/r02_bench_$c.err
  python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_$c.json').read().strip().splitlines()[-1])
print('$c', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
done
