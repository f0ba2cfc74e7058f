timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for c in sv fhn lv_fix_theta lv_batch ar_default; do timeout 200 python bench.py --config $c --steps 50 --warmup 5 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['workload'][:30], d['value'], d['ms_per_step'], d['e2e']['value'])"; done
