#!/bin/bash
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/final_bench.json
timeout 600 python bench.py --impl reference > gpurun_out/final_bench_ref.json 2>> gpurun_out/final_bench.err; echo "ref rc=$?"; tail -c 400 gpurun_out/final_bench_ref.json
