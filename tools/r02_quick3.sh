for f in 8 16 32 64; do
echo "== flush $f"
NMA_WS_FLUSH=$f timeout 300 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_regime.py -m gpu -q -s -k "weight_gradient or bench_scale" 2>&1 | grep -E "wgrad:|worst|passed|failed|rows" | head -12
done
