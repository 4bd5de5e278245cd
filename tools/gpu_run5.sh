bash tools/gpu_perf.sh r02d 2>&1 | tail -34
echo "=== tests"
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --deselect tests/test_gpu_configs.py -p no:cacheprovider 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_configs.py -q -s -k "verbatim or config1_fp32" -p no:cacheprovider 2>&1 | grep -E "parity\]|passed|failed|Error" | tail -20
echo "=== loader"
timeout 300 python tools/gpu_loader_bench.py 2000000 2>&1 | tail -3
echo "=== sanitizer"
bash tools/gpu_sanitize.sh 2>&1 | tail -12
