timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 --deselect tests/test_gpu_configs.py -p no:cacheprovider --timeout=240 > gpurun_out/r02_pytest_small.log 2>&1; tail -12 gpurun_out/r02_pytest_small.log
timeout 1500 python -m pytest tests/test_gpu_configs.py -q -s -p no:cacheprovider --timeout=240 > gpurun_out/r02_pytest_configs.log 2>&1; grep -E "parity\]|passed|failed|Error" gpurun_out/r02_pytest_configs.log | tail -30
timeout 200 python tools/gpu_loader_bench.py 1000000 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
