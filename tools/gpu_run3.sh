bash tools/gpu_perf.sh r02b 2>&1 | tail -40
echo "=== f32 probe"
timeout 300 python tools/gpu_f32_probe.py 4000000 0,1,2,3,7 2>&1 | tail -12
echo "=== ncu full f32"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_tc_kernel -s 3 -c 1 -o gpurun_out/r02_f32 python tools/gpu_f32_probe.py 2000000 0 > gpurun_out/r02_f32_ncu.log 2>&1
echo "ncu rc=$?"
echo "=== tests"
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --deselect tests/test_gpu_configs.py -p no:cacheprovider 2>&1 | tail -15
timeout 600 python -m pytest tests/test_gpu_configs.py -q -s -k "verbatim" -p no:cacheprovider 2>&1 | grep -E "parity\]|passed|failed|Error" | tail
