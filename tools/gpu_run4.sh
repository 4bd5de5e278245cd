echo "=== f32 probe"; timeout 300 python tools/gpu_f32_probe.py 4000000 > gpurun_out/r02c_f32probe.log 2>&1; tail -6 gpurun_out/r02c_f32probe.log
bash tools/gpu_perf.sh r02c 2>&1 | tail -32
echo "=== tests"
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --deselect tests/test_gpu_configs.py -p no:cacheprovider 2>&1 | tail -8
timeout 1500 python -m pytest tests/test_gpu_configs.py -q -s -p no:cacheprovider 2>&1 | grep -E "parity\]|passed|failed|Error" | tail -30
echo "=== sanitizer (memcheck only, short)"
timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_parity.py -q -x -k "golden or tie_probe" -p no:cacheprovider > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_memcheck.log | tail -4
