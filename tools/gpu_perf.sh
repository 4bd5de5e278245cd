#!/bin/bash
# GPU box: bench line (with the `configs` extras), ncu launch list of the same command, timing probe
mkdir -p gpurun_out
TAG=${1:-r02}
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/${TAG}_clocks.csv &
SMI=$!
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"
kill $SMI
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 300 python tools/gpu_probe.py > gpurun_out/${TAG}_probe.log 2>&1
echo "probe rc=$?"
tail -c 6000 gpurun_out/${TAG}_bench.json
tail -5 gpurun_out/${TAG}_bench.err
tail -12 gpurun_out/${TAG}_probe.log
