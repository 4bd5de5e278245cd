#!/bin/bash
mkdir -p gpurun_out
R=${ROWS:-10000000}
timeout 600 python tools/gpu_probe.py $R > gpurun_out/probe.log 2>&1; echo "probe exit $?"; cat gpurun_out/probe.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --rows $R --no-e2e --no-cpu > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_tc -s 3 -c 1 -f -o gpurun_out/prof_scan \
    python bench.py --steps 1 --warmup 3 --rows $R --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
