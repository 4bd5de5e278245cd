#!/bin/bash
# bench + ncu launch list + one full ncu capture of the scan kernel (1 GPU)
mkdir -p gpurun_out
R=${ROWS:-10000000}
timeout 900 python bench.py --steps 10 --warmup 3 --rows $R > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; tail -c 3000 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --rows $R --no-e2e --no-cpu > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_tc -s 3 -c 1 -f -o gpurun_out/prof_scan \
    python bench.py --steps 1 --warmup 3 --rows $R --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out/
