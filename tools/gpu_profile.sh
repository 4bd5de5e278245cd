#!/bin/bash
# round profile artifacts: bench line, ncu launch list of the same command, full captures of the scan kernel
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; echo "bench exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?"
# steady state: 3 launches per step (dense prefix, main scan); skip the warm-up steps
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_tc -s 8 -c 2 -f -o gpurun_out/r01_scan_tc \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
# Q = 1000 (imagenet shape, 4 Q blocks): does the bank still stream from HBM once?
timeout 900 ncu --set full --clock-control none -k regex:scan_tc -s 6 -c 3 -f -o gpurun_out/r01_scan_tc_q1000 \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --classes 1000 --t2t-only > gpurun_out/ncu_q1000.log 2>&1; echo "ncu q1000 exit $?"
tail -c 700 gpurun_out/bench_r01.json
