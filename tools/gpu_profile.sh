#!/bin/bash
# round profile artifacts: bench line, ncu launch list of the same command, one full capture of the scan kernel
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; echo "bench exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_tc -s 6 -c 2 -f -o gpurun_out/r01_scan_tc \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
timeout 900 ncu --set full --clock-control none -k regex:"select_kernel|partition_kernel|t2i_rescore|t2i_walk|final_tau" -s 15 -c 5 -f -o gpurun_out/r01_select \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_sel.log 2>&1; echo "ncu select exit $?"
tail -c 600 gpurun_out/bench_r01.json
