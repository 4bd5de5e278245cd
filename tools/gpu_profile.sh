#!/bin/bash
# round-2 profile artifacts: full captures of the scan kernel (bf16 bench workload, fp32 banks), DRAM traffic with several Q blocks
mkdir -p gpurun_out
# bench workload: launches per step = dense prefix + main scan; skip the warm-up steps, capture one main scan
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_tc -s 7 -c 1 -f -o gpurun_out/r02_scan_tc \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/r02_ncu_full.log 2>&1; echo "ncu full exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_tc_kernel -s 3 -c 1 -f -o gpurun_out/r02_scan_f32_fixed \
    python tools/gpu_f32_probe.py 4000000 > gpurun_out/r02_ncu_f32.log 2>&1; echo "ncu f32 exit $?"
timeout 600 ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:scan_tc --csv --log-file gpurun_out/r02_l2share.csv python tools/gpu_l2share.py 10000000 200,400,1000 > gpurun_out/r02_l2share.log 2>&1; echo "l2share exit $?"
