#!/bin/bash
# round-2 final evidence: bench line, ncu launch list of the same command, full capture of the selecting scan, phase trace
mkdir -p gpurun_out
bash tools/gpu_perf.sh r02f
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_tc -s 7 -c 1 -f -o gpurun_out/r02f_scan_tc \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/r02f_ncu_full.log 2>&1; echo "ncu full exit $?"
python tools/gpu_trace_probe.py 2>&1 | grep -E "trace|thr=" | sed 's/us since first entry (min\/max over CTAs): //' > gpurun_out/r02f_trace.log
python tools/gpu_ab_step.py swat_b200/libswat_b200.so 2>&1 | tail -3 > gpurun_out/r02f_ab.log
python tools/gpu_ab_step.py tools/ab/libswat_b200_old.so 2>&1 | tail -3 >> gpurun_out/r02f_ab.log
python tools/gpu_ab_step.py swat_b200/libswat_b200.so 2>&1 | tail -3 >> gpurun_out/r02f_ab.log
python tools/gpu_ab_step.py tools/ab/libswat_b200_old.so 2>&1 | tail -3 >> gpurun_out/r02f_ab.log
cat gpurun_out/r02f_ab.log
