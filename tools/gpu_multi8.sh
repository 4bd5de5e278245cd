#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
CHECK_ROWS=16777216 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/r02_dist_check_$N.log 2>&1
echo "dist check exit $?"; grep -E "world=|rror" gpurun_out/r02_dist_check_$N.log | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
echo "bench exit $?"; tail -n 3 gpurun_out/r02_bench_n$N.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_n$N.json").read().strip().split("\n")[-1])
print("N", d["n_gpus"], "ms/step", d["ms_per_step"], "value", d["value"], "roof", d["roofline"]["frac"], "e2e", d["e2e"] and d["e2e"]["value"], "launches", d["gpu_launches"], d["clocks"])
c=d["configs"]
print("cold", c["cold_call_ms"])
for e in c["cfg5_qsweep"]+[c["cfg5_strong_scaling"]]:
    print(e["workload"][:80], "| ms", round(e["ms_per_step"],3), "kern", e["scan_kernel_ms"] and round(e["scan_kernel_ms"],3), "step_frac", round(e["roofline"]["step_frac"],3), "kern_frac", e["roofline"]["kernel_frac"] and round(e["roofline"]["kernel_frac"],3))
PY
