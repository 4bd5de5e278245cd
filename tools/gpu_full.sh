#!/bin/bash
# GPU box: what the driver runs at round end, in its order -- the whole -m gpu suite, smoke(), the reference arm, the bench line
mkdir -p gpurun_out
TAG=${1:-r02h}
t0=$(date +%s)
timeout 2400 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - t0 )) s)"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
t0=$(date +%s)
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2; echo "smoke ($(( $(date +%s) - t0 )) s)"
t0=$(date +%s)
timeout 900 python bench.py --impl reference > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$? ($(( $(date +%s) - t0 )) s)"
t0=$(date +%s)
timeout 1200 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$? ($(( $(date +%s) - t0 )) s)"
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().split("\n")[-1])
print("ms/step", d["ms_per_step"], "value", d["value"], "roof", d["roofline"]["frac"], "step_frac", d["stats"]["step_frac_of_roofline"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "clocks", d["clocks"])
c=d["configs"]
print("cold", c["cold_call_ms"])
for e in c.get("cfg3_sweep9",[])+c["cfg5_qsweep"]+[c["cfg5_strong_scaling"]]+c.get("cfg1_fp32",[]):
    print(e["workload"][:70], "| ms", round(e["ms_per_step"],3), "kern", e["scan_kernel_ms"] and round(e["scan_kernel_ms"],3), "step_frac", round(e["roofline"]["step_frac"],3), "kern_frac", e["roofline"]["kernel_frac"] and round(e["roofline"]["kernel_frac"],3), e.get("escalations_per_step"))
r=json.loads(open("gpurun_out/${TAG}_bench_ref.json").read().strip().split("\n")[-1])
print("reference arm:", r["value"], r["unit"], r["cpu_baseline"]["sample"])
PY
