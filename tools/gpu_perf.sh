#!/bin/bash
# GPU box: bench line (with the `configs` extras), ncu launch list of the same command, timing probe
mkdir -p gpurun_out
TAG=${1:-r02}
KREG='regex:scan_tc|scan_simt|select_kernel|partition|final_tau|reset_kernel|bootstrap|rescore|walk_kernel|merge'
if [ -n "$SMI_LOOP" ]; then
  nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/${TAG}_clocks.csv &
  SMI=$!
fi
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"
[ -n "$SMI" ] && kill $SMI
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 300 python tools/gpu_probe.py > gpurun_out/${TAG}_probe.log 2>&1
echo "probe rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().split("\n")[-1])
print("ms/step", d["ms_per_step"], "value", d["value"], "roof", d["roofline"]["frac"], "step_frac", d["stats"]["step_frac_of_roofline"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
c=d["configs"]
print("cold", c["cold_call_ms"], c.get("cold_call_breakdown_ms")); print("e2e", d["e2e"].get("ms_per_step"), d["e2e"].get("last_step_device_ms"))
for e in c.get("cfg3_sweep9",[])+c["cfg5_qsweep"]+[c["cfg5_strong_scaling"]]+c.get("cfg1_fp32",[]):
    print(e["workload"][:70], "| ms", round(e["ms_per_step"],3), "kern", e["scan_kernel_ms"] and round(e["scan_kernel_ms"],3), "step_frac", round(e["roofline"]["step_frac"],3), "kern_frac", e["roofline"]["kernel_frac"] and round(e["roofline"]["kernel_frac"],3), e.get("escalations_per_step"))
PY
tail -14 gpurun_out/${TAG}_probe.log
