#!/bin/bash
# GPU box: small -m gpu suite, config-1 step breakdown, main bench line and the ncu launch list of the last steps
mkdir -p gpurun_out
TAG=${1:-r02g}
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 --deselect tests/test_gpu_configs.py -p no:cacheprovider --timeout=240 > gpurun_out/${TAG}_pytest_small.log 2>&1
echo "small rc=$?"; tail -5 gpurun_out/${TAG}_pytest_small.log
python tools/gpu_cfg1_probe.py 1000000 4 2>&1 | tee gpurun_out/${TAG}_cfg1.log | grep -v " 0 wall"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-extras > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().split("\n")[-1])
print("ms/step", d["ms_per_step"], "value", d["value"], "roof", d["roofline"]["frac"], "kernel_ms", d["roofline"]["kernel_ms"], "e2e", d["e2e"]["value"])
PY
KREG='regex:scan_tc|scan_simt|select_kernel|partition|final_tau|reset_kernel|bootstrap|rescore|walk_kernel|merge|splice|remap'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/${TAG}_ncu.log 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/${TAG}_launches.csv")))
i0=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
h=rows[i0]
for r in rows[i0+1:][-9:]:
    d=dict(zip(h,r))
    print(d["ID"], d["Kernel Name"][:60], d["Grid Size"], d["Metric Value"])
PY
