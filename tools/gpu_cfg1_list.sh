#!/bin/bash
# per-kernel times of BASELINE config 1 (1 M x 512 fp32 rows, T2T-500): ncu launch list of the probe
KREG='regex:scan_tc|scan_simt|select_kernel|partition|final_tau|reset_kernel|bootstrap|rescore|walk_kernel|merge|splice|remap'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 40 --csv --log-file gpurun_out/r02f_cfg1_launches.csv python tools/gpu_cfg1_probe.py 1000000 3 > gpurun_out/r02f_cfg1_ncu.log 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/r02f_cfg1_launches.csv")))
i0=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
h=rows[i0]
for r in rows[i0+1:][9:27]:
    d=dict(zip(h,r))
    print(d["ID"], d["Kernel Name"].split("::")[-1][:50], d["Grid Size"], d["Metric Value"])
PY
python tools/gpu_cfg1_probe.py 1000000 5 2>&1 | grep " 4 wall"
