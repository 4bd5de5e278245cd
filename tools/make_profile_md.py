"""Turn the raw ncu outputs of tools/gpu_profile.sh (gpurun_out/) into the tracked summaries under profiles/.

    python tools/make_profile_md.py [round_tag=r01]

Needs `ncu` (reads .ncu-rep files offline; no GPU).  Writes
  profiles/<tag>_launches.csv / <tag>_launches_summary.md   launch list of the bench command
  profiles/<tag>_scan_tc_ncu.md                             --set full metrics of the scan launches (Q=200 and Q=1000)
  profiles/scan_traffic.json                                DRAM bytes of the selecting scan (bench.py's roofline.traffic)
"""
import csv
import io
import json
import os
import subprocess
import sys
from collections import OrderedDict, defaultdict

TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"
OUT = "profiles"
SRC = "gpurun_out"
os.makedirs(OUT, exist_ok=True)


def short(name):
    name = name.replace("void swat::<unnamed>::", "").replace("swat::<unnamed>::", "").replace("void ", "")
    return name.split("(")[0]


# ------------------------------------------------------------------------------------- launch list
raw = [l for l in open(f"{SRC}/{TAG}_launches.csv", errors="replace") if l.startswith('"')]
open(f"{OUT}/{TAG}_launches.csv", "w").writelines(raw)
rows = list(csv.DictReader(io.StringIO("".join(raw))))
ours, other_n, other_ns = [], 0, 0.0
for r in rows:
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    ns = float(r["Metric Value"])
    if "swat::" in r["Kernel Name"]:
        ours.append((short(r["Kernel Name"]), ns))
    else:
        other_n += 1
        other_ns += ns
agg = defaultdict(lambda: [0, 0.0])
for k, ns in ours:
    agg[k][0] += 1
    agg[k][1] += ns
tot = sum(v[1] for v in agg.values())
md = [f"# Round {TAG[1:]} — ncu launch list of `python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu` (1 x B200)", "",
      "`ncu --metrics gpu__time_duration.sum --clock-control none` — per-launch times are cold-cache and serialised; compare SHARES.",
      f"Raw list: `profiles/{TAG}_launches.csv`.  Launches of this library's kernels over 5 pipeline steps (3 warm-up + 2 timed;",
      f"the {other_n} other launches, {other_ns / 1e6:.1f} ms, are torch kernels generating the synthetic banks before the timed region).",
      "`scan_tc_kernel<2, 0, 0, 1>` is the dense prefix pass of the threshold bootstrap, `<2, 0, 0, 0>` the selecting scan.", "",
      "| kernel | launches | total ms | share of our kernels | mean us |", "|---|---:|---:|---:|---:|"]
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    md.append(f"| `{k}` | {n} | {ns / 1e6:.3f} | {100 * ns / tot:.1f}% | {ns / n / 1e3:.1f} |")
# last step in launch order
names = [k for k, _ in ours]
last_reset = max(i for i, k in enumerate(names) if k.startswith("reset_kernel"))
md += ["", "Last pipeline step, in launch order (steady state):", "", "| # | kernel | us |", "|---|---|---:|"]
for i, (k, ns) in enumerate(ours[last_reset:]):
    md.append(f"| {i} | `{k}` | {ns / 1e3:.1f} |")
open(f"{OUT}/{TAG}_launches_summary.md", "w").write("\n".join(md) + "\n")

# ------------------------------------------------------------------------------------- full captures
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__inst_executed_op_global_red.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "launch__shared_mem_per_block_dynamic"]


def capture(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    H, U = rows[0], rows[1]
    res = []
    for V in rows[2:]:
        d = OrderedDict()
        d["kernel"] = "scan_tc_kernel" + V[H.index("Kernel Name")].split("scan_tc_kernel")[-1].split("(")[0]
        for w in sorted(WANT):
            if w in H:
                d[w] = (V[H.index(w)], U[H.index(w)])
        res.append(d)
    return res


def stalls(rep, top=12):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return []
    H = rows[1]
    si, ci, ei = H.index("# Samples"), H.index("Source"), H.index("Instructions Executed")
    st = [i for i, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]
    seen, data, tot = set(), [], 0.0
    for r in rows[2:]:
        try:
            v = float(r[si])
        except Exception:
            continue
        key = (r[ci], r[ei], r[si])
        if key in seen:          # the source page lists every kernel of the report; identical lines repeat
            continue
        seen.add(key)
        tot += v
        best = sorted(((float(r[i] or 0), H[i]) for i in st), reverse=True)[0]
        data.append((v, r[ci].strip(), r[ei], best))
    return [f"| {100 * v / tot:.1f}% | {e} | `{s[:70]}` | {b[1]} |" for v, s, e, b in sorted(data, reverse=True)[:top]]


md = [f"# Round {TAG[1:]} — `ncu --set full --clock-control none --import-source on -k regex:scan_tc` (1 x B200)", "",
      "Raw reports are kept out of git (10+ MB each); regenerate with `tools/gpu_profile.sh` + `tools/make_profile_md.py`.",
      "Per-launch numbers under ncu are cold-cache, serialised and (tensor-bound shapes) power-capped: see the SM clock line."]
traffic = None
for title, rep in (("BASELINE workload: 10 M x 512 bf16, C = Q = 200, T2T500+T2I0.25 (`bench.py --steps 1 --warmup 3`): dense prefix pass "
                    "of the threshold bootstrap, then the selecting scan", f"{SRC}/{TAG}_scan_tc.ncu-rep"),
                   ("imagenet shape: 10 M x 512 bf16, C = Q = 1000, T2T top-500 (`--classes 1000 --t2t-only`; 4 query blocks, unit plan: 2 "
                    "launches per scan)", f"{SRC}/{TAG}_scan_tc_q1000.ncu-rep")):
    if not os.path.exists(rep):
        continue
    md += ["", f"## {title}"]
    for i, d in enumerate(capture(rep)):
        md += ["", f"### launch {i}: `scan_tc_kernel{d['kernel'].split('scan_tc_kernel')[-1]}`", "", "| metric | value | unit |", "|---|---:|---|"]
        for k, v in d.items():
            if k != "kernel":
                md.append(f"| {k} | {v[0]} | {v[1]} |")
        if "q1000" not in rep and d["kernel"].rstrip().endswith("0>"):
            rd, wr = d["dram__bytes_read.sum"], d["dram__bytes_write.sum"]
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            traffic = {"kernel": "scan_tc_kernel<2,0,0,0> (selecting scan of the bench workload)",
                       "dram_bytes_read": float(rd[0]) * scale[rd[1]], "dram_bytes_write": float(wr[0]) * scale[wr[1]],
                       "algorithmic_bytes_per_launch": 10_000_000 * 1024 - 32768 * 1024,
                       "source": f"ncu --set full, profiles/{TAG}_scan_tc_ncu.md"}
            traffic["dram_bytes_per_launch"] = traffic["dram_bytes_read"] + traffic["dram_bytes_write"]
    s = stalls(rep)
    if s:
        md += ["", "Top sampled SASS lines (share of samples, executions, instruction, dominant stall):", "", "| samples | exec | SASS | stall |",
               "|---:|---:|---|---|"] + s
open(f"{OUT}/{TAG}_scan_tc_ncu.md", "w").write("\n".join(md) + "\n")
if traffic:
    json.dump(traffic, open(f"{OUT}/scan_traffic.json", "w"), indent=1)
print("wrote", f"{OUT}/{TAG}_launches_summary.md", f"{OUT}/{TAG}_scan_tc_ncu.md", "traffic" if traffic else "")
