"""Summarise an .ncu-rep: key raw metrics + top stalled SASS lines (needs ncu, no GPU)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H, U = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_lsu.sum",
        "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum", "smsp__inst_executed_op_global_red.sum", "launch__grid_size", "launch__block_size"]
for V in rows[2:]:
    print("== kernel:", V[H.index("Kernel Name")][:70])
    for h, u, v in zip(H, U, V):
        if h in want:
            print(f"  {h:70s} {v} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
H = rows[1]
si, ci, ei = H.index("# Samples"), H.index("Source"), H.index("Instructions Executed")
stall = [i for i, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]
data, tot = [], 0.0
for r in rows[2:]:
    try:
        v = float(r[si])
    except Exception:
        continue
    tot += v
    st = sorted([(float(r[i]), H[i]) for i in stall], reverse=True)[:2]
    data.append((v, r[ci].strip(), r[ei], st))
print("total samples", tot)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for v, s, e, st in sorted(data, reverse=True)[:n]:
    print(f"{100*v/tot:5.1f}% exec={e:>10} {s[:64]:64s} {st[0][1]}={st[0][0]:.0f} {st[1][1]}={st[1][0]:.0f}")
