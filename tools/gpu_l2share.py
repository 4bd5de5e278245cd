"""One selecting scan per (query count, lock_window) -- run under ncu to read DRAM bytes / L2 hit rate per launch:
   ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum -k regex:scan_tc --csv ... python tools/gpu_l2share.py"""
import sys, torch
sys.path.insert(0, ".")
from swat_b200 import _lib, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dev = torch.device("cuda", 0)
ctx = _lib.Context(0)
ctx.set_option("bootstrap_rows", 0)
qc, _, _ = synth.make_queries(1000, 1, seed=0, dtype=torch.bfloat16)
cap, _, _ = synth.make_bank(N, qc, seed=0, device=dev, dtype=torch.bfloat16, chunk=1 << 20, with_images=False)
for Q in [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "200,400,1000").split(",")]:
    qs = _lib.Queries(ctx, qc[:Q].float())
    for lw in (0, 4):
        ctx.set_option("lock_window", lw)
        job = _lib.Job(ctx, qs, 576, -1e-4)
        job.scan(cap); torch.cuda.synchronize()       # first scan: thresholds warm up
        job.reset(); job.scan(cap); torch.cuda.synchronize()
        print("Q", Q, "lock_window", lw, "done", flush=True)
        job.close()
    qs.close()
