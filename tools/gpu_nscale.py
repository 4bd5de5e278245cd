"""Scan time per row vs bank size (Q=200): looks for scale-dependent effects (TLB, power, list growth)."""
import sys, subprocess, torch
sys.path.insert(0, ".")
from swat_b200 import _lib, synth
dev = torch.device("cuda", 0)
ctx = _lib.Context(0)
qc, queries, _ = synth.make_queries(200, 1, seed=0, dtype=torch.bfloat16)
NMAX = 100_000_000
cap, _, _ = synth.make_bank(NMAX, qc, seed=0, device=dev, dtype=torch.bfloat16, chunk=1 << 20, with_images=False)
qs = _lib.Queries(ctx, queries.float())
def t1(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)
def clk():
    return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
for thr, kf in ((0.999, 500), (0.0, 500)):
    job = _lib.Job(ctx, qs, kf, thr)
    for rep in range(2):
        for n in (10_000_000, 20_000_000, 50_000_000, 100_000_000):
            for off in (0, NMAX - n):
                v = cap[off:off + n]
                ms = min(t1(lambda: (job.reset(), job.scan(v))) for _ in range(3))
                print(f"thr={thr} N={n} off={off}: {ms:.3f} ms  {n/ms/1e6:.3f} G rows/s  {n*1024/ms/1e6:.0f} GB/s  [{clk()}]", flush=True)
    job.close()
