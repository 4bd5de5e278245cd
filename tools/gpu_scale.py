"""Scan time vs bank size and threshold: separates start-up burst from steady-state survivor cost."""
import sys, torch
sys.path.insert(0, ".")
from swat_b200 import _lib, synth
C = 200
dev = torch.device("cuda", 0)
qc, queries, _ = synth.make_queries(C, 1, seed=0, dtype=torch.bfloat16)
cap, _, _ = synth.make_bank(20_000_000, qc, seed=0, device=dev, dtype=torch.bfloat16, chunk=1 << 20, with_images=False)
ctx = _lib.Context(0)
qs = _lib.Queries(ctx, queries.float())
def t(n, kf, thr):
    job = _lib.Job(ctx, qs, kf, thr); ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        job.reset(); e0.record(); job.scan(cap[:n]); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    job.close(); ts.sort(); return ts[2]
for n in (1_250_000, 2_500_000, 5_000_000, 10_000_000, 20_000_000):
    print(f"N={n}: floor {t(n,500,0.999):.3f}  k500 {t(n,500,0.0):.3f}  k500@thr0.2 {t(n,500,0.2):.3f}  k1024 {t(n,1024,0.0):.3f}  k4096 {t(n,4096,0.0):.3f} ms")
