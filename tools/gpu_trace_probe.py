"""Phase stamps of the scan kernel (SWAT_SCAN_TRACE=1): where the fixed cost of a launch goes, and how far apart the
pairs finish under the static and the dynamic tile plan."""
import os, sys, torch
os.environ["SWAT_SCAN_TRACE"] = "1"
sys.path.insert(0, ".")
from swat_b200 import _lib, synth
dev = torch.device("cuda", 0)
ctx = _lib.Context(0)
qc, q, _ = synth.make_queries(200, 1, seed=0, dtype=torch.bfloat16)
cap, _, _ = synth.make_bank(10_000_000, qc, seed=0, device=dev, dtype=torch.bfloat16, chunk=1 << 20, with_images=False)
qs = _lib.Queries(ctx, q.float())
for thr, boot in ((0.999, 0), (0.0, 32768)):
    ctx.set_option("bootstrap_rows", boot)
    job = _lib.Job(ctx, qs, 1024, thr)       # thr 0.999: thresholds closed, no survivors, the pure pipeline
    for dyn in (0, 1, 0, 1):
        ctx.set_option("dyn_tiles", dyn)
        for n in (1_000_000, 10_000_000):
            print(f"thr={thr} dyn_tiles={dyn} n={n}", file=sys.stderr, flush=True)
            job.reset(); job.scan(cap[:n]); torch.cuda.synchronize()
    job.close()
