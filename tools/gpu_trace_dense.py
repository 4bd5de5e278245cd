"""Phase stamps (SWAT_SCAN_TRACE=1) of a dense-mode launch the size of the bootstrap prefix (32 K rows x 200 classes)."""
import os, sys, torch
os.environ["SWAT_SCAN_TRACE"] = "1"
sys.path.insert(0, ".")
from swat_b200 import _lib, synth
dev = torch.device("cuda", 0)
ctx = _lib.Context(0)
qc, q, _ = synth.make_queries(200, 1, seed=0, dtype=torch.bfloat16)
cap, _, _ = synth.make_bank(1_000_000, qc, seed=0, device=dev, dtype=torch.bfloat16, chunk=1 << 20, with_images=False)
qs = _lib.Queries(ctx, q.float())
for n in (32768, 32768, 32768, 262144, 262144):
    out = _lib.scores_dense(ctx, qs, cap[:n], engine="tc"); torch.cuda.synchronize()
# the bootstrap's own dense launch (transposed output) followed by the selecting scan
job = _lib.Job(ctx, qs, 1024, 0.0)
for _ in range(3):
    job.reset(); job.scan(cap); torch.cuda.synchronize()
