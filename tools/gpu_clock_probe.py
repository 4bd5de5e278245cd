"""SM clock the scan kernel observes (SWAT_DEBUG=2) on ONE busy GPU, imagenet shape: where the automatic lockstep threshold sits."""
import os, sys, torch
os.environ["SWAT_DEBUG"] = "2"
sys.path.insert(0, ".")
from swat_b200 import _lib, synth
dev = torch.device("cuda", 0)
ctx = _lib.Context(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
qc, q, _ = synth.make_queries(1000, 1, seed=1, dtype=torch.bfloat16)
cap, _, _ = synth.make_bank(N, qc, seed=1, device=dev, dtype=torch.bfloat16, chunk=1 << 20, with_images=False)
qs = _lib.Queries(ctx, q.float())
for i in range(12):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); _lib.topk(ctx, qs, cap, 500, 0.0); e1.record(); torch.cuda.synchronize()
    print(f"step {i}: {e0.elapsed_time(e1):.2f} ms", file=sys.stderr, flush=True)
