"""One class-mean line of the nine-dataset sweep, call by call (SWAT_DEBUG=1 shows allocations / escalations)."""
import json, os, sys, time, torch
os.environ["SWAT_DEBUG"] = "1"
sys.path.insert(0, ".")
from swat_b200 import _lib, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
dev = torch.device("cuda", 0)
ctx = _lib.Context(0)
qc, q1000, _ = synth.make_queries(1000, 1, seed=1, dtype=torch.bfloat16)
cap, _, _ = synth.make_bank(N, qc, seed=1, device=dev, dtype=torch.bfloat16, chunk=1 << 20, with_images=False)
for C in (102, 200):
    qs = _lib.Queries(ctx, qc[:C].float())
    for i in range(5):
        t0 = time.perf_counter()
        _lib.topk(ctx, qs, cap, 500, 0.0); torch.cuda.synchronize()
        print(f"C={C} call {i}: wall {(time.perf_counter()-t0)*1e3:.2f} ms", json.dumps(ctx.last_timing()), flush=True)
    qs.close()
