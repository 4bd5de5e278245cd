"""BASELINE config 1 (1 M x 512 fp32 rows, C = Q = 200, T2T top-500): step breakdown.  Run plain for event timings, or
under `ncu --metrics gpu__time_duration.sum` for the per-kernel launch list."""
import json, sys, time
import torch
sys.path.insert(0, ".")
from swat_b200 import _lib, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = torch.device("cuda", 0)
ctx = _lib.Context(0)
qc, q, _ = synth.make_queries(200, 1, seed=3, dtype=torch.float32)
cap, img, _ = synth.make_bank(N, qc, seed=3, device=dev, dtype=torch.float32, chunk=1 << 18)
qs = _lib.Queries(ctx, q)
torch.cuda.synchronize()
for name, kw in (("t2t", {}), ("t2t+t2i", {"t2i_bank": img, "t2i_threshold": 0.25})):
    for i in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        _lib.topk(ctx, qs, cap, 500, 0.0, **kw)
        e1.record(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(name, i, f"wall {dt*1e3:.3f} ms events {e0.elapsed_time(e1):.3f} ms", json.dumps(ctx.last_timing()), flush=True)
