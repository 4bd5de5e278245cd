"""Steady-state step time of swat_topk (T2T and T2T+T2I) vs the size of the threshold-bootstrap prefix."""
import sys, statistics, torch
sys.path.insert(0, ".")
from swat_b200 import _lib, synth
N = 10_000_000
dev = torch.device("cuda", 0)
ctx = _lib.Context(0)
qc, queries, _ = synth.make_queries(200, 1, seed=0, dtype=torch.bfloat16)
cap, img, _ = synth.make_bank(N, qc, seed=0, device=dev, dtype=torch.bfloat16, chunk=1 << 20)
qs = _lib.Queries(ctx, queries.float())
def t1(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)
opts = (0, 8192, 16384, 32768, 65536, 131072)
for name, kw in (("t2t", {}), ("t2t+t2i", {"t2i_bank": img})):
    for _ in range(3):
        _lib.topk(ctx, qs, cap, 500, 0.0, **kw)      # settle the per-class depth hints
    res = {o: [] for o in opts}
    for rep in range(7):
        for o in (opts if rep % 2 == 0 else opts[::-1]):
            ctx.set_option("bootstrap_rows", o)
            ms = t1(lambda: _lib.topk(ctx, qs, cap, 500, 0.0, **kw))
            if rep: res[o].append((ms, ctx.last_timing()["scan_ms"]))
    print(name, "  ".join(f"{o}: {statistics.median(v[0] for v in r):.3f} (scan {statistics.median(v[1] for v in r):.3f})" for o, r in res.items()), flush=True)
