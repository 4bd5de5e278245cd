"""Experiment: uniform pairs per Q block and TMA L2 prefetch distance for multi-Q-block (tensor-bound) scans."""
import sys, torch
sys.path.insert(0, ".")
from swat_b200 import _lib, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
dev = torch.device("cuda", 0)
ctx = _lib.Context(0)
qc, _, _ = synth.make_queries(64, 1, seed=0, dtype=torch.bfloat16)
cap, _, _ = synth.make_bank(N, qc, seed=0, device=dev, dtype=torch.bfloat16, chunk=1 << 20, with_images=False)
def timeit(fn, reps=3):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2]
import statistics
def t1(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)
cfgs = [(0,), (1,)]
for Q in (400, 1000, 1221, 2048, 4096, 5191, 8192):
    _, queries, _ = synth.make_queries(Q, 1, seed=1, dtype=torch.bfloat16)
    qs = _lib.Queries(ctx, queries.float())
    job = _lib.Job(ctx, qs, 500, 0.0)
    res = {c: [] for c in cfgs}
    for rep in range(6):
        for c in (cfgs if rep % 2 == 0 else cfgs[::-1]):       # interleaved so drift (clocks, temperature) hits every config alike
            ctx.set_option("unit_plan", c[0])
            ms = t1(lambda: (job.reset(), job.scan(cap)))
            if rep > 0: res[c].append(ms)
    print(f"Q={Q}: " + "  ".join(f"unit_plan={c[0]}: {statistics.median(v):.2f}/{min(v):.2f}" for c, v in res.items()), flush=True)
    job.close(); qs.close()
