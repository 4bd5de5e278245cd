"""Cost of a class with (almost) no rows passing T2I at benchmark scale: bank-swap pass vs the fp32 in-pass predicate."""
import sys, time, json, torch
sys.path.insert(0, ".")
from swat_b200 import _lib, synth
N = 10_000_000
dev = torch.device("cuda", 0)
qc, queries, _ = synth.make_queries(200, 1, seed=0, dtype=torch.bfloat16)
cap, img, _ = synth.make_bank(N, qc, seed=0, device=dev, dtype=torch.bfloat16, chunk=1 << 20)
g = torch.Generator().manual_seed(5)
q = queries.float().clone()
for c in (7, 99):                                  # two classes whose prompt matches nothing in the bank
    q[c] = torch.nn.functional.normalize(torch.randn(512, generator=g), dim=0)
q = q.to(torch.bfloat16).float()
for swap in (1, 0):
    ctx = _lib.Context(0, swap_pass=swap)
    qs = _lib.Queries(ctx, q)
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        s, r, t, c = _lib.topk(ctx, qs, cap, 500, 0.0, t2i_bank=img)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        tm = ctx.last_timing()
        print(json.dumps(tm)); print(f"swap_pass={swap} call {rep}: {dt * 1e3:.2f} ms  scans={tm["scan_launches"]:.0f} escalations={tm["escalations"]:.0f} "
              f"scan_ms={tm['scan_ms']:.2f}  counts[7]={int(c[7])} counts[99]={int(c[99])} total={int(c.sum())}", flush=True)
    qs.close(); ctx.close()
