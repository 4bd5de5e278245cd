"""One selecting scan per query count (run under ncu to read DRAM bytes / L2 hit rate per launch)."""
import sys, torch
sys.path.insert(0, ".")
from swat_b200 import _lib, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
dev = torch.device("cuda", 0)
ctx = _lib.Context(0)
ctx.set_option("bootstrap_rows", 0)
qc, _, _ = synth.make_queries(64, 1, seed=0, dtype=torch.bfloat16)
cap, _, _ = synth.make_bank(N, qc, seed=0, device=dev, dtype=torch.bfloat16, chunk=1 << 20, with_images=False)
for Q in [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "200,400,1000,4096").split(",")]:
    _, queries, _ = synth.make_queries(Q, 1, seed=1, dtype=torch.bfloat16)
    qs = _lib.Queries(ctx, queries.float())
    job = _lib.Job(ctx, qs, 500, 0.0)
    job.scan(cap); torch.cuda.synchronize()
    print("Q", Q, "done", flush=True)
    job.close(); qs.close()
