"""fp32 banks on the tcgen05 scan (converter warps): scan-only timing for the two shared-memory splits."""
import sys
import torch
sys.path.insert(0, ".")
from swat_b200 import _lib, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
dev = torch.device("cuda", 0)
qc, q, _ = synth.make_queries(200, 1, seed=3, dtype=torch.float32)
cap, _, _ = synth.make_bank(N, qc, seed=3, device=dev, dtype=torch.float32, chunk=1 << 18, with_images=False)
def ev(fn, reps=4):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2]
ref = None
for op in (3, 2):
    ctx = _lib.Context(0, f32_op_stages=op)
    qs = _lib.Queries(ctx, q)
    eps = _lib.scan_eps(qs, torch.float32)
    job = _lib.Job(ctx, qs, 1536, -eps)
    ms = ev(lambda: (job.reset(), job.scan(cap)))
    s, r, c, t = job.select()
    same = None if ref is None else bool(torch.equal(r, ref))
    ref = r.clone() if ref is None else ref
    print(f"f32_op_stages={op} eps={eps:.5f}: scan {ms:.3f} ms  {N * 2048 / ms / 1e6:.0f} GB/s  frac {N * 2048 / ms / 1e6 / 6537:.3f}  overflow={job.overflowed()} same_rows={same}", flush=True)
    j2 = _lib.Job(ctx, qs, 500, 0.999)
    ms = ev(lambda: (j2.reset(), j2.scan(cap)))
    print(f"   thr=0.999 (no survivors): scan {ms:.3f} ms {N * 2048 / ms / 1e6:.0f} GB/s", flush=True)
    j2.close(); job.close(); qs.close(); ctx.close()
