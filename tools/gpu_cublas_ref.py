"""What a plain library GEMM of the same shape costs: torch.matmul (cuBLAS) bank[N,512] x Q^T[512,Q] -> scores[N,Q] bf16,
no selection, the score matrix written to HBM -- next to the fused scan (scores never leave the SM, top-k' kept)."""
import sys, statistics, torch
sys.path.insert(0, ".")
from swat_b200 import _lib, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dev = torch.device("cuda", 0)
ctx = _lib.Context(0)
qc, _, _ = synth.make_queries(1024, 1, seed=0, dtype=torch.bfloat16)
cap, _, _ = synth.make_bank(N, qc, seed=0, device=dev, dtype=torch.bfloat16, chunk=1 << 20, with_images=False)
def t1(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)
print(f"| Q | cuBLAS GEMM ms (writes {N:,} x Q bf16) | TFLOP/s | fused scan + top-500 ms | TFLOP/s |")
print("|---:|---:|---:|---:|---:|")
for Q in (64, 200, 256, 400, 512, 1000, 1024):
    q = qc[:Q].to(dev)
    out = torch.empty(N, Q, dtype=torch.bfloat16, device=dev)
    qs = _lib.Queries(ctx, qc[:Q].float())
    job = _lib.Job(ctx, qs, 500, 0.0)
    a, b = [], []
    for rep in range(6):
        ms_g = t1(lambda: torch.matmul(cap, q.t(), out=out))
        ms_s = t1(lambda: (job.reset(), job.scan(cap)))
        if rep: a.append(ms_g); b.append(ms_s)
    g, s = statistics.median(a), statistics.median(b)
    fl = 2.0 * N * 512 * Q
    print(f"| {Q} | {g:.2f} | {fl / g / 1e9:.0f} | {s:.2f} | {fl / s / 1e9:.0f} |", flush=True)
    job.close(); qs.close(); del out
