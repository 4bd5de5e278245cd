"""Scan-only time: one query per class against synonym groups with a MAX / MEAN reduce at the same column count
(is the grouped-reduce epilogue the bottleneck?).  CUDA events, median of 5, thresholds live (bootstrap on)."""
import sys, torch
sys.path.insert(0, ".")
from swat_b200 import _lib, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
dev = torch.device("cuda", 0)
ctx = _lib.Context(0)
qc, _, _ = synth.make_queries(256, 1, seed=0, dtype=torch.bfloat16)
cap, _, _ = synth.make_bank(N, qc, seed=0, device=dev, dtype=torch.bfloat16, chunk=1 << 20, with_images=False)
def ev_time(fn, reps=5):
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    return sorted(out)
def sizes_for(C, S):
    sizes = [S // C] * C
    for i in range(S - sum(sizes)): sizes[i] += 1
    return sizes
for C, S in ((37, 114), (64, 128), (100, 200), (128, 256)):
    for mode in ("none", "max", "mean"):
        if mode == "none":
            q = qc[:S].float(); coq = None; red = "none"
        else:
            sizes = sizes_for(C, S)
            coq = torch.repeat_interleave(torch.arange(C, dtype=torch.int32), torch.tensor(sizes))
            g = torch.Generator().manual_seed(1)
            u = torch.nn.functional.normalize(torch.randn(S, 512, generator=g), dim=-1)
            q = torch.nn.functional.normalize(qc[coq.long()].float() + 0.3 * u, dim=-1).to(torch.bfloat16).float()
            red = mode
        qs = _lib.Queries(ctx, q, class_of_query=coq, reduce=red) if coq is not None else _lib.Queries(ctx, q)
        job = _lib.Job(ctx, qs, 576, -1e-4)
        t = ev_time(lambda: (job.reset(), job.scan(cap)))
        print(f"Q={S} C={C if coq is not None else S} reduce={mode}: scan {t[len(t)//2]:.3f} ms ({N*1024/t[len(t)//2]/1e6:.0f} GB/s)  {[round(x,3) for x in t]}", flush=True)
        job.close(); qs.close()
