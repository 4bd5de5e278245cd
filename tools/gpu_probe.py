"""Timing probe on the GPU box: scan-only time for several configurations + pipeline breakdown."""
import json, sys, time
import torch
sys.path.insert(0, ".")
from swat_b200 import _lib, synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
C = int(sys.argv[2]) if len(sys.argv) > 2 else 200
dev = torch.device("cuda", 0)
ctx = _lib.Context(0)
qc, queries, _ = synth.make_queries(C, 1, seed=0, dtype=torch.bfloat16)
cap, img, _ = synth.make_bank(N, qc, seed=0, device=dev, dtype=torch.bfloat16, chunk=1 << 20)
qs = _lib.Queries(ctx, queries.float())
torch.cuda.synchronize()

def ev_time(fn, reps=5):
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    return sorted(out)

for kf in (500, 1024, 4096):
    job = _lib.Job(ctx, qs, kf, 0.0)
    def scan():
        job.reset(); job.scan(cap)
    t = ev_time(scan)
    sel = ev_time(lambda: job.select())
    print(f"scan k_fetch={kf}: {t} ms  -> {N*1024/t[len(t)//2]/1e6:.0f} GB/s ; select {sel} ; overflow={job.overflowed()}")
    job.close()
# dense-free "pure GEMM" ceiling: threshold so high nothing survives
job = _lib.Job(ctx, qs, 500, 0.999)
t = ev_time(lambda: (job.reset(), job.scan(cap)))
print(f"scan thr=0.999 (no survivors): {t} ms -> {N*1024/t[len(t)//2]/1e6:.0f} GB/s")
job.close()
for name, kw in (("t2t", {}), ("t2t+t2i", {"t2i_bank": img})):
    for _ in range(2):
        t0 = time.perf_counter()
        _lib.topk(ctx, qs, cap, 500, 0.0, **kw)
        dt = time.perf_counter() - t0
        print(name, f"wall {dt*1e3:.2f} ms", json.dumps(ctx.last_timing()))
# tail breakdown: candidates -> exact re-score + walk
from swat_b200 import dist as sdist
job = _lib.Job(ctx, qs, 1024, -1e-4)
job.reset(); job.scan(cap)
sc, rw, cn, tr = job.select()
t = ev_time(lambda: _lib.rescore_walk(ctx, qs, cap, sc, rw, cn, tr, 500, 0.0, aux_bank=img))
print("rescore+walk (1024 candidates x 200 classes, both banks):", t)
job.close()
for _ in range(3):
    t0 = time.perf_counter(); _lib.topk(ctx, qs, cap, 500, 0.0, t2i_bank=img); dt = time.perf_counter() - t0
    print("t2t+t2i", f"wall {dt*1e3:.2f} ms", json.dumps(ctx.last_timing()))
# host pipeline: candidates' rows read in place from pinned banks (zero-copy) vs gathered on the host
h_cap = torch.empty(cap.shape, dtype=cap.dtype, pin_memory=True); h_cap.copy_(cap)
h_img = torch.empty(img.shape, dtype=img.dtype, pin_memory=True); h_img.copy_(img)
torch.cuda.synchronize()
for zc in (1, 0, 1, 0):
    ctx.set_option("zero_copy", zc)
    t0 = time.perf_counter(); _lib.topk_host(ctx, qs, h_cap, 500, 0.0, t2i_bank=h_img); dt = time.perf_counter() - t0
    print(f"topk_host zero_copy={zc}: wall {dt*1e3:.1f} ms  {N/dt/1e6:.1f} M rows/s", json.dumps(ctx.last_timing()))
# several Q blocks: pairs sharing a tile range with / without the lockstep window
del h_cap, h_img
for Q in (400, 1000):
    qcq, qq, _ = synth.make_queries(Q, 1, seed=5, dtype=torch.bfloat16)
    qsq = _lib.Queries(ctx, qq.float())
    jq = _lib.Job(ctx, qsq, 576, -1e-4)
    for lw in (0, 4, 0, 4, 2, 8):
        ctx.set_option("lock_window", lw)
        t = ev_time(lambda: (jq.reset(), jq.scan(cap)), reps=4)
        print(f"Q={Q} lock_window={lw}: scan {t[len(t)//2]:.3f} ms  ({[round(x,3) for x in t]})", flush=True)
    jq.close(); qsq.close()
ctx.set_option("lock_window", 4)
