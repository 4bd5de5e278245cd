"""Two Q blocks where two HBM passes cost more than the MMAs (256 < Q <~ 400): scan time vs the lockstep window,
interleaved (CUDA events, median of 5)."""
import sys, torch
sys.path.insert(0, ".")
from swat_b200 import _lib, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
dev = torch.device("cuda", 0)
ctx = _lib.Context(0)
qc, _, _ = synth.make_queries(1000, 1, seed=0, dtype=torch.bfloat16)
cap, _, _ = synth.make_bank(N, qc, seed=0, device=dev, dtype=torch.bfloat16, chunk=1 << 20, with_images=False)
def ev_time(fn, reps=5):
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    return sorted(out)
for Q in (271, 343, 400, 512):
    qsq = _lib.Queries(ctx, qc[:Q].float())
    jq = _lib.Job(ctx, qsq, 576, -1e-4)
    hbm = N * 1024 / 6.537e12 * 1e3
    tens = N * 2 * 512 * Q / 1.3665e15 * 1e3
    print(f"Q={Q}: 1x HBM {hbm:.2f} ms, tensor {tens:.2f} ms", flush=True)
    for lw in (0, 4, 8, 16, 32, 0, 4, 8, 16, 32):
        ctx.set_option("lock_window", lw)
        t = ev_time(lambda: (jq.reset(), jq.scan(cap)))
        print(f"Q={Q} lock_window={lw}: scan {t[len(t)//2]:.3f} ms  ({[round(x,3) for x in t]})", flush=True)
    jq.close(); qsq.close()
