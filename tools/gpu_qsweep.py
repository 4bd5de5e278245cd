"""Query-count sweep (BASELINE config 5 shape) + config-1 fp32 path timing on one GPU."""
import json, sys, torch
sys.path.insert(0, ".")
from swat_b200 import _lib, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dev = torch.device("cuda", 0)
peaks = json.load(open("MEASURED_PEAKS.json")) if __import__("os").path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650, "bf16_tflops_sustained": 1400}
hbm, tf = peaks["hbm_gbs"] * 1e9, peaks["bf16_tflops_sustained"] * 1e12
ctx = _lib.Context(0)
# one bank built around 8192 class vectors; the sweep's Q queries are the first Q of them, so every class has its
# ~0.05 N / 8192 relevant rows whatever Q is (queries unrelated to the bank are the worst case for a streaming
# top-k: ~2.5x more survivors)
qc, _, _ = synth.make_queries(8192, 1, seed=0, dtype=torch.bfloat16)
cap, _, _ = synth.make_bank(N, qc, seed=0, device=dev, dtype=torch.bfloat16, chunk=1 << 20, with_images=False)
def timeit(fn, reps=4):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2]
print(f"N={N} rows; roof = min(HBM {hbm/1e9:.0f} GB/s / 1KB, {tf/1e12:.0f} TF / (1024 Q))")
md = [f"# Query-count sweep (BASELINE config 5 shape): {N:,} x 512 bf16 rows resident on 1 x B200, C = Q, top-500 selecting scan", "",
      f"Roofline = slower of 1 KB/row at {hbm/1e9:.0f} GB/s (measured copy bandwidth; a read-only stream can exceed it) and "
      f"Q x 1024 flop/row at {tf/1e12:.1f} TF/s (measured sustained bf16).  CUDA events, median of 4.", "",
      "| Q | scan ms | G rows/s | roof G rows/s | frac | bound |", "|---:|---:|---:|---:|---:|---|"]
for Q in (16, 64, 128, 200, 256, 400, 512, 1000, 1024, 2048, 4096, 8192):
    queries = qc[:Q]
    qs = _lib.Queries(ctx, queries.float())
    job = _lib.Job(ctx, qs, 500, 0.0)
    ms = timeit(lambda: (job.reset(), job.scan(cap)))
    roof = min(hbm / 1024, tf / (1024.0 * Q))
    print(f"Q={Q:5d}: scan {ms:8.3f} ms  {N/ms/1e6:7.3f} G rows/s  roof {roof/1e9:6.3f} G rows/s  frac {N/ms*1e3/roof:5.3f}  overflow={job.overflowed()}", flush=True)
    md.append(f"| {Q} | {ms:.3f} | {N/ms/1e6:.3f} | {roof/1e9:.3f} | {N/ms*1e3/roof:.3f} | {'hbm' if hbm/1024 <= tf/(1024.0*Q) else 'tensor'} |")
    job.close(); qs.close()
# imagenet-like synonym groups: C=1000 classes, 5191 queries, MAX reduce
sizes = [1 + (i * 37) % 10 for i in range(1000)]
s = sum(sizes); sizes[0] += 5191 - s if 5191 - s > -sizes[0] else 0
qc2, queries, coq = synth.make_queries(1000, sizes, seed=2, dtype=torch.bfloat16)
qs = _lib.Queries(ctx, queries.float(), coq, 1000, "max")
job = _lib.Job(ctx, qs, 500, 0.0)
ms = timeit(lambda: (job.reset(), job.scan(cap)), reps=3)
Q = queries.shape[0]; roof = min(hbm / 1024, tf / (1024.0 * Q))
md.append(f"| {Q} (C=1000, MAX over synonym groups) | {ms:.3f} | {N/ms/1e6:.3f} | {roof/1e9:.3f} | {N/ms*1e3/roof:.3f} | tensor |")
__import__("os").makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/qsweep.md", "w").write("\n".join(md) + "\n")
print(f"imagenet-like C=1000 Q={Q} MAX: scan {ms:.3f} ms {N/ms/1e6:.3f} G rows/s roof {roof/1e9:.3f} frac {N/ms*1e3/roof:.3f} overflow={job.overflowed()}")
job.close(); qs.close()
# config 1: 1M x 512 fp32, C=200, exact fp32 kernel
qc3, q3, _ = synth.make_queries(200, 1, seed=3, dtype=torch.float32)
cap32, img32, _ = synth.make_bank(1_000_000, qc3, seed=3, device=dev, dtype=torch.float32, chunk=1 << 18)
qs = _lib.Queries(ctx, q3)
for name, kw in (("T2T", {}), ("T2T+T2I", {"t2i_bank": img32})):
    _lib.topk(ctx, qs, cap32, 500, 0.0, **kw)
    ms = timeit(lambda: _lib.topk(ctx, qs, cap32, 500, 0.0, **kw), reps=3)
    print(f"config1 fp32 1M x 512, C=200 {name}: {ms:.2f} ms/step  {1e6/ms/1e3:.2f} M rows/s  timing {ctx.last_timing()}")
