"""Phase timing of swat_topk_host (pinned host banks): where the end-to-end time goes."""
import sys, time, json, torch
sys.path.insert(0, ".")
from swat_b200 import _lib, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
ctx = _lib.Context(0)
qc, queries, _ = synth.make_queries(200, 1, seed=0, dtype=torch.bfloat16)
cap, img, _ = synth.make_bank(N, qc, seed=0, device="cuda", dtype=torch.bfloat16, chunk=1 << 20)
hcap = torch.empty(cap.shape, dtype=cap.dtype, pin_memory=True); hcap.copy_(cap)
himg = torch.empty(img.shape, dtype=img.dtype, pin_memory=True); himg.copy_(img)
del cap, img
torch.cuda.synchronize()
qs = _lib.Queries(ctx, queries.float())
# raw H2D rate of one big pinned copy
d = torch.empty_like(hcap, device="cuda")
for _ in range(2):
    t0 = time.perf_counter(); d.copy_(hcap, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"raw pinned H2D: {hcap.numel() * 2 / dt / 1e9:.1f} GB/s ({dt * 1e3:.1f} ms)")
del d
for chunk in (0, 131072, 524288, 1048576):
    if chunk: ctx.set_option("host_chunk_rows", chunk)
    for rep in range(3):
        t0 = time.perf_counter()
        _lib.topk_host(ctx, qs, hcap, 500, 0.0, t2i_bank=himg)
        dt = time.perf_counter() - t0
    print(f"chunk={chunk or 'default'}: wall {dt * 1e3:.1f} ms  {N / dt / 1e6:.1f} M rows/s", json.dumps(ctx.last_timing()))
