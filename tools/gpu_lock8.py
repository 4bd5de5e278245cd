"""imagenet shape (C = Q = 1000, T2T-500, 50 M rows per GPU) under torchrun on all GPUs of the box: sharded step time with
the lockstep window off / on / automatic, interleaved.  With 8 GPUs busy the box is power-capped: held in step the pairs of
a tile range read the bank from HBM once, and the saved traffic buys clock."""
import os, sys, torch
if int(os.environ.get("RANK", "0")) == 0:
    os.environ["SWAT_DEBUG"] = "1"
import torch.distributed as dist
sys.path.insert(0, ".")
from swat_b200 import _lib, synth
from swat_b200 import dist as sdist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
ctx = _lib.Context(local)
qc, q, _ = synth.make_queries(1000, 1, seed=1, dtype=torch.bfloat16)
cap, _, _ = synth.make_bank(N, qc, seed=1, device=dev, dtype=torch.bfloat16, chunk=1 << 20, with_images=False, row_offset=rank * N)
qs = _lib.Queries(ctx, q.float())
def run(reps):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        sdist.topk_sharded(ctx, qs, cap, 500, 0.0, row_offset=rank * N, world=world)
    e1.record(); dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
run(3)
for lw in (-1, 0, 4, -1, 0, 4):
    ctx.set_option("lock_window", lw)
    run(2)
    ms = run(6)
    if rank == 0:
        print(f"world={world} lock_window={lw}: {ms:.2f} ms per step ({N * world / ms / 1e6:.2f} G rows/s)", flush=True)
dist.destroy_process_group()
