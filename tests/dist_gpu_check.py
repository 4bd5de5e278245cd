"""Multi-GPU check, run under torchrun on the GPU box (tools/gpu_multi.sh):
every rank scans its row shard, ONE NCCL all-gather, merge walk; the result must equal the
single-GPU pipeline over the whole bank bit for bit (shard-count invariance)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from swat_b200 import _lib, synth
from swat_b200 import dist as sdist


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    ctx = _lib.Context(lr)
    N, C, k = int(os.environ.get('CHECK_ROWS', 4_000_000)), 200, 500
    chunk = 1 << 18
    assert N % (world * chunk) == 0 or world == 1 or True
    qc, queries, _ = synth.make_queries(C, 1, seed=7, dtype=torch.bfloat16)
    qs = _lib.Queries(ctx, queries.float())
    a, b = sdist.shard_range(N, rank, world)
    # the full bank is generated chunk by chunk with per-chunk seeds, so a shard that starts on a chunk
    # boundary equals the matching slice of the full bank
    a = a // chunk * chunk if rank else 0
    b = N if rank == world - 1 else (sdist.shard_range(N, rank + 1, world)[0] // chunk * chunk)
    cap, img, _ = synth.make_bank(b - a, qc, seed=7, device=dev, dtype=torch.bfloat16, chunk=chunk, row_offset=a, tie_block=0)
    ok = True
    def log(*a):
        print(f'[rank {rank}]', *a, flush=True)
    for t2i in (None, img):
        log('sharded start', t2i is not None)
        res = sdist.topk_sharded(ctx, qs, cap, k, 0.0, t2i_bank=t2i, row_offset=a, world=world)
        log('sharded done')
        if rank == 0:
            fcap, fimg, _ = synth.make_bank(N, qc, seed=7, device=dev, dtype=torch.bfloat16, chunk=chunk, tie_block=0)
            full = _lib.topk(ctx, qs, fcap, k, 0.0, t2i_bank=None if t2i is None else fimg)
            bit = torch.equal(res[1], full[1]) and torch.equal(res[3], full[3]) and torch.equal(res[0], full[0])
            if t2i is not None:
                bit = bit and torch.equal(res[2], full[2])
            # scores are canonical (one fixed-order fp32 dot per returned row), so sharded and single-GPU results must
            # agree bit for bit whichever engine or escalation path either side took; `same` only explains a failure
            same = torch.equal(res[3], full[3]) and torch.allclose(res[0], full[0], atol=2e-5)
            mism = int(((res[1] != full[1]) & (full[1] >= 0)).sum())
            for c in ((res[1] != full[1]).any(1)).nonzero().flatten().tolist():
                n = int(full[3][c])
                same = same and set(res[1][c, :n].tolist()) == set(full[1][c, :n].tolist())
            print(f"world={world} {'T2T+T2I' if t2i is not None else 'T2T'}: sharded == single-GPU: bit-identical {bit}, parity {same} "
                  f"({mism} positions swapped between near-ties); accepted {int(full[3].sum())}", flush=True)
            ok = ok and bit
            del fcap, fimg
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
