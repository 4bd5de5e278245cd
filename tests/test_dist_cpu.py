"""Host-side logic of the row-sharded multi-GPU path on CPU: world_size-2 gloo all-gather of packed
candidate lists + merge walk (the CUDA merge is replaced by the oracle through `merge_fn`)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def test_shard_range_partitions_rows():
    from swat_b200.dist import shard_range
    for n, w in ((10, 3), (400_000_000, 8), (7, 8), (0, 2), (1_000_001, 4)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_roundtrip():
    from swat_b200.dist import pack, unpack
    g = torch.Generator().manual_seed(0)
    C, k, W = 5, 7, 3
    parts = []
    for w in range(W):
        parts.append((torch.randn(C, k, generator=g), torch.randint(-1, 2 ** 40, (C, k), generator=g),
                      torch.randn(C, k, generator=g), torch.randint(0, k + 1, (C,), generator=g, dtype=torch.int32),
                      torch.randint(0, 2, (C,), generator=g, dtype=torch.int32)))
    buf = torch.cat([pack(*p) for p in parts])
    s, r, t, c, tr = unpack(buf, W, C, k, True)
    for w in range(W):
        assert torch.equal(s[w], parts[w][0]) and torch.equal(r[w], parts[w][1]) and torch.equal(t[w], parts[w][2])
        assert torch.equal(c[w], parts[w][3]) and torch.equal(tr[w], parts[w][4])
    buf = torch.cat([pack(p[0], p[1], None, p[3], p[4]) for p in parts])
    s, r, t, c, tr = unpack(buf, W, C, k, False)
    assert t is None and torch.equal(r[2], parts[2][1])


def test_packed_layout_alignment():
    """int64 rows must stay 8-byte aligned in every rank's slice of the gathered buffer; views alias the buffer."""
    from swat_b200.dist import PackedCandidates, packed_layout
    for C, k, t in ((5, 7, True), (5, 7, False), (3, 1, False), (200, 1024, True), (1, 1, True)):
        lay = packed_layout(C, k, t)
        assert lay["len"] % 2 == 0 and lay["rows"] == 0 and lay["flags"] < lay["len"]
        p = PackedCandidates(C, k, t, "cpu")
        p.rows.fill_(2 ** 40 + 3); p.scores.fill_(1.5); p.counts.fill_(k); p.trunc.fill_(1); p.flags.fill_(3)
        if t:
            p.t2i.fill_(0.25)
        q = PackedCandidates(C, k, t, "cpu", buf=p.buf.clone())
        assert int(q.rows[C - 1, k - 1]) == 2 ** 40 + 3 and float(q.scores[0, 0]) == 1.5 and int(q.flags[0]) == 3
        assert int(q.counts[C - 1]) == k and int(q.trunc[0]) == 1 and (not t or float(q.t2i[C - 1, 0]) == 0.25)


def _oracle_merge(scores, rows, t2i, counts, trunc, k, thr):
    """CPU stand-in for swat_merge_topk with the same contract (predicate walk + frontier check)."""
    G, C, kf = scores.shape
    out_s = torch.zeros(C, k); out_r = torch.full((C, k), -1, dtype=torch.int64); out_t = torch.zeros(C, k)
    out_c = torch.zeros(C, dtype=torch.int32); inc = torch.zeros(C, dtype=torch.int32)
    for c in range(C):
        ent, frontier = [], None
        for g in range(G):
            n = int(counts[g, c])
            for j in range(n):
                if t2i is None or float(t2i[g, c, j]) >= thr:
                    ent.append((-float(scores[g, c, j]), int(rows[g, c, j]), float(t2i[g, c, j]) if t2i is not None else 0.0))
            if int(trunc[g, c]) and n > 0:
                f = (-float(scores[g, c, n - 1]), int(rows[g, c, n - 1]))
                frontier = f if frontier is None or f < frontier else frontier
        ent.sort()
        ent = ent[:k]
        for i, (ns, r, t) in enumerate(ent):
            out_s[c, i], out_r[c, i], out_t[c, i] = -ns, r, t
        out_c[c] = len(ent)
        if frontier is not None and (len(ent) < k or (ent[-1][0], ent[-1][1]) > frontier):
            inc[c] = 1
    return out_s, out_r, out_t if t2i is not None else None, out_c, inc


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import swat_oracle as so
    from swat_b200 import dist as sdist, synth
    N, C, k, kf = 6000, 5, 30, 120
    qc, queries, _ = synth.make_queries(C, 1, seed=2, dtype=torch.bfloat16)
    cap, img, _ = synth.make_bank(N, qc, seed=2, dtype=torch.bfloat16, rho=0.4, tie_block=60, chunk=1 << 12)
    capf, imgf, qf = cap.float().numpy(), img.float().numpy(), queries.float().numpy()
    a, b = sdist.shard_range(N, rank, world)
    # local stage computed by the oracle on this rank's rows: T2T top-kf candidates + their T2I
    S = so.score_matrix(capf[a:b], qf); I = so.score_matrix(imgf[a:b], qf)
    rows = torch.full((C, kf), -1, dtype=torch.int64); sc = torch.zeros(C, kf); ti = torch.zeros(C, kf)
    cnt = torch.zeros(C, dtype=torch.int32); tr = torch.zeros(C, dtype=torch.int32)
    for c in range(C):
        sel = so.select_walk(S[:, c], kf, 0.0)
        rows[c, :sel.size] = torch.from_numpy(sel + a); sc[c, :sel.size] = torch.from_numpy(S[sel, c]); ti[c, :sel.size] = torch.from_numpy(I[sel, c])
        cnt[c] = sel.size; tr[c] = int((S[:, c] >= 0).sum() > kf)
    res = sdist.gather_merge((sc, rows, ti, cnt, tr), k, 0.25, world, merge_fn=_oracle_merge)
    full = so.topk_walk(capf, qf, k, 0.0, t2i_bank=imgf, t2i_threshold=0.25)
    ok = all(res[1][c, :int(res[3][c])].tolist() == full[0][c, :full[3][c]].tolist() for c in range(C))
    ok = ok and res[3].tolist() == full[3].tolist() and int(res[4].sum()) == 0
    # T2T only path (no aux): top-k of the union
    res2 = sdist.gather_merge((sc[:, :k].contiguous(), rows[:, :k].contiguous(), None, torch.minimum(cnt, torch.tensor(k, dtype=torch.int32)), tr),
                              k, 0.25, world, merge_fn=_oracle_merge)
    full2 = so.topk_walk(capf, qf, k, 0.0)
    ok = ok and all(res2[1][c, :int(res2[3][c])].tolist() == full2[0][c, :full2[3][c]].tolist() for c in range(C))
    # packed exchange buffer written in place (what the GPU path does): same result, and every rank sees every rank's flags
    pk = sdist.PackedCandidates(C, kf, True, "cpu")
    pk.scores.copy_(sc); pk.rows.copy_(rows); pk.t2i.copy_(ti); pk.counts.copy_(cnt); pk.trunc.copy_(tr); pk.flags.fill_(rank * 2)
    gathered = sdist.gather_packed(pk.buf, world)
    res3 = sdist.merge_packed(gathered, pk.lay, world, k, 0.25, merge_fn=_oracle_merge)
    ok = ok and torch.equal(res3[1], res[1]) and torch.equal(res3[3], res[3]) and torch.equal(res3[0], res[0])
    ok = ok and sdist.unpack_flags(gathered, world, C, kf, True).tolist() == [2 * r for r in range(world)]
    ret[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.timeout(240)
def test_world2_gloo_gather_merge_matches_single_shard():
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(200)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert ret.get(0) is True and ret.get(1) is True
