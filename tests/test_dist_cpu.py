"""Host-side logic of the row-sharded multi-GPU path on CPU: world_size-2 gloo all-gather of packed
per-shard walk results + merge (the CUDA merge is replaced by the oracle through `merge_fn`)."""
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
        lim = torch.randn(C, generator=g); lim[w] = float("-inf")
        parts.append((torch.randn(C, k, generator=g), torch.randint(-1, 2 ** 40, (C, k), generator=g),
                      torch.randn(C, k, generator=g), torch.randint(0, k + 1, (C,), generator=g, dtype=torch.int32), lim))
    buf = torch.cat([pack(*p) for p in parts])
    s, r, t, c, lim = unpack(buf, W, C, k, True)
    for w in range(W):
        assert torch.equal(s[w], parts[w][0]) and torch.equal(r[w], parts[w][1]) and torch.equal(t[w], parts[w][2])
        assert torch.equal(c[w], parts[w][3]) and torch.equal(lim[w], parts[w][4])
    buf = torch.cat([pack(p[0], p[1], None, p[3], p[4]) for p in parts])
    s, r, t, c, lim = unpack(buf, W, C, k, False)
    assert t is None and torch.equal(r[2], parts[2][1])


def test_packed_layout_alignment():
    """int64 rows must stay 8-byte aligned in every rank's slice of the gathered buffer; views alias the buffer."""
    from swat_b200.dist import PackedResults, packed_layout
    for C, k, t in ((5, 7, True), (5, 7, False), (3, 1, False), (200, 500, True), (1, 1, True)):
        lay = packed_layout(C, k, t)
        assert lay["len"] % 2 == 0 and lay["rows"] == 0 and lay["flags"] < lay["len"]
        p = PackedResults(C, k, t, "cpu")
        p.rows.fill_(2 ** 40 + 3); p.scores.fill_(1.5); p.counts.fill_(k); p.limit.fill_(float("-inf")); p.flags.fill_(3)
        if t:
            p.t2i.fill_(0.25)
        q = PackedResults(C, k, t, "cpu", buf=p.buf.clone())
        assert int(q.rows[C - 1, k - 1]) == 2 ** 40 + 3 and float(q.scores[0, 0]) == 1.5 and int(q.flags[0]) == 3
        assert int(q.counts[C - 1]) == k and float(q.limit[0]) == float("-inf") and (not t or float(q.t2i[C - 1, 0]) == 0.25)


def test_default_k_fetch_mirrors_the_library():
    from swat_b200.dist import default_k_fetch
    assert default_k_fetch(500, False, 1e-4) == 576 and default_k_fetch(500, True, 1e-4) == 1024
    assert default_k_fetch(500, False, 5e-3) == 1536 and default_k_fetch(500, True, 5e-3) == 2048
    assert default_k_fetch(4096, False, 1e-4) == 4096 and default_k_fetch(1, False, 1e-4) == 96


def _oracle_merge(scores, rows, t2i, counts, limit, k):
    """CPU stand-in for swat_merge_topk with the same contract: k best of the union under (score desc, row asc);
    incomplete when the result reaches down to some shard's limit."""
    G, C, kin = scores.shape
    out_s = torch.zeros(C, k); out_r = torch.full((C, k), -1, dtype=torch.int64); out_t = torch.zeros(C, k)
    out_c = torch.zeros(C, dtype=torch.int32); inc = torch.zeros(C, dtype=torch.int32)
    for c in range(C):
        ent = []
        for g in range(G):
            for j in range(int(counts[g, c])):
                ent.append((-float(scores[g, c, j]), int(rows[g, c, j]), float(t2i[g, c, j]) if t2i is not None else 0.0))
        ent.sort()
        ent = ent[:k]
        for i, (ns, r, t) in enumerate(ent):
            out_s[c, i], out_r[c, i], out_t[c, i] = -ns, r, t
        out_c[c] = len(ent)
        lim = float(limit[:, c].max())
        if lim > float("-inf") and (len(ent) < k or -ent[-1][0] <= lim):
            inc[c] = 1
    return out_s, out_r, out_t if t2i is not None else None, out_c, inc


def _local_walk(S, I, a, k, kf, t2i_thr):
    """Oracle stand-in for one rank's local stage: T2T top-kf candidates, accept walk, limit = score of the last
    candidate when the list was truncated and fewer than k were accepted."""
    C = S.shape[1]
    from oracle import swat_oracle as so
    rows = torch.full((C, k), -1, dtype=torch.int64); sc = torch.zeros(C, k); ti = torch.zeros(C, k)
    cnt = torch.zeros(C, dtype=torch.int32); lim = torch.full((C,), float("-inf"))
    for c in range(C):
        sel = so.select_walk(S[:, c], kf, 0.0)
        trunc = int((S[:, c] >= 0).sum()) > kf
        acc = [r for r in sel.tolist() if I is None or I[r, c] >= t2i_thr]
        if trunc and len(sel):          # rows at or below the frontier are not vouched for
            acc = [r for r in acc if S[r, c] > S[sel[-1], c]]
        acc = acc[:k]
        n = len(acc)
        rows[c, :n] = torch.tensor(acc, dtype=torch.int64) + a
        sc[c, :n] = torch.from_numpy(S[acc, c]) if n else sc[c, :n]
        if I is not None and n:
            ti[c, :n] = torch.from_numpy(I[acc, c])
        cnt[c] = n
        if trunc and n < k:
            lim[c] = float(S[sel[-1], c])
    return sc, rows, (ti if I is not None else None), cnt, lim


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
    # local stage computed by the oracle on this rank's rows: walked T2T top-kf candidates + the shard's limit
    S = so.score_matrix(capf[a:b], qf); I = so.score_matrix(imgf[a:b], qf)
    local = _local_walk(S, I, a, k, kf, 0.25)
    res = sdist.gather_merge(local, k, world, merge_fn=_oracle_merge)
    full = so.topk_walk(capf, qf, k, 0.0, t2i_bank=imgf, t2i_threshold=0.25)
    ok = all(res[1][c, :int(res[3][c])].tolist() == full[0][c, :full[3][c]].tolist() for c in range(C))
    ok = ok and res[3].tolist() == full[3].tolist() and int(res[4].sum()) == 0
    # T2T only path (no aux): top-k of the union
    local2 = _local_walk(S, None, a, k, kf, 0.0)
    res2 = sdist.gather_merge(local2, k, world, merge_fn=_oracle_merge)
    full2 = so.topk_walk(capf, qf, k, 0.0)
    ok = ok and all(res2[1][c, :int(res2[3][c])].tolist() == full2[0][c, :full2[3][c]].tolist() for c in range(C))
    # a walk that is too shallow must be reported: 12 candidates per shard cannot yield 30 accepted rows
    local3 = _local_walk(S, I, a, k, 12, 0.25)
    res3 = sdist.gather_merge(local3, k, world, merge_fn=_oracle_merge)
    ok = ok and int(res3[4].sum()) > 0
    # packed exchange buffer written in place (what the GPU path does): same result, and every rank sees every rank's flags
    pk = sdist.PackedResults(C, k, True, "cpu")
    pk.scores.copy_(local[0]); pk.rows.copy_(local[1]); pk.t2i.copy_(local[2]); pk.counts.copy_(local[3]); pk.limit.copy_(local[4])
    pk.flags.fill_(rank * 2)
    gathered = sdist.gather_packed(pk.buf, world)
    res4 = sdist.merge_packed(gathered, pk.lay, world, k, merge_fn=_oracle_merge)
    ok = ok and torch.equal(res4[1], res[1]) and torch.equal(res4[3], res[3]) and torch.equal(res4[0], res[0])
    ok = ok and sdist.unpack_flags(gathered, world, C, k, True).tolist() == [2 * r for r in range(world)]
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
