"""Parity at BASELINE.json's config sizes (``-m gpu``), against the oracle:

* config 1 -- 1 M x 512 **fp32** caption shard, semi-aves shape (C = 200): Q = 200 class prompts and Q = 400 synonym
  pairs with MEAN and MAX, unpartitioned and Zipf-partitioned, T2T and T2T+T2I, every class against ``so.topk_walk``
  (the CPU oracle); plus the oracle's verbatim port of the reference loop (sample_retrieval.py:724-825) on the
  partitioned variant (all classes) and on two classes of the unpartitioned one.  fp32 banks run on the tcgen05 kernel
  (converter warps) and every returned score is the exact fp32 re-score.
* config 2 -- 10 M x 512 bf16 caption + image, C = Q = 200, k = 500: T2T and T2T500+T2I0.25 on all classes against a
  chunked fp32 torch-on-GPU restatement (tests/gpu_restate.py) that this module first proves equal to ``so.topk_walk``
  on the first 1 M rows; planted needles, walk invariants, idempotence, shard-count invariance (bit-identical).
* config 3 -- one 50 M x 512 bf16 bank: the imagenet line (C = Q = 1000) and the semi-aves synonym line (Q = 400, MAX),
  all classes against the restatement.

Parity rule (north star): same counts; rows identical up to swaps among scores that agree to TIE_TOL; rows on one side
only sit within 1e-3 of the k-th score; scores within 1e-3.  Interior swaps and boundary differences are counted and
bounded, not just tolerated.
"""
import numpy as np
import pytest
import torch

from oracle import swat_oracle as so
from tests.gpu_restate import compare_walks, restate_topk_walk

pytestmark = pytest.mark.gpu

TIE_TOL = 2e-6          # canonical fp32 scores vs the oracle's fp32 GEMM: summation order only
K = 500


def _report(tag, res, swap_div=500):
    swaps, boundary, flips, total = res
    print(f"[parity] {tag}: {total} rows compared, {swaps} positions swapped between near-ties (<= {TIE_TOL}), "
          f"{boundary} boundary rows differ, {flips} predicate flips (|T2I - 0.25| <= 2e-6)")
    assert swaps <= max(8, total // swap_div), f"{tag}: {swaps} near-tie swaps in {total} rows"
    assert boundary <= max(4, total // 5000), f"{tag}: {boundary} boundary differences"
    assert flips <= 3, f"{tag}: {flips} predicate flips"


# ----------------------------------------------------------------------------------------------- config 1
@pytest.fixture(scope="module")
def cfg1():
    from swat_b200 import _lib, synth
    ctx = _lib.Context(0)
    qc, q200, _ = synth.make_queries(200, 1, seed=0, dtype=torch.float32)
    _, q400, coq400 = synth.make_queries(200, 2, seed=0, dtype=torch.float32)
    banks = {}
    for part in (False, True):
        cap, img, labels = synth.make_bank(1_000_000, qc, seed=1, dtype=torch.float32, partitioned=part, chunk=1 << 18, device="cuda")
        banks[part] = (cap, img, labels.to(torch.int32))
    yield dict(lib=_lib, ctx=ctx, qc=qc, q200=q200, q400=q400, coq400=coq400, banks=banks)
    ctx.close()


@pytest.mark.parametrize("part", [False, True], ids=["unpartitioned", "zipf-partitioned"])
@pytest.mark.parametrize("qcfg", ["Q200", "Q400-mean", "Q400-max"])
def test_config1_fp32_1M_vs_cpu_oracle(cfg1, qcfg, part):
    w = cfg1
    lib, ctx = w["lib"], w["ctx"]
    cap, img, labels = w["banks"][part]
    if qcfg == "Q200":
        q, coq, red = w["q200"], None, "none"
    else:
        q, coq, red = w["q400"], w["coq400"], qcfg.split("-")[1]
    qs = lib.Queries(ctx, q, coq, 200, red)
    assert lib.scan_eps(qs, torch.float32) > 1e-3            # fp32 banks take the bf16-converting tensor-core scan
    rc = labels if part else None
    capf, imgf, qf = cap.cpu().numpy(), img.cpu().numpy(), q.numpy()
    coq_np = None if coq is None else coq.numpy()
    lab_np = labels.cpu().numpy() if part else None
    for with_t2i in (False, True):
        l0 = ctx.launch_count
        g = lib.topk(ctx, qs, cap, K, 0.0, t2i_bank=img if with_t2i else None, t2i_threshold=0.25, row_class=rc)
        o = so.topk_walk(capf, qf, K, 0.0, t2i_bank=imgf if with_t2i else None, t2i_threshold=0.25, class_of_query=coq_np,
                         n_classes=200, reduce=red, row_labels=lab_np)
        res = compare_walks(g, o, TIE_TOL, what=f"cfg1 {qcfg} part={part} t2i={with_t2i}", aux_thr=0.25 if with_t2i else None)
        _report(f"cfg1 {qcfg} part={part} t2i={with_t2i} launches={ctx.launch_count - l0} esc={ctx.last_timing()['escalations']}", res)
        if with_t2i:
            assert np.all(g[2].cpu().numpy()[g[1].cpu().numpy() >= 0] >= 0.25)
    qs.close()


def test_config1_verbatim_port(cfg1):
    """The reference's loop as written (per class: GEMV, Python sorted, accept walk): all 200 classes on the Zipf-partitioned
    shard, two classes on the unpartitioned one."""
    w = cfg1
    lib, ctx = w["lib"], w["ctx"]
    qs = lib.Queries(ctx, w["q200"])
    prompts = {str(c): {"mean": w["q200"][c].numpy()} for c in range(200)}
    # partitioned
    cap, img, labels = w["banks"][True]
    capf, imgf, lab = cap.cpu().numpy(), img.cpu().numpy(), labels.cpu().numpy()
    order = np.argsort(lab, kind="stable")
    starts = np.searchsorted(lab[order], np.arange(201))
    feats = {}
    for c in range(200):
        rows = order[starts[c]:starts[c + 1]]
        feats[str(c)] = {"file_paths": [f"/s/{c}/{r}.jpg" for r in rows.tolist()], "feats": imgf[rows], "caption_feats": capf[rows]}
    for with_t2i, fn in ((False, so.verbatim_t2t_ranked_sampler), (True, so.verbatim_t2t_ranked_t2i_tshd_sampler)):
        ms, nd = fn(prompts, K, 0.0, feats)[:2]
        g = lib.topk(ctx, qs, cap, K, 0.0, t2i_bank=img if with_t2i else None, row_class=labels)
        # the port's per-class lists (indices into the class's regrouped rows) -> [C,K] arrays of global rows
        r_rows = np.full((200, K), -1, dtype=np.int64); r_s = np.zeros((200, K), np.float32); r_t = np.zeros((200, K), np.float32)
        r_c = np.zeros(200, np.int32)
        i = 0
        for c in range(200):
            n = int(nd[str(c)])
            r_c[c] = n
            if n == 0:
                continue
            cls_rows = order[starts[c]:starts[c + 1]]
            r_rows[c, :n] = cls_rows[ms["row_list"][i]]; r_s[c, :n] = ms["score_list"][i]
            if with_t2i:
                r_t[c, :n] = ms["t2i_list"][i]
            i += 1
        res = compare_walks(g, (r_rows, r_s, r_t if with_t2i else None, r_c), TIE_TOL, what=f"cfg1 verbatim part t2i={with_t2i}",
                            aux_thr=0.25 if with_t2i else None)
        # the port (like the reference) scores each class with its own GEMV: bit-identical rows of the planted block of
        # 1000 ties come out 1 ulp apart depending on their position in the BLAS blocking, so their order is not index order
        _report(f"cfg1 verbatim port, Zipf-partitioned, t2i={with_t2i}", res, swap_div=50)
    # unpartitioned: classes 5 and 117
    cap, img, _ = w["banks"][False]
    capf, imgf = cap.cpu().numpy(), img.cpu().numpy()
    paths = [f"/s/0/{r}.jpg" for r in range(capf.shape[0])]
    sub = qs.subset([5, 117])
    g = lib.topk(ctx, sub, cap, K, 0.0, t2i_bank=img)
    feats = {str(c): {"file_paths": paths, "feats": imgf, "caption_feats": capf} for c in (5, 117)}
    ms, nd = so.verbatim_t2t_ranked_t2i_tshd_sampler({str(c): prompts[str(c)] for c in (5, 117)}, K, 0.0, feats)[:2]
    for i, c in enumerate((5, 117)):
        n = int(g[3][i])
        assert n == int(nd[str(c)])
        got = [paths[r] for r in g[1][i, :n].cpu().tolist()]
        assert sum(a != b for a, b in zip(got, ms["file_list"][i])) <= 2
    qs.close()


# ----------------------------------------------------------------------------------------------- config 2
N2, C2 = 10_000_000, 200
NEEDLE_CLASSES = [3, 17, 42, 77, 101, 150, 188, 199]


@pytest.fixture(scope="module")
def cfg2():
    from swat_b200 import _lib, synth
    dev = torch.device("cuda", 0)
    ctx = _lib.Context(0)
    qc, queries, _ = synth.make_queries(C2, 1, seed=0, dtype=torch.bfloat16)
    cap, img, _ = synth.make_bank(N2, qc, seed=0, device=dev, dtype=torch.bfloat16, chunk=1 << 20)
    needles = synth.plant_needles(cap, qc, NEEDLE_CLASSES, 600, seed=0, img=img)
    qs = _lib.Queries(ctx, queries.float())
    torch.cuda.synchronize()
    yield dict(lib=_lib, ctx=ctx, qs=qs, cap=cap, img=img, queries=queries, needles=needles, dev=dev)
    qs.close(); ctx.close()


def test_restatement_equals_cpu_oracle_at_1M(cfg2):
    """Pins the GPU restatement (the checker of the 10 M and 50 M tests) to the CPU oracle on the first 1 M rows."""
    w = cfg2
    cap, img, q = w["cap"][:1_000_000], w["img"][:1_000_000], w["queries"].float()
    capf, imgf, qf = cap.float().cpu().numpy(), img.float().cpu().numpy(), q.numpy()
    for t2i in (False, True):
        o = so.topk_walk(capf, qf, K, 0.0, t2i_bank=imgf if t2i else None, t2i_threshold=0.25)
        r = restate_topk_walk(cap, q.cuda(), K, 0.0, t2i_bank=img if t2i else None, t2i_threshold=0.25)
        res = compare_walks((r[1], r[0], r[2], r[3]), o, TIE_TOL, what=f"restatement t2i={t2i}", aux_thr=0.25 if t2i else None)
        _report(f"restatement vs CPU oracle t2i={t2i}", res)
    # synonym groups with MAX and MEAN on a 200 k prefix
    from swat_b200 import synth
    _, q400, coq = synth.make_queries(C2, 2, seed=0, dtype=torch.bfloat16)
    for red in ("max", "mean"):
        o = so.topk_walk(capf[:200_000], q400.float().numpy(), 100, 0.0, class_of_query=coq.numpy(), n_classes=C2, reduce=red)
        r = restate_topk_walk(cap[:200_000], q400.float().cuda(), 100, 0.0, class_of_query=coq, n_classes=C2, reduce=red)
        res = compare_walks((r[1], r[0], r[2], r[3]), o, TIE_TOL, what=f"restatement {red}")
        _report(f"restatement vs CPU oracle {red}", res)


def _invariants(scores, rows, counts, k, n_rows, n_cls):
    scores, rows, counts = scores.cpu().numpy(), rows.cpu().numpy(), counts.cpu().numpy()
    assert scores.shape == rows.shape == (n_cls, k) and counts.shape == (n_cls,)
    for c in range(n_cls):
        n = int(counts[c])
        assert 0 <= n <= k
        s, r = scores[c, :n], rows[c, :n]
        assert np.all(rows[c, n:] == -1) and np.all(r >= 0) and np.all(r < n_rows)
        assert len(np.unique(r)) == n
        d = np.diff(s)
        assert np.all(d <= 0), f"class {c}: scores not descending"
        ties = np.nonzero(d == 0)[0]
        assert np.all(r[ties] < r[ties + 1]), f"class {c}: ties not in ascending row order"
        assert np.all(s >= 0.0)


def _check_needles(w, rows, scores):
    """Planted rows (cosine ladder 0.90-0.99) outscore all but a stray natural row or two, so the head of their class
    is known in closed form."""
    q = w["queries"].float().cuda()
    for c in NEEDLE_CLASSES:
        got = rows[c].cpu().tolist()
        needle_rows = set(w["needles"][c].tolist())
        assert len(set(got) & needle_rows) >= K - 5, f"class {c}: planted rows missing from the head"


def test_config2_10M_bf16_vs_restatement(cfg2):
    w = cfg2
    lib, ctx, qs, cap, img = w["lib"], w["ctx"], w["qs"], w["cap"], w["img"]
    q = w["queries"].float().cuda()
    for t2i in (False, True):
        g = lib.topk(ctx, qs, cap, K, 0.0, t2i_bank=img if t2i else None, t2i_threshold=0.25)
        _invariants(g[0], g[1], g[3], K, N2, C2)
        _check_needles(w, g[1], g[0])
        r = restate_topk_walk(cap, q, K, 0.0, t2i_bank=img if t2i else None, t2i_threshold=0.25)
        res = compare_walks(g, r, TIE_TOL, what=f"cfg2 t2i={t2i}", aux_thr=0.25 if t2i else None)
        _report(f"cfg2 10M bf16 t2i={t2i} esc={ctx.last_timing()['escalations']}", res)
        if t2i:
            assert np.all(g[2].cpu().numpy()[g[1].cpu().numpy() >= 0] >= 0.25)
        again = lib.topk(ctx, qs, cap, K, 0.0, t2i_bank=img if t2i else None, t2i_threshold=0.25)       # idempotence (steady state)
        assert torch.equal(again[1], g[1]) and torch.equal(again[0], g[0]) and torch.equal(again[3], g[3])


def test_config2_shard_count_invariance_bit_identical(cfg2):
    """G in {1,2,4,8} row shards + merge == the single-shard pipeline, bit for bit (canonical scores)."""
    from swat_b200 import dist
    w = cfg2
    lib, ctx, qs, cap, img = w["lib"], w["ctx"], w["qs"], w["cap"], w["img"]
    full = lib.topk(ctx, qs, cap, K, 0.0, t2i_bank=img, t2i_threshold=0.25)
    full_t = lib.topk(ctx, qs, cap, K, 0.0)
    for G in (2, 4, 8):
        for t2i, ref in ((True, full), (False, full_t)):
            parts = []
            for r in range(G):
                a, b = dist.shard_range(N2, r, G)
                parts.append(dist.local_walk(ctx, qs, cap[a:b], K, 2048, 0.0, img[a:b] if t2i else None, 0.25, row_offset=a))
            s, rws, t, c, lim = dist.unpack(torch.cat([dist.pack(*p) for p in parts]), G, C2, K, t2i)
            ms, mr, mt, mc, inc = lib.merge_topk(ctx, s, rws, c, aux=t, limit=lim, k_out=K)
            assert int(inc.sum()) == 0, f"G={G} t2i={t2i}"
            assert torch.equal(mr, ref[1]) and torch.equal(mc, ref[3]) and torch.equal(ms, ref[0]), f"G={G} t2i={t2i}"
            if t2i:
                assert torch.equal(mt, ref[2])


def test_config2_cross_engine_bit_identical(cfg2):
    """The fp32-FMA kernel (different arithmetic, different code path, in-pass T2I predicate) ranks the candidates of a
    few classes; after the canonical re-score the result equals the tensor-core pipeline's bit for bit."""
    w = cfg2
    lib, ctx, qs, cap, img = w["lib"], w["ctx"], w["qs"], w["cap"], w["img"]
    full = lib.topk(ctx, qs, cap, K, 0.0, t2i_bank=img, t2i_threshold=0.25)
    sub_classes = [0, 42, 117, 199]
    sub = qs.subset(sub_classes)
    n = 2_000_000                                   # the SIMT kernel is slow: a 2 M-row prefix
    ref = lib.topk(ctx, sub, cap[:n], K, 0.0, t2i_bank=img[:n], t2i_threshold=0.25)
    eps = lib.scan_eps(sub, torch.bfloat16, "simt")
    job = lib.Job(ctx, sub, 640, 0.0 - eps)
    job.scan(cap[:n], t2i_bank=img[:n], t2i_threshold=0.25 - eps, engine="simt")
    sc, rw, cn, tr = job.select()
    assert not job.overflowed()
    job.close()
    o = lib.rescore_walk(ctx, sub, cap[:n], sc, rw, cn, tr, K, 0.0, aux_bank=img[:n], aux_threshold=0.25, eps=eps)
    assert int(o[5].sum()) == 0
    assert torch.equal(o[1], ref[1]) and torch.equal(o[0], ref[0]) and torch.equal(o[2], ref[2]) and torch.equal(o[3], ref[3])
    assert full is not None


# ----------------------------------------------------------------------------------------------- config 3
@pytest.fixture(scope="module")
def cfg3():
    from swat_b200 import _lib, synth
    dev = torch.device("cuda", 0)
    ctx = _lib.Context(0)
    qc, q1000, _ = synth.make_queries(1000, 1, seed=3, dtype=torch.bfloat16)
    cap, _, _ = synth.make_bank(50_000_000, qc, seed=3, device=dev, dtype=torch.bfloat16, chunk=1 << 20, with_images=False)
    torch.cuda.synchronize()
    yield dict(lib=_lib, ctx=ctx, cap=cap, qc=qc, q1000=q1000)
    ctx.close()


@pytest.mark.parametrize("line", ["imagenet-Q1000", "semi-aves-Q400-max"])
def test_config3_50M_lines_vs_restatement(cfg3, line):
    from swat_b200 import synth
    w = cfg3
    lib, ctx, cap = w["lib"], w["ctx"], w["cap"]
    if line == "imagenet-Q1000":
        q, coq, C, red = w["q1000"].float(), None, 1000, "none"
    else:
        # 200 of the bank's classes with two synonyms each, MAX over the pair
        g = torch.Generator().manual_seed(5)
        u = torch.nn.functional.normalize(torch.randn(400, 512, generator=g), dim=-1)
        base = w["qc"][:200].float().repeat_interleave(2, 0)
        q = torch.nn.functional.normalize(base + 0.3 * u, dim=-1).to(torch.bfloat16).float()
        coq, C, red = torch.arange(200, dtype=torch.int32).repeat_interleave(2), 200, "max"
    qs = lib.Queries(ctx, q, coq, C, red)
    g = lib.topk(ctx, qs, cap, K, 0.0)
    r = restate_topk_walk(cap, q.cuda(), K, 0.0, class_of_query=coq, n_classes=C, reduce=red)
    res = compare_walks(g, r, TIE_TOL, what=f"cfg3 {line}")
    _report(f"cfg3 50M {line} esc={ctx.last_timing()['escalations']}", res)
    _invariants(g[0], g[1], g[3], K, 50_000_000, C)
    qs.close()
