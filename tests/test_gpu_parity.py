"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle and the golden vectors
generated from the reference.  Everything here needs a B200 (``-m gpu``)."""
import numpy as np
import pytest
import torch

from oracle import swat_oracle as so
from tests.golden_util import assert_walk_equal, load_bank_case

pytestmark = pytest.mark.gpu

# Parity contract (BASELINE.json north_star): indices identical after tie-breaking by row index,
# score deltas allowed only between near-ties / at the k-th boundary; scores within 1e-3 absolute.
# The tensor-core path accumulates exact bf16 products in fp32 in a different order than the CPU,
# so near-ties are compared with TIE_TOL, far below the 1e-3 the contract allows.
TIE_TOL = 2e-5
SCORE_TOL = 1e-3


@pytest.fixture(scope="module")
def lib():
    from swat_b200 import _lib
    return _lib


@pytest.fixture(scope="module", params=[2, 1], ids=["cta_group2", "cta_group1"])
def ctx(request, lib):
    c = lib.Context(0, cta_group=request.param)
    yield c
    c.close()


@pytest.fixture(scope="module")
def ctx2(lib):
    c = lib.Context(0)
    yield c
    c.close()


def _rand_unit(n, seed, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    x = torch.nn.functional.normalize(torch.randn(n, 512, generator=g), dim=-1)
    return x.to(dtype)


def check_result(scores, rows, counts, o_rows, o_scores, o_counts, S, tol, t2i=None, o_t2i=None, what=""):
    scores, rows, counts = scores.cpu().numpy(), rows.cpu().numpy(), counts.cpu().numpy()
    swaps = 0
    for c in range(rows.shape[0]):
        n = int(counts[c])
        assert n == int(o_counts[c]), f"{what} class {c}: count {n} != oracle {int(o_counts[c])}"
        swaps += assert_walk_equal(rows[c, :n], o_rows[c, :n], lambda r, c=c: S[r, c], tol, boundary_tol=SCORE_TOL, what=f"{what} class {c}")
        assert np.all(rows[c, n:] == -1)
        np.testing.assert_allclose(scores[c, :n], S[rows[c, :n], c], atol=SCORE_TOL)
        np.testing.assert_allclose(scores[c, :n], o_scores[c, :n], atol=SCORE_TOL)
        assert np.all(np.diff(scores[c, :n]) <= 0), f"{what} class {c}: scores not descending"
    return swaps


@pytest.mark.parametrize("n_rows,n_cls", [(1000, 37), (256, 16), (77, 5), (4099, 200)])
def test_dense_scores_match_oracle(lib, ctx, n_rows, n_cls):
    bank = _rand_unit(n_rows, 1, torch.bfloat16)
    q = _rand_unit(n_cls, 2, torch.bfloat16).float()
    ref = so.score_matrix(bank.float().numpy(), q.numpy())
    qs = lib.Queries(ctx, q)
    d_bank = bank.cuda()
    tc = lib.scores_dense(ctx, qs, d_bank, engine="tc").cpu().numpy()
    simt = lib.scores_dense(ctx, qs, d_bank, engine="simt").cpu().numpy()
    np.testing.assert_allclose(simt, ref, atol=2e-6)
    np.testing.assert_allclose(tc, ref, atol=1e-5)
    f32 = lib.scores_dense(ctx, lib.Queries(ctx, _rand_unit(n_cls, 2)), _rand_unit(n_rows, 1).cuda()).cpu().numpy()
    np.testing.assert_allclose(f32, so.score_matrix(_rand_unit(n_rows, 1).numpy(), _rand_unit(n_cls, 2).numpy()), atol=2e-6)


@pytest.mark.parametrize("reduce", ["mean", "max", "min"])
def test_dense_scores_grouped(lib, ctx, reduce):
    sizes = [1, 3, 2, 5, 1, 4, 18, 2, 2, 7]
    coq = np.repeat(np.arange(len(sizes)), sizes).astype(np.int32)
    bank = _rand_unit(777, 3, torch.bfloat16)
    q = _rand_unit(len(coq), 4, torch.bfloat16).float()
    ref = so.score_matrix(bank.float().numpy(), q.numpy(), coq, len(sizes), reduce)
    qs = lib.Queries(ctx, q, coq, len(sizes), reduce)
    for eng in ("tc", "simt"):
        got = lib.scores_dense(ctx, qs, bank.cuda(), engine=eng).cpu().numpy()
        np.testing.assert_allclose(got, ref, atol=1e-5, err_msg=f"{reduce}/{eng}")


@pytest.mark.parametrize("name", ["bank_bf16", "bank_f32"])
def test_golden_reference_parity(lib, ctx, name):
    """The reference's own outputs (tests/golden, made by oracle/gen_golden.py): T2T-rank and
    T2T-rank-T2I-tshd, partitioned (the reference's real use) and unpartitioned."""
    z, meta, cap, img, q = load_bank_case(name)
    k = int(z["k"]); labels = z["labels"]; C = q.shape[0]
    dt = torch.bfloat16 if name == "bank_bf16" else torch.float32
    d_cap = torch.from_numpy(cap).to(dt).cuda(); d_img = torch.from_numpy(img).to(dt).cuda()
    qs = lib.Queries(ctx, torch.from_numpy(q))
    S = so.score_matrix(cap, q)
    row_class = torch.from_numpy(labels.astype(np.int32)).cuda()
    for tag, rc in (("unpart", None), ("part", row_class)):
        for m, t2i in (("t2t", None), ("t2t_t2i", d_img)):
            scores, rows, ti, counts = lib.topk(ctx, qs, d_cap, k, 0.0, t2i_bank=t2i, t2i_threshold=0.25, row_class=rc)
            ref_rows = z[f"{tag}_{m}_rows"]; ref_labels = z[f"{tag}_{m}_labels"]
            rows = rows.cpu().numpy(); counts = counts.cpu().numpy()
            for c in range(C):
                cid = int(z["class_ids"][c])
                exp = ref_rows[ref_labels == cid]
                assert int(counts[c]) == len(exp) == meta["counts"][tag][m][str(cid)], f"{name} {tag} {m} class {cid}"
                assert_walk_equal(rows[c, :counts[c]], exp, lambda r, c=c: S[r, c], TIE_TOL, boundary_tol=SCORE_TOL,
                                  what=f"{name} {tag} {m} class {cid}")
            if ti is not None:
                I = so.score_matrix(img, q)
                ti = ti.cpu().numpy()
                for c in range(C):
                    n = int(counts[c])
                    np.testing.assert_allclose(ti[c, :n], I[rows[c, :n], c], atol=SCORE_TOL)
                    assert np.all(ti[c, :n] >= 0.25 - 1e-6)


def test_tie_probe(lib, ctx2):
    z = np.load("tests/golden/primitives.npz")
    caps, imgs, q = z["probe_caps"], z["probe_imgs"], z["probe_q"]
    qs = lib.Queries(ctx2, torch.from_numpy(q[None]))
    s, r, t, c = lib.topk(ctx2, qs, torch.from_numpy(caps).cuda(), 4, 0.0)
    assert r[0, :int(c[0])].tolist() == [3, 1, 4, 5]
    s, r, t, c = lib.topk(ctx2, qs, torch.from_numpy(caps).cuda(), 4, 0.0, t2i_bank=torch.from_numpy(imgs).cuda())
    assert r[0, :int(c[0])].tolist() == [1, 4, 5, 2]


@pytest.mark.parametrize("n_rows,n_cls,k", [(200_000, 64, 500), (50_001, 200, 100), (3000, 8, 4000)])
def test_topk_random_bank_vs_oracle(lib, ctx, n_rows, n_cls, k):
    """Thresholds, histogram refresh and candidate buffers under load; k > rows-per-class included."""
    from swat_b200 import synth
    qc, queries, coq = synth.make_queries(n_cls, 1, seed=5, dtype=torch.bfloat16)
    cap, img, labels = synth.make_bank(n_rows, qc, seed=5, dtype=torch.bfloat16, rho=0.2, tie_block=300, chunk=1 << 16)
    capf, imgf, qf = cap.float().numpy(), img.float().numpy(), queries.float().numpy()
    S = so.score_matrix(capf, qf)
    qs = lib.Queries(ctx, queries.float())
    o = so.topk_walk(capf, qf, k, 0.0)
    g = lib.topk(ctx, qs, cap.cuda(), k, 0.0)
    check_result(g[0], g[1], g[3], o[0], o[1], o[3], S, TIE_TOL, what="t2t")
    o = so.topk_walk(capf, qf, k, 0.0, t2i_bank=imgf, t2i_threshold=0.25)
    g = lib.topk(ctx, qs, cap.cuda(), k, 0.0, t2i_bank=img.cuda(), t2i_threshold=0.25)
    check_result(g[0], g[1], g[3], o[0], o[1], o[3], S, TIE_TOL, what="t2t+t2i")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "f32"])
@pytest.mark.parametrize("reduce", ["max", "mean", "min"])
def test_topk_synonym_groups(lib, ctx, reduce, dtype):
    """Uneven synonym groups (1..6 queries per class): the host pads the query block so that no class straddles the
    column where the second epilogue warp set starts; fp32 banks run the 4-warp epilogue over those padding columns."""
    from swat_b200 import synth
    sizes = [1 + (i * 7) % 6 for i in range(90)]            # 90 classes, 1..6 synonyms each -> 2 Q blocks at cta_group 1
    qc, queries, coq = synth.make_queries(len(sizes), sizes, seed=9, dtype=torch.bfloat16)
    cap, img, _ = synth.make_bank(30_000, qc, seed=9, dtype=dtype, rho=0.3, tie_block=100, chunk=1 << 15)
    capf, imgf, qf = cap.float().numpy(), img.float().numpy(), queries.float().numpy()
    coq_np = coq.numpy()
    S = so.score_matrix(capf, qf, coq_np, len(sizes), reduce)
    qs = lib.Queries(ctx, queries.float(), coq, len(sizes), reduce)
    o = so.topk_walk(capf, qf, 200, 0.0, t2i_bank=imgf, class_of_query=coq_np, n_classes=len(sizes), reduce=reduce)
    g = lib.topk(ctx, qs, cap.cuda(), 200, 0.0, t2i_bank=img.cuda())
    check_result(g[0], g[1], g[3], o[0], o[1], o[3], S, TIE_TOL, what=f"{reduce} {dtype}")
    dense = lib.scores_dense(ctx, qs, cap.cuda(), engine="tc").cpu().numpy()      # dense mode walks the same columns
    np.testing.assert_allclose(dense, S, atol=1e-5 if dtype == torch.bfloat16 else 2e-2)


@pytest.mark.parametrize("reduce", ["max", "mean"])
def test_full_width_grouped_block(lib, ctx2, reduce):
    """128 classes x 2 synonyms = 256 query columns: the class boundary at column 128 is where the second epilogue
    warp set starts, no padding is needed and the whole set is ONE resident block (it used to be planned as two)."""
    from swat_b200 import synth
    sizes = [2] * 128
    qc, queries, coq = synth.make_queries(len(sizes), sizes, seed=13, dtype=torch.bfloat16)
    cap, _, _ = synth.make_bank(120_000, qc, seed=13, dtype=torch.bfloat16, rho=0.3, tie_block=100, chunk=1 << 15, with_images=False)
    capf, qf, coq_np = cap.float().numpy(), queries.float().numpy(), coq.numpy()
    qs = lib.Queries(ctx2, queries.float(), coq, len(sizes), reduce)
    g = lib.topk(ctx2, qs, cap.cuda(), 150, 0.0)
    S = so.score_matrix(capf, qf, coq_np, len(sizes), reduce)
    o = so.topk_walk(capf, qf, 150, 0.0, class_of_query=coq_np, n_classes=len(sizes), reduce=reduce)
    check_result(g[0], g[1], g[3], o[0], o[1], o[3], S, TIE_TOL, what=f"full-width {reduce}")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "f32"])
def test_dynamic_tile_plan_equals_static(lib, ctx2, dtype):
    """One Q block: pairs claim tiles from a global counter (dyn_tiles, the default) instead of a fixed stride.  Which
    pair scores a tile cannot matter: results are bit-identical to the static plan, on a bank whose tile count is
    not a multiple of the pair count and ends in a ragged tile."""
    from swat_b200 import synth
    qc, queries, _ = synth.make_queries(48, 1, seed=17, dtype=torch.bfloat16)
    cap, img, _ = synth.make_bank(74 * 256 * 9 + 12_345, qc, seed=17, dtype=dtype, rho=0.2, tie_block=200, chunk=1 << 16)
    qs = lib.Queries(ctx2, queries.float())
    capd, imgd = cap.cuda(), img.cuda()
    out = {}
    try:
        for dyn in (1, 0):
            ctx2.set_option("dyn_tiles", dyn)
            out[dyn] = [lib.topk(ctx2, qs, capd, 300, 0.0), lib.topk(ctx2, qs, capd, 300, 0.0, t2i_bank=imgd, t2i_threshold=0.25)]
    finally:
        ctx2.set_option("dyn_tiles", 1)
    for a, b in zip(out[1], out[0]):
        assert torch.equal(a[3], b[3]) and torch.equal(a[1], b[1]) and torch.equal(a[0], b[0])
    capf, qf = cap.float().numpy(), queries.float().numpy()
    o = so.topk_walk(capf, qf, 300, 0.0)
    g = out[1][0]
    check_result(g[0], g[1], g[3], o[0], o[1], o[3], capf @ qf.T, TIE_TOL, what=f"dynamic tiles {dtype}")


def test_lockstep_window_is_result_neutral(lib, ctx2):
    """Four Q blocks: held in step (lock_window > 0, what the automatic mode switches on for a power-capped GPU) the
    pairs of a tile range read the bank once; which pair is ahead cannot matter to the result."""
    from swat_b200 import synth
    qc, queries, _ = synth.make_queries(1000, 1, seed=23, dtype=torch.bfloat16)
    cap, _, _ = synth.make_bank(400_000, qc, seed=23, dtype=torch.bfloat16, rho=0.3, tie_block=300, chunk=1 << 16, with_images=False)
    qs = lib.Queries(ctx2, queries.float())
    capd = cap.cuda()
    out = {}
    try:
        for lw in (0, 4, -1):
            ctx2.set_option("lock_window", lw)
            out[lw] = lib.topk(ctx2, qs, capd, 100, 0.0)
    finally:
        ctx2.set_option("lock_window", -1)
    for lw in (4, -1):
        assert torch.equal(out[lw][1], out[0][1]) and torch.equal(out[lw][0], out[0][0]) and torch.equal(out[lw][3], out[0][3])
    capf, qf = cap.float().numpy(), queries.float().numpy()
    o = so.topk_walk(capf, qf, 100, 0.0)
    check_result(out[4][0], out[4][1], out[4][3], o[0], o[1], o[3], capf @ qf.T, TIE_TOL, what="lockstep")


@pytest.mark.parametrize("n_cls,syn", [(700, False), (230, True)])
def test_unit_plan_many_query_blocks(lib, ctx, n_cls, syn):
    """More queries than one resident block holds, block count not dividing the CTA pairs: the scan runs as
    several launches of balanced (query block x tile range) units.  Must equal the oracle and, bit for bit,
    the single-launch schedule."""
    from swat_b200 import synth
    sizes = [1 + (i * 5) % 7 for i in range(n_cls)] if syn else 1          # 230 classes -> ~900 queries
    qc, queries, coq = synth.make_queries(n_cls, sizes, seed=21, dtype=torch.bfloat16)
    cap, _, _ = synth.make_bank(360_000, qc, seed=21, dtype=torch.bfloat16, rho=0.3, tie_block=300, chunk=1 << 16, with_images=False)
    capf, qf = cap.float().numpy(), queries.float().numpy()
    red = "max" if syn else "none"
    coq_np = coq.numpy()
    qs = lib.Queries(ctx, queries.float(), coq, n_cls, red)
    capd = cap.cuda()
    l0 = ctx.launch_count
    g = lib.topk(ctx, qs, capd, 100, 0.0)
    n_launch = ctx.launch_count - l0
    ctx.set_option("unit_plan", 0)
    try:
        l0 = ctx.launch_count
        h = lib.topk(ctx, qs, capd, 100, 0.0)
        assert ctx.launch_count - l0 < n_launch, "unit plan did not engage"
    finally:
        ctx.set_option("unit_plan", 1)
    assert torch.equal(g[1], h[1]) and torch.equal(g[0], h[0]) and torch.equal(g[3], h[3])
    S = so.score_matrix(capf, qf, coq_np, n_cls, red)
    o = so.topk_walk(capf, qf, 100, 0.0, class_of_query=coq_np, n_classes=n_cls, reduce=red)
    check_result(g[0], g[1], g[3], o[0], o[1], o[3], S, TIE_TOL, what="unit plan")


def test_exclusion_bitmap_and_threshold(lib, ctx2):
    bank = _rand_unit(5000, 11, torch.bfloat16)
    q = _rand_unit(12, 12, torch.bfloat16)
    S = so.score_matrix(bank.float().numpy(), q.float().numpy())
    ex = np.zeros(5000, dtype=bool); ex[::3] = True
    bits = np.packbits(ex, bitorder="little")
    bits = np.concatenate([bits, np.zeros((-len(bits)) % 4, np.uint8)]).view(np.int32)
    qs = lib.Queries(ctx2, q.float())
    for thr in (0.0, 0.05, -1.0):
        o = so.topk_walk(bank.float().numpy(), q.float().numpy(), 64, thr, exclude=ex)
        g = lib.topk(ctx2, qs, bank.cuda(), 64, thr, exclude=torch.from_numpy(bits).cuda())
        check_result(g[0], g[1], g[3], o[0], o[1], o[3], S, TIE_TOL, what=f"thr {thr}")
        assert not np.any(ex[g[1].cpu().numpy()[g[1].cpu().numpy() >= 0]])


def test_overflow_retry_and_t2i_escalation(lib):
    """Tiny survivor lists and class candidate buffers force both overflow retries; a T2I predicate
    almost nothing passes forces over-fetch escalation and finally the exact in-pass predicate."""
    ctx = lib.Context(0, cand_cap=48, list_entries=20_000, overfetch=64)
    bank = _rand_unit(60_000, 21, torch.bfloat16)
    img = _rand_unit(60_000, 22, torch.bfloat16)
    q = _rand_unit(24, 23, torch.bfloat16)
    img[1000:60_000:997] = q[3]                                  # ~60 rows pass T2I for class 3 only
    bf, imf, qf = bank.float().numpy(), img.float().numpy(), q.float().numpy()
    S = so.score_matrix(bf, qf)
    qs = lib.Queries(ctx, q.float())
    o = so.topk_walk(bf, qf, 40, 0.0)
    g = lib.topk(ctx, qs, bank.cuda(), 40, 0.0)
    check_result(g[0], g[1], g[3], o[0], o[1], o[3], S, TIE_TOL, what="overflow")
    assert ctx.last_timing()["escalations"] >= 1
    o = so.topk_walk(bf, qf, 40, 0.0, t2i_bank=imf, t2i_threshold=0.25)
    g = lib.topk(ctx, qs, bank.cuda(), 40, 0.0, t2i_bank=img.cuda(), t2i_threshold=0.25)
    check_result(g[0], g[1], g[3], o[0], o[1], o[3], S, TIE_TOL, what="escalation")
    assert int(o[3][3]) > 0 and int(o[3].sum()) == int(o[3][3])
    ctx.close()


def test_bank_swap_escalation(lib):
    """Classes with fewer than k rows passing T2I: after the over-fetch ladder the pipeline enumerates the passers
    from the IMAGE bank (one tensor-core pass + exact re-scores) instead of the fp32 two-bank scan.  Class 3 has ~60
    passers (resolved by the swap pass), class 5 has 5000 passers with noise captions (more than the swap pass can
    enumerate: falls through to the in-pass predicate), every other class has none."""
    bank = _rand_unit(60_000, 71, torch.bfloat16)
    img = _rand_unit(60_000, 72, torch.bfloat16)
    q = _rand_unit(24, 73, torch.bfloat16)
    img[1000:60_000:997] = q[3]
    img[7:60_000:12] = q[5]
    bf, imf, qf = bank.float().numpy(), img.float().numpy(), q.float().numpy()
    S = so.score_matrix(bf, qf)
    o = so.topk_walk(bf, qf, 400, 0.0, t2i_bank=imf, t2i_threshold=0.25)
    assert 0 < int(o[3][3]) < 400 and int(o[3][5]) == 400 and int(o[3].sum()) == int(o[3][3]) + 400
    scans = {}
    for swap in (1, 0):
        ctx = lib.Context(0, swap_pass=swap)
        qs = lib.Queries(ctx, q.float())
        for rep in range(2):          # second call: the remembered per-class depths send the short classes straight to the swap pass
            g = lib.topk(ctx, qs, bank.cuda(), 400, 0.0, t2i_bank=img.cuda(), t2i_threshold=0.25)
            check_result(g[0], g[1], g[3], o[0], o[1], o[3], S, TIE_TOL, what=f"swap={swap} rep={rep}")
            I = so.score_matrix(imf, qf)
            for c in (3, 5):
                n = int(g[3][c]); r = g[1][c, :n].cpu().numpy()
                np.testing.assert_allclose(g[2][c, :n].cpu().numpy(), I[r, c], atol=SCORE_TOL)
                assert np.all(I[r, c] >= 0.25 - 1e-6)
        scans[swap] = ctx.last_timing()["scan_launches"]
        qs.close(); ctx.close()
    assert scans[1] >= 2 and scans[0] >= 2
    # partitioned data (a row is eligible for its own class only): 3 classes x 20 000 rows, class 1 has 50 passers
    lab = np.repeat(np.arange(3), 20_000).astype(np.int32)
    img2 = _rand_unit(60_000, 74, torch.bfloat16)
    q2 = q[:3].clone()
    img2[20_000:40_000:400] = q2[1]
    bf2, imf2, qf2 = bank.float().numpy(), img2.float().numpy(), q2.float().numpy()
    o = so.topk_walk(bf2, qf2, 100, 0.0, t2i_bank=imf2, t2i_threshold=0.25, row_labels=lab)
    assert int(o[3][0]) == 0 and int(o[3][2]) == 0 and 0 < int(o[3][1]) < 100
    ctx = lib.Context(0)
    qs = lib.Queries(ctx, q2.float())
    g = lib.topk(ctx, qs, bank.cuda(), 100, 0.0, t2i_bank=img2.cuda(), t2i_threshold=0.25, row_class=torch.from_numpy(lab).cuda())
    check_result(g[0], g[1], g[3], o[0], o[1], o[3], so.score_matrix(bf2, qf2), TIE_TOL, what="partitioned swap")
    qs.close(); ctx.close()


def test_partitioned_escalation_mix(lib):
    """Partitioned data (a row is eligible for its own class only) where the classes need DIFFERENT escalation paths in
    one call: class 0 fills in the first pass, class 1 has ~50 T2I passers (bank-swap pass on a sub-query set with
    renumbered row classes), class 2 has 6000 passers whose captions rank low (more than the swap pass can enumerate:
    two-pass in-pass predicate), class 3 has none.  Round 1 spliced rows of the wrong class here (sub-query sets kept the
    original row_class array)."""
    n_per, C = 20_000, 4
    lab = np.repeat(np.arange(C), n_per).astype(np.int32)
    q = _rand_unit(C, 83, torch.bfloat16)
    bank = _rand_unit(C * n_per, 81, torch.bfloat16)
    img = _rand_unit(C * n_per, 82, torch.bfloat16)
    unit = lambda x: torch.nn.functional.normalize(x.float(), dim=-1).to(torch.bfloat16)
    img[0:n_per:2] = q[0]                                              # class 0: every second row passes, plenty in its T2T head
    img[n_per:2 * n_per:400] = q[1]                                    # class 1: 50 passers
    rows2 = torch.arange(2 * n_per, 3 * n_per, 3)[:6000]               # class 2: 6000 passers ...
    img[rows2] = q[2]
    bank[rows2] = unit(-0.05 * q[2].float()[None, :] + bank[rows2].float())  # ... whose captions rank BELOW most of the class: < 300 of them in its T2T top-4096
    bf, imf, qf = bank.float().numpy(), img.float().numpy(), q.float().numpy()
    o = so.topk_walk(bf, qf, 300, 0.0, t2i_bank=imf, t2i_threshold=0.25, row_labels=lab)
    assert int(o[3][0]) == 300 and 0 < int(o[3][1]) < 300 and int(o[3][2]) == 300 and int(o[3][3]) == 0
    rank2 = np.argsort(np.argsort(-so.score_matrix(bf[2 * n_per:3 * n_per], qf)[:, 2]))     # T2T rank inside class 2
    assert rank2[o[0][2][299] - 2 * n_per] > 4096                       # its walk really ends beyond the widest over-fetch
    S = so.score_matrix(bf, qf)
    for swap in (1, 0):
        ctx = lib.Context(0, swap_pass=swap)
        qs = lib.Queries(ctx, q.float())
        for rep in range(2):
            g = lib.topk(ctx, qs, bank.cuda(), 300, 0.0, t2i_bank=img.cuda(), t2i_threshold=0.25, row_class=torch.from_numpy(lab).cuda())
            check_result(g[0], g[1], g[3], o[0], o[1], o[3], S, TIE_TOL, what=f"partitioned mix swap={swap} rep={rep}")
            rows = g[1].cpu().numpy()
            for c in range(C):
                r = rows[c][rows[c] >= 0]
                assert np.all(lab[r] == c), f"class {c} received rows of another class"
        assert ctx.last_timing()["scan_launches"] >= 2
        qs.close(); ctx.close()


def test_host_pipeline_equals_resident(lib, ctx2):
    """Scores are canonical (one fixed-order fp32 dot per returned row), so the resident pipeline, the host pipeline
    with pinned banks (candidates' rows read in place over PCIe) and the host pipeline with pageable banks (host-side
    gather) agree bit for bit, whichever escalation path each of them took."""
    from swat_b200 import synth
    qc, queries, _ = synth.make_queries(40, 1, seed=31, dtype=torch.bfloat16)
    cap, img, labels = synth.make_bank(70_000, qc, seed=31, dtype=torch.bfloat16, rho=0.2, tie_block=200, chunk=1 << 16)
    qs = lib.Queries(ctx2, queries.float())
    ctx2.set_option("host_chunk_rows", 8192)
    capf, imgf, qf = cap.float().numpy(), img.float().numpy(), queries.float().numpy()
    S = so.score_matrix(capf, qf)
    for t2i_dev, t2i_pin, t2i_page in ((None, None, None), (img.cuda(), img.pin_memory(), img.clone())):
        r = lib.topk(ctx2, qs, cap.cuda(), 300, 0.0, t2i_bank=t2i_dev, row_offset=1000)
        h = lib.topk_host(ctx2, qs, cap.pin_memory(), 300, 0.0, t2i_bank=t2i_pin, row_offset=1000)
        ctx2.set_option("zero_copy", 0)
        g = lib.topk_host(ctx2, qs, cap.clone(), 300, 0.0, t2i_bank=t2i_page, row_offset=1000)
        ctx2.set_option("zero_copy", 1)
        for x in (h, g):
            assert torch.equal(r[3].cpu(), x[3]) and torch.equal(r[1].cpu(), x[1]) and torch.equal(r[0].cpu(), x[0])
            if t2i_dev is not None:
                assert torch.equal(r[2].cpu(), x[2])
        o = so.topk_walk(capf, qf, 300, 0.0, t2i_bank=None if t2i_dev is None else imgf, t2i_threshold=0.25)
        rows = torch.where(r[1] >= 0, r[1] - 1000, r[1])
        check_result(r[0], rows, r[3], o[0], o[1], o[3], S, TIE_TOL, what="host vs resident")
    ctx2.set_option("host_chunk_rows", 1 << 18)


def test_merge_and_shard_invariance(lib, ctx2):
    """Row-sharded scan + one gather + merge equals the single-shard result for any shard count
    (SURVEY 8e), bit for bit: every shard walks its own candidates and ships at most k accepted rows per class
    plus its limit (swat_b200/dist.py); the merge keeps the k best of the union."""
    from swat_b200 import dist, synth
    qc, queries, _ = synth.make_queries(30, 1, seed=41, dtype=torch.bfloat16)
    cap, img, _ = synth.make_bank(90_000, qc, seed=41, dtype=torch.bfloat16, rho=0.2, tie_block=500, chunk=1 << 16)
    qs = lib.Queries(ctx2, queries.float())
    d_cap, d_img = cap.cuda(), img.cuda()
    full_t2t = lib.topk(ctx2, qs, d_cap, 250, 0.0)
    full = lib.topk(ctx2, qs, d_cap, 250, 0.0, t2i_bank=d_img)
    for G in (1, 2, 3, 8):
        bounds = [dist.shard_range(90_000, r, G) for r in range(G)]
        assert bounds[0][0] == 0 and bounds[-1][1] == 90_000 and all(a[1] == b[0] for a, b in zip(bounds[:-1], bounds[1:]))
        # T2T only
        parts = [dist.local_walk(ctx2, qs, d_cap[a:b], 250, 1024, 0.0, None, row_offset=a) for a, b in bounds]     # deep enough to see past the block of 500 ties
        buf = torch.cat([dist.pack(*p) for p in parts])
        s, r, t, c, lim = dist.unpack(buf, G, 30, 250, False)
        ms, mr, mt, mc, inc = lib.merge_topk(ctx2, s, r, c, limit=lim, k_out=250)
        assert int(inc.sum()) == 0, f"T2T G={G}"
        assert torch.equal(mr, full_t2t[1]) and torch.equal(mc, full_t2t[3]) and torch.equal(ms, full_t2t[0]), f"T2T G={G}"
        # T2I walk per shard, merged
        parts = [dist.local_walk(ctx2, qs, d_cap[a:b], 250, 1024, 0.0, d_img[a:b], 0.25, row_offset=a) for a, b in bounds]
        buf = torch.cat([dist.pack(*p) for p in parts])
        s, r, t, c, lim = dist.unpack(buf, G, 30, 250, True)
        ms, mr, mt, mc, inc = lib.merge_topk(ctx2, s, r, c, aux=t, limit=lim, k_out=250)
        assert int(inc.sum()) == 0, f"G={G}"
        assert torch.equal(mr, full[1]) and torch.equal(mc, full[3]), f"T2I G={G}"
        assert torch.equal(ms, full[0]) and torch.equal(mt, full[2])
        # the exchange path proper: kernels write into packed buffers (no overflow check before the exchange), the
        # merge reads the gathered buffer in place through the per-shard stride
        pks = []
        for a, b in bounds:
            pk = dist.PackedResults(30, 250, True, d_cap.device)
            dist.local_walk(ctx2, qs, d_cap[a:b], 250, 1024, 0.0, d_img[a:b], 0.25, row_offset=a, packed=pk, check=False)
            pks.append(pk)
        gathered = torch.cat([pk.buf for pk in pks])
        ps, pr, pt, pc, pinc = dist.merge_packed(gathered, pks[0].lay, G, 250, ctx=ctx2)
        assert torch.equal(pr, full[1]) and torch.equal(pc, full[3]) and torch.equal(ps, full[0]) and torch.equal(pt, full[2])
        assert int(pinc.sum()) == 0 and dist.unpack_flags(gathered, G, 30, 250, True).tolist() == [0] * G
    # a frontier violation must be reported: k_fetch too small to find 250 passing rows
    parts = [dist.local_walk(ctx2, qs, d_cap[a:b], 250, 256, 0.0, d_img[a:b], 0.25, row_offset=a) for a, b in bounds[:2]]
    s, r, t, c, lim = dist.unpack(torch.cat([dist.pack(*p) for p in parts]), 2, 30, 250, True)
    assert bool((lim > float("-inf")).any())
    inc = lib.merge_topk(ctx2, s, r, c, aux=t, limit=lim, k_out=250)[4]
    assert int(inc.sum()) > 0
    # single-process world: the whole sharded pipeline, including partitioned data with an exclusion bitmap
    res = dist.topk_sharded(ctx2, qs, d_cap, 250, 0.0, t2i_bank=d_img, world=1)
    assert torch.equal(res[1], full[1]) and torch.equal(res[0], full[0])
    rc = torch.randint(0, 30, (90_000,), generator=torch.Generator().manual_seed(3), dtype=torch.int32).cuda()
    ex = torch.randint(-2 ** 31, 2 ** 31 - 1, ((90_000 + 31) // 32,), generator=torch.Generator().manual_seed(4), dtype=torch.int64).to(torch.int32).cuda()
    one = lib.topk(ctx2, qs, d_cap, 60, 0.0, t2i_bank=d_img, row_class=rc, exclude=ex)
    res = dist.topk_sharded(ctx2, qs, d_cap, 60, 0.0, t2i_bank=d_img, world=1, row_class=rc, exclude=ex)
    assert torch.equal(res[1], one[1]) and torch.equal(res[0], one[0]) and torch.equal(res[3], one[3])


def test_streaming_job_matches_single_view(lib, ctx2):
    bank = _rand_unit(40_000, 51, torch.bfloat16).cuda()
    q = _rand_unit(20, 52, torch.bfloat16)
    qs = lib.Queries(ctx2, q.float())
    job = lib.Job(ctx2, qs, 128, 0.0)
    job.scan(bank)
    a = job.select(); assert not job.overflowed()
    job.reset()
    for s0 in range(0, 40_000, 7777):
        job.scan(bank[s0:s0 + 7777], row_base=s0)
    b = job.select(); assert not job.overflowed()
    for x, y in zip(a[:3], b[:3]):
        assert torch.equal(x, y)
    job.close()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_ragged_and_tiny_banks(lib, ctx, dtype):
    """Banks smaller than, equal to and just past a tile (128 / 256 rows), a single row, a single query, k larger than
    the bank, everything excluded, a threshold nothing reaches -- resident and host pipelines, both engines."""
    for n_rows, n_cls in ((1, 1), (1, 17), (127, 3), (128, 3), (129, 17), (255, 1), (256, 5), (257, 17), (1000, 200)):
        bank = _rand_unit(n_rows, 100 + n_rows, dtype)
        img = _rand_unit(n_rows, 200 + n_rows, dtype)
        q = _rand_unit(n_cls, 300 + n_cls, torch.bfloat16).float()
        bf, imf, qf = bank.float().numpy(), img.float().numpy(), q.numpy()
        S = so.score_matrix(bf, qf)
        qs = lib.Queries(ctx, q)
        for k in (1, 40, 300):
            o = so.topk_walk(bf, qf, k, 0.0)
            g = lib.topk(ctx, qs, bank.cuda(), k, 0.0)
            check_result(g[0], g[1], g[3], o[0], o[1], o[3], S, TIE_TOL, what=f"N={n_rows} C={n_cls} k={k}")
            o = so.topk_walk(bf, qf, k, 0.0, t2i_bank=imf, t2i_threshold=0.0)
            g = lib.topk(ctx, qs, bank.cuda(), k, 0.0, t2i_bank=img.cuda(), t2i_threshold=0.0)
            check_result(g[0], g[1], g[3], o[0], o[1], o[3], S, TIE_TOL, what=f"T2I N={n_rows} C={n_cls} k={k}")
            h = lib.topk_host(ctx, qs, bank, k, 0.0, t2i_bank=img, t2i_threshold=0.0)
            assert torch.equal(h[1], g[1].cpu()) and torch.equal(h[3], g[3].cpu())
        # nothing can be accepted: threshold above every score, or every row excluded
        g = lib.topk(ctx, qs, bank.cuda(), 10, 2.0)
        assert int(g[3].sum()) == 0 and bool((g[1] == -1).all())
        bits = torch.full(((n_rows + 31) // 32,), -1, dtype=torch.int32)
        g = lib.topk(ctx, qs, bank.cuda(), 10, 0.0, exclude=bits.cuda())
        assert int(g[3].sum()) == 0 and bool((g[1] == -1).all())
        qs.close()


def test_bad_arguments_raise(lib, ctx2):
    q = _rand_unit(4, 1)
    with pytest.raises(lib.SwatError):
        lib.Queries(ctx2, q, np.array([0, 0, 1, 1], np.int32), 2, "none")      # NONE needs one query per class
    with pytest.raises(lib.SwatError):
        lib.Queries(ctx2, q, np.array([1, 0, 1, 0], np.int32), 2, "max")       # not grouped
    qs = lib.Queries(ctx2, q)
    with pytest.raises(lib.SwatError):
        lib.topk(ctx2, qs, _rand_unit(64, 2).cuda(), 100000)                   # k beyond the supported maximum
    with pytest.raises((ValueError, TypeError)):
        lib.topk(ctx2, qs, torch.zeros(64, 256).cuda(), 10)
    # fp32 banks are scanned as bf16-rounded rows under an error bound that assumes L2-normalised rows: un-normalised
    # features must fail loudly (the re-score checks the bound on every candidate), never return unproven rows
    big = (_rand_unit(4096, 5) * 40.0).cuda()
    with pytest.raises(lib.SwatError, match="L2-normalised"):
        lib.topk(ctx2, qs, big, 10)
    ok = lib.topk(ctx2, qs, _rand_unit(4096, 5).cuda(), 10)
    assert int(ok[3].min()) == 10


@pytest.mark.parametrize("reduce", ["none", "max"])
def test_threshold_bootstrap_is_exact(lib, reduce):
    """A fresh job seeds its thresholds from a dense prefix (swat_job_scan bootstrap).  The result must be
    identical, bit for bit, to the un-bootstrapped scan, and match the oracle."""
    from swat_b200 import synth
    sizes = [1] * 48 if reduce == "none" else [1 + i % 4 for i in range(48)]
    qc, queries, coq = synth.make_queries(48, sizes, seed=61, dtype=torch.bfloat16)
    cap, img, _ = synth.make_bank(120_000, qc, seed=61, dtype=torch.bfloat16, rho=0.2, tie_block=400, chunk=1 << 16)
    res = {}
    for boot in (0, 8192):
        ctx = lib.Context(0, bootstrap_rows=boot)
        qs = lib.Queries(ctx, queries.float(), coq, 48, reduce)
        before = ctx.launch_count
        res[boot] = [x.cpu() if x is not None else None for x in lib.topk(ctx, qs, cap.cuda(), 300, 0.0, t2i_bank=img.cuda())]
        res[boot].append(ctx.launch_count - before)
        qs.close(); ctx.close()
    assert res[8192][4] > res[0][4]                      # the bootstrap really ran (two extra launches per scan)
    for a, b in zip(res[0][:4], res[8192][:4]):
        assert torch.equal(a, b)
    capf, imgf, qf = cap.float().numpy(), img.float().numpy(), queries.float().numpy()
    S = so.score_matrix(capf, qf, coq.numpy(), 48, reduce)
    o = so.topk_walk(capf, qf, 300, 0.0, t2i_bank=imgf, class_of_query=coq.numpy(), n_classes=48, reduce=reduce)
    check_result(res[8192][0], res[8192][1], res[8192][3], o[0], o[1], o[3], S, TIE_TOL, what=f"bootstrap {reduce}")


def test_per_class_depth_job(lib, ctx2):
    """swat_job_set_class_depth: each class keeps its own number of best rows; the deep classes' lists are
    prefixes-compatible with a uniform deep job."""
    bank = _rand_unit(50_000, 71, torch.bfloat16).cuda()
    q = _rand_unit(12, 72, torch.bfloat16)
    qs = lib.Queries(ctx2, q.float())
    depth = torch.tensor([64, 512, 64, 64, 300, 64, 64, 64, 64, 64, 512, 1], dtype=torch.int32)
    job = lib.Job(ctx2, qs, 512, 0.0)
    job.set_class_depth(depth)
    job.scan(bank)
    s, r, c, tr = job.select(); assert not job.overflowed()
    job.reset(); job.set_class_depth(None); job.scan(bank)
    s2, r2, c2, _ = job.select(); assert not job.overflowed()
    job.close()
    assert c.tolist() == depth.tolist() and int(c2.min()) == 512
    for i, d in enumerate(depth.tolist()):
        assert torch.equal(r[i, :d], r2[i, :d]) and torch.equal(s[i, :d], s2[i, :d])
        assert bool((r[i, d:] == -1).all()) and int(tr[i]) == 1
    with pytest.raises(lib.SwatError):
        j = lib.Job(ctx2, qs, 100, 0.0); j.set_class_depth(torch.full((12,), 101, dtype=torch.int32))
