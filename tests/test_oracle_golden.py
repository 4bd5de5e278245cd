"""Pin the oracle (oracle/swat_oracle.py) to outputs of the reference's own functions.

The fixtures under tests/golden/ were produced by oracle/gen_golden.py, which imports
/root/reference/retrieval/sample_retrieval.py and runs it unmodified on CPU."""
import hashlib

import numpy as np
import pytest

from oracle import swat_oracle as so
from tests.golden_util import GOLDEN, assert_walk_equal, load_bank_case, make_paths


def test_primitives_match_reference():
    z = np.load(f"{GOLDEN}/primitives.npz")
    X, P, F = z["X"], z["P"], z["F"]
    np.testing.assert_allclose(so.similarity(P, X), z["t2t_R3"], atol=2e-6)
    np.testing.assert_allclose(so.similarity(P, X), z["t2i_R3"], atol=2e-6)
    np.testing.assert_allclose(so.similarity(P[:1], X), z["t2t_R1"], atol=2e-6)
    one = so.similarity(P[:1], X[:1])
    assert isinstance(one, list) and len(one) == 1
    np.testing.assert_allclose(one, z["t2t_single_row"], atol=2e-6)
    for mode in ("min", "max", "mean"):
        np.testing.assert_allclose(so.similarity_p2p(F, X, mode), z[f"p2p_{mode}"], atol=2e-6)
    with pytest.raises(ValueError):
        so.similarity_p2p(F, X, "median")


def test_tie_probe_matches_reference():
    """9 rows, rows 2/6/7 identical T2T; row 3 has the best T2T but fails T2I."""
    z = np.load(f"{GOLDEN}/primitives.npz")
    caps, imgs, q = z["probe_caps"], z["probe_imgs"], z["probe_q"]
    assert z["probe_t2t_rows"].tolist() == [3, 1, 4, 5]
    assert z["probe_t2t_t2i_rows"].tolist() == [1, 4, 5, 2]
    r, s, _, c = so.topk_walk(caps, q[None], 4, 0.0)
    assert r[0, :c[0]].tolist() == [3, 1, 4, 5]
    r, s, i, c = so.topk_walk(caps, q[None], 4, 0.0, t2i_bank=imgs, t2i_threshold=0.25)
    assert r[0, :c[0]].tolist() == [1, 4, 5, 2]
    feats = {"0": {"file_paths": [f"/r/0/{i}.jpg" for i in range(9)], "feats": imgs, "caption_feats": caps}}
    ms, nd, _ = so.verbatim_t2t_ranked_sampler({"0": {"mean": q}}, 4, 0.0, feats)
    assert ms["row_list"][0].tolist() == [3, 1, 4, 5]
    ms, nd, _ = so.verbatim_t2t_ranked_t2i_tshd_sampler({"0": {"mean": q}}, 4, 0.0, feats)
    assert ms["row_list"][0].tolist() == [1, 4, 5, 2]


@pytest.mark.parametrize("name", ["bank_bf16", "bank_f32"])
def test_samplers_match_reference(name):
    z, meta, cap, img, q = load_bank_case(name)
    class_ids = z["class_ids"]; labels = z["labels"]; k = int(z["k"])
    C = len(class_ids)
    # bf16-valued inputs: exact list equality.  fp32 inputs: the reference's own BLAS returns
    # 1-ulp-different scores for bit-identical rows, so ties are compared within 2e-6.
    tol = 0.0 if name == "bank_bf16" else 2e-6
    S_cap, S_img = so.score_matrix(cap, q), so.score_matrix(img, q)

    def same(got, ref_key):
        S = S_img if ref_key.endswith("_t2i_rows") and "t2t" not in ref_key else S_cap
        ref = z[ref_key]; labs = z[ref_key.replace("_rows", "_labels")]
        assert len(got) == len(ref)
        pos = 0
        for cid in np.unique(labs):
            n = int((labs == cid).sum()); c = int(np.nonzero(class_ids == cid)[0][0])
            assert_walk_equal(got[pos:pos + n], ref[pos:pos + n], lambda r: S[r, c], tol, what=f"{ref_key} class {cid}")
            pos += n
    np.testing.assert_allclose(so.similarity(q[:1], cap[:64]), z["sim64"], atol=2e-6)
    np.testing.assert_allclose(so.similarity(q[:1], img[:64]), z["t2i64"], atol=2e-6)
    paths, cmap = make_paths(labels, class_ids)
    raw = {"caption_features": cap, "image_features": img, "labels": class_ids[labels], "filepath": paths}
    prompts = {str(class_ids[c]): {"mean": q[c]} for c in range(C)}
    # --- regrouper: key order and per-class row order
    feats = so.transform_extracted_fea(raw)
    assert list(feats.keys()) == meta["regroup_keys"]
    for i, kk in enumerate(feats.keys()):
        assert feats[kk]["row_ids"].tolist() == z[f"regroup_rows_{i}"].tolist()
    # --- partitioned, verbatim port incl. the diagnostic text files
    t2i_rank = lambda *a, **kw: so.verbatim_t2t_ranked_sampler(*a, rank_on_images=True, **kw)      # t2i_ranked_sampler :1195-1243
    for m, fn in (("t2t", so.verbatim_t2t_ranked_sampler), ("t2t_t2i", so.verbatim_t2t_ranked_t2i_tshd_sampler), ("t2i", t2i_rank)):
        ms, nd, diag = fn(prompts, k, 0.0, feats, caption_map=cmap)
        assert nd == meta["counts"]["part"][m]
        got = np.concatenate([feats[str(int(l[0]))]["row_ids"][r] for r, l in zip(ms["row_list"], ms["label_list"])])
        same(got, f"part_{m}_rows")
        assert np.concatenate(ms["label_list"]).tolist() == z[f"part_{m}_labels"].tolist()
        if not tol:
            np.testing.assert_allclose(np.concatenate(ms["feature_list"]).astype(np.float64).sum(1), z[f"part_{m}_featsum"], atol=1e-5)
        d = meta["diag"]["part"][m]
        assert diag["sampled_list"][:3] == d["sampled_head"] or tol
        if tol or m == "t2i":
            # diagnostic files list rows in walk order with round(score, 4): byte-comparable only without
            # ties and when no score sits within a BLAS rounding difference (1e-7) of a 4th-decimal boundary
            continue
        assert hashlib.sha256("\n".join(diag["sampled_list"]).encode()).hexdigest() == d["sampled_sha"]
        assert hashlib.sha256("\n".join(diag["filtered_list"]).encode()).hexdigest() == d["filtered_sha"]
    # --- partitioned, vectorised restatement: same rows
    dense = labels.astype(np.int64)
    for m, bank, t2i in (("t2t", cap, None), ("t2t_t2i", cap, img), ("t2i", img, None)):
        rows, sc, ti, cnt = so.topk_walk(bank, q, k, 0.0, t2i_bank=t2i, row_labels=dense)
        got = np.concatenate([rows[c, :cnt[c]] for c in range(C)])
        same(got, f"part_{m}_rows")
        assert {str(class_ids[c]): int(cnt[c]) for c in range(C)} == meta["counts"]["part"][m]
    # --- unpartitioned (every class scans the whole bank): vectorised and verbatim port
    for m, bank, t2i in (("t2t", cap, None), ("t2t_t2i", cap, img), ("t2i", img, None)):
        rows, sc, ti, cnt = so.topk_walk(bank, q, k, 0.0, t2i_bank=t2i, row_chunk=500)
        got = np.concatenate([rows[c, :cnt[c]] for c in range(C)])
        same(got, f"unpart_{m}_rows")
        assert {str(class_ids[c]): int(cnt[c]) for c in range(C)} == meta["counts"]["unpart"][m]
    feats_u = {str(class_ids[c]): {"file_paths": paths, "feats": img, "caption_feats": cap} for c in range(C)}
    ms, nd, _ = so.verbatim_t2t_ranked_t2i_tshd_sampler(prompts, k, 0.0, feats_u)
    same(np.concatenate(ms["row_list"]), "unpart_t2t_t2i_rows")


def test_split_line_format():
    lines = so.format_split_lines([["/a/1/5.jpg", "/a/1/6.jpg"], ["/a/2/7.jpg"]], [np.array([1, 1]), np.array([2])])
    assert lines == ["/a/1/5.jpg 1 0\n", "/a/1/6.jpg 1 0\n", "/a/2/7.jpg 2 0\n"]


def test_merge_is_shard_count_invariant():
    """Top-k under (score desc, row asc) is associative: merging shard-local results equals the
    single-shard result for any shard count.  Scores are computed once so BLAS
    blocking cannot perturb ties."""
    rng = np.random.default_rng(0)
    S = rng.standard_normal((3000, 3)).astype(np.float32)
    S[100:140] = S[7]                            # ties across shard boundaries
    k = 50
    full_rows = [so.select_walk(S[:, c], k, -1.0) for c in range(3)]
    for G in (1, 2, 3, 8):
        bounds = np.linspace(0, 3000, G + 1).astype(int)
        rows = np.full((G, 3, k), -1, dtype=np.int64); sc = np.zeros((G, 3, k), np.float32); cnt = np.zeros((G, 3), np.int32)
        for g, (a, b) in enumerate(zip(bounds[:-1], bounds[1:])):
            for c in range(3):
                sel = so.select_walk(S[a:b, c], k, -1.0)
                rows[g, c, :sel.size] = sel + a; sc[g, c, :sel.size] = S[a:b, c][sel]; cnt[g, c] = sel.size
        r, s_, c_ = so.merge_topk(rows, sc, cnt, k)
        for c in range(3):
            assert r[c].tolist() == full_rows[c].tolist()
            np.testing.assert_array_equal(s_[c], S[full_rows[c], c])


def _fewshot_case():
    import json
    from tests.golden_util import bf16_bits_to_f32
    z = np.load(f"{GOLDEN}/bank_fewshot.npz")
    meta = json.load(open(f"{GOLDEN}/bank_fewshot.json"))
    cap, img, q = bf16_bits_to_f32(z["cap_bf16"]), bf16_bits_to_f32(z["img_bf16"]), bf16_bits_to_f32(z["q_bf16"])
    few = bf16_bits_to_f32(z["few_bf16"])                       # [C,16,512]
    labels = z["labels"]; C = q.shape[0]
    paths, cmap = make_paths(labels, np.arange(C))
    raw = {"caption_features": cap, "image_features": img, "labels": labels, "filepath": paths}
    prompts = {str(c): {"mean": q[c]} for c in range(C)}
    fewshot = {c: [few[c, i] for i in range(few.shape[1])] for c in range(C)}
    return z, meta, cap, img, q, few, raw, prompts, fewshot, paths, cmap


def test_fewshot_samplers_match_reference():
    """i2i_ranked_sampler_p2p, i2t_rank_sampler, t2t_rank_i2t_tshd_sampler, t2t_rank_i2i_tshd_sampler
    (reference :1016-1133, :831-953) -- verbatim port vs the reference's outputs."""
    z, meta, cap, img, q, few, raw, prompts, fewshot, paths, cmap = _fewshot_case()
    k = int(z["k"])
    feats = so.transform_extracted_fea(raw)
    runs = {
        "i2i_rank": lambda: so.verbatim_t2t_ranked_sampler(prompts, k, 0.0, feats, rank_on_images=True, rank_fewshot=fewshot),
        "i2t_rank": lambda: so.verbatim_t2t_ranked_sampler(prompts, k, 0.0, feats, rank_on_images=False, rank_fewshot=fewshot),
        "t2t_i2t": lambda: so.verbatim_t2t_ranked_t2i_tshd_sampler(prompts, k, 0.0, feats, t2i_threshold=0.25, pred_fewshot=fewshot, pred_on_captions=True),
        "t2t_i2i": lambda: so.verbatim_t2t_ranked_t2i_tshd_sampler(prompts, k, 0.0, feats, t2i_threshold=0.65, pred_fewshot=fewshot, pred_on_captions=False),
    }
    for name, fn in runs.items():
        ms, nd, _ = fn()
        assert nd == meta["counts"][name], name
        got = np.concatenate([feats[str(int(l[0]))]["row_ids"][r] for r, l in zip(ms["row_list"], ms["label_list"])])
        assert got.tolist() == z[f"{name}_rows"].tolist(), name
        assert np.concatenate(ms["label_list"]).tolist() == z[f"{name}_labels"].tolist()
    assert z["t2t_i2i_rows"].tolist() != z["t2t_i2t_rows"].tolist()      # the two predicates really select differently


@pytest.mark.parametrize("name", ["bank_bf16", "bank_f32"])
def test_near_duplicates_match_reference(name):
    """remove_near_duplicates2 (:237-275): duplicate fractions per class and the (file-id keyed) dict."""
    z, meta, cap, img, q = load_bank_case(name)
    class_ids, labels = z["class_ids"], z["labels"]
    paths, _ = make_paths(labels, class_ids)
    raw = {"caption_features": cap, "image_features": img, "labels": class_ids[labels], "filepath": paths}
    feats = so.transform_extracted_fea(raw)
    dd, frac, avg = so.remove_near_duplicates2(feats)
    ref = meta["near_dup"]
    np.testing.assert_allclose(frac, ref["fractions"], atol=1e-12)
    assert abs(avg - ref["avg"]) < 1e-12 and max(frac) > 0
    row = {p: i for i, p in enumerate(paths)}
    assert {k: sorted(row[p] for p in v) for k, v in dd.items() if v} == ref["dict"]


@pytest.mark.parametrize("name", ["bank_bf16", "bank_f32"])
def test_zeroshot_filter_and_random_sampler_match_reference(name):
    """zeroshot_clip_img_filter (:278-329) and random_sampler (:592-661): the oracle's ports against the reference's
    outputs (the random sampler consumes Python's RNG exactly like the reference, so the seed pins the result)."""
    import random
    z, meta, cap, img, q = load_bank_case(name)
    class_ids, labels = z["class_ids"], z["labels"]
    paths, _ = make_paths(labels, class_ids)
    raw = {"caption_features": cap, "image_features": img, "labels": class_ids[labels], "filepath": paths}
    feats = so.transform_extracted_fea(raw)
    row = {p: i for i, p in enumerate(paths)}
    W = np.zeros((int(class_ids.max()) + 1, 512), np.float32)
    W[class_ids] = q
    zs, _ = so.zeroshot_clip_img_filter(feats, W)
    assert {k: sorted(row[p] for p in v) for k, v in zs.items() if v} == meta["zeroshot"]
    prompts = {str(int(class_ids[c])): {"mean": q[c]} for c in range(len(class_ids))}
    dd, _, _ = so.remove_near_duplicates2(feats)
    for tag, thr, th, use_dups in (("plain", 0.0, False, False), ("t2i", 0.2, False, True), ("tailhead", 0.2, True, False)):
        random.seed(1234)
        files, counts, sampled, filtered = so.verbatim_random_sampler(prompts, int(z["k"]), thr, feats,
                                                                      duplicates_dict=dd if use_dups else None, tail_head=th)
        ref = meta["random"][tag]
        assert [row[p] for fl in files for p in fl] == ref["rows"], tag
        assert counts == ref["counts"], tag
