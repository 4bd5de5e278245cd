"""Drop-in boundary on the GPU: the host mirror of the reference's samplers and the CLI, against the
reference's golden outputs and the oracle's verbatim port."""
import json
import logging
import os
import pickle
from argparse import Namespace

import numpy as np
import pytest
import torch

from oracle import swat_oracle as so
from tests.golden_util import assert_walk_equal, load_bank_case, make_paths

pytestmark = pytest.mark.gpu
TIE_TOL = 2e-5


def _case(name):
    z, meta, cap, img, q = load_bank_case(name)
    class_ids, labels = z["class_ids"], z["labels"]
    paths, cmap = make_paths(labels, class_ids)
    raw = {"caption_features": torch.from_numpy(cap), "image_features": torch.from_numpy(img),
           "labels": torch.from_numpy(class_ids[labels]), "filepath": paths}
    prompts = {str(class_ids[c]): {"mean": torch.from_numpy(q[c])} for c in range(len(class_ids))}
    return z, meta, cap, img, q, raw, prompts, paths, cmap


@pytest.mark.parametrize("name,dtype", [("bank_bf16", "bf16"), ("bank_f32", "f32")])
def test_samplers_match_reference_outputs(tmp_path, name, dtype):
    """t2t_ranked_sampler / t2t_ranked_t2i_tshd_sampler with the reference's signature and return
    structure, partitioned (transform_extracted_fea) and unpartitioned (aliased dict)."""
    from swat_b200 import retrieval
    z, meta, cap, img, q, raw, prompts, paths, cmap = _case(name)
    S_cap, S_img = so.score_matrix(cap, q), so.score_matrix(img, q)
    cmap_path = str(tmp_path / "cap.map")
    pickle.dump(cmap, open(cmap_path, "wb"))
    args = Namespace(dataset="synthetic", output_folder=str(tmp_path / "out"), prefix="T2T", bank_dtype=dtype, caption_map_path=cmap_path)
    lg = logging.getLogger("t")
    path_row = {p: i for i, p in enumerate(paths)}
    feats_p = retrieval.transform_extracted_fea(raw)
    feats_u = {k: {"file_paths": paths, "feats": raw["image_features"], "caption_feats": raw["caption_features"]} for k in prompts}
    for tag, feats in (("part", feats_p), ("unpart", feats_u)):
        for m, fn in (("t2t", retrieval.t2t_ranked_sampler), ("t2t_t2i", retrieval.t2t_ranked_t2i_tshd_sampler),
                      ("t2i", retrieval.t2i_ranked_sampler)):
            S = S_img if m == "t2i" else S_cap
            ms, nd = fn(args, lg, prompts, int(z["k"]), 0.0, feats)
            assert nd == meta["counts"][tag][m], f"{name} {tag} {m}"
            ref_rows, ref_labels = z[f"{tag}_{m}_rows"], z[f"{tag}_{m}_labels"]
            assert torch.cat(ms["label_list"]).tolist() == ref_labels.tolist()
            pos = 0
            for files, labs, feat in zip(ms["file_list"], ms["label_list"], ms["feature_list"]):
                n = len(files)
                cid = int(labs[0]); c = int(np.nonzero(z["class_ids"] == cid)[0][0])
                got = [path_row[p] for p in files]
                assert_walk_equal(got, ref_rows[pos:pos + n], lambda r, c=c: S[r, c], TIE_TOL, boundary_tol=1e-3, what=f"{name} {tag} {m} {cid}")
                np.testing.assert_allclose(feat.numpy(), img[got], atol=0)          # feature_list = the image features of the winners
                pos += n
            sl = "sampled_list.txt" if m == "t2t_t2i" else "T2T_sampled_list.txt"
            lines = open(os.path.join(args.output_folder, sl)).read().split("\n")
            assert len(lines) == sum(nd.values()) and lines[0].split(", ")[-1].startswith("synthetic caption")
            if tag == "part":
                # filtered_list (:463-469, :761-764, :815-818): the rejected rows the reference's walk met, in walk order
                fl = "filtered_list.txt" if m == "t2t_t2i" else "T2T_filtered_list.txt"
                text = open(os.path.join(args.output_folder, fl)).read()
                got_f = [path_row[l.split(", ")[-2]] for l in text.split("\n")] if text else []
                ref_f = z[f"part_{m}_filtered_rows"].tolist()
                assert len(got_f) == len(ref_f) == meta["diag"]["part"][m]["n_filtered"]
                assert sorted(got_f) == sorted(ref_f), f"{name} {m}: filtered rows differ"
                moved = sum(a != b for a, b in zip(got_f, ref_f))
                assert moved <= 4, f"{name} {m}: {moved} filtered rows out of order"        # near-ties only
                if name == "bank_bf16":             # bf16-valued inputs: the 4-decimal scores in the text agree as well
                    import hashlib
                    same_sha = hashlib.sha256(text.encode()).hexdigest() == meta["diag"]["part"][m]["filtered_sha"]
                    assert same_sha or moved > 0, f"{name} {m}: filtered_list text differs from the reference's"


def test_cli_writes_reference_outputs(tmp_path, monkeypatch):
    """python sample_retrieval.py --prefix ... : {prefix}.txt, {prefix}_num_imgs_sampled.json, sampling.log,
    copy into data/{dataset}/ -- compared with the oracle's verbatim port of the reference script."""
    from swat_b200 import sample_retrieval as cli, shards
    z, meta, cap, img, q, raw, prompts, paths, cmap = _case("bank_bf16")
    monkeypatch.chdir(tmp_path)
    os.makedirs("retrieved/semi-aves"); os.makedirs("data/semi-aves/prompts")
    pth = "retrieved/semi-aves/semi-aves_vitb32_openclip_laion400m_mined.pth"
    shards.save_mined_pth(pth, raw["caption_features"], raw["image_features"], raw["labels"], paths)
    torch.save({"alternates": prompts}, "data/semi-aves/prompts/semi-aves_vitb32_openclip_laion400m_prompt_tensors.pth")
    pickle.dump(cmap, open("cap.map", "wb"))
    for method, prefix, key in (("T2T-rank", "T2T40", "t2t"), ("T2T-rank-T2I-tshd", "T2T40+T2I0.25", "t2t_t2i")):
        fn, n = cli.main(["--prefix", prefix, "--dataset", "semi-aves", "--root", "retrieved", "--num_samples", str(int(z["k"])),
                          "--sampling_method", method, "--bank_dtype", "bf16", "--data_dir", "data", "--caption_map_path", "cap.map",
                          "--log_mode", "file"])
        out = f"output/semi-aves_vitb32_openclip_laion400m_{prefix}"
        assert fn == f"{out}/{prefix}.txt" and os.path.exists(f"{out}/sampling.log")
        text = open(fn).read()
        assert open(f"data/semi-aves/{prefix}.txt").read() == text
        counts = json.load(open(f"{out}/{prefix}_num_imgs_sampled.json"))
        assert counts == meta["counts"]["part"][key] and n == sum(counts.values())
        ref_lines = [f"{paths[r]} {l} 0" for r, l in zip(z[f"part_{key}_rows"].tolist(), z[f"part_{key}_labels"].tolist())]
        got_lines = text.strip("\n").split("\n")
        assert len(got_lines) == len(ref_lines)
        assert [l.split(" ")[1:] for l in got_lines] == [l.split(" ")[1:] for l in ref_lines]      # class-major, labels, source flag
        # Byte-identical except where the reference itself is not reproducible: its MKL GEMV returns scores 1 ulp apart
        # for BIT-IDENTICAL rows (position inside the BLAS blocking), so its order among exact duplicates is not index
        # order.  Every differing line must be such a swap: same multiset of lines, scores of the swapped rows equal.
        assert sorted(got_lines) == sorted(ref_lines)
        S = so.score_matrix(cap, q)
        row_of = {p: i for i, p in enumerate(paths)}
        cls_of = {int(cid): c for c, cid in enumerate(z["class_ids"].tolist())}
        differing = 0
        for a, b in zip(got_lines, ref_lines):
            if a != b:
                differing += 1
                c = cls_of[int(a.split(" ")[1])]
                assert abs(S[row_of[a.split(" ")[0]], c] - S[row_of[b.split(" ")[0]], c]) <= 1e-6, (a, b)
        assert differing <= 8, f"{differing}/{len(ref_lines)} lines differ"
        for line in got_lines:                                                                     # MyDataset's parser (dataset_utils.py:148-154)
            p, lab, src = line.strip("\n").split(" ")
            assert int(src) == 0 and int(lab) in z["class_ids"].tolist()


def test_sampling_seam_exclusion_wiring_and_random(tmp_path, monkeypatch):
    """``sampling(args, logger, model, preprocess, metrics, dataset_root)`` with the reference's positional signature and
    its module-global prompt tensors (:1471, :1489); ``--image_dedup`` / ``--zeroshot_img_filter`` compute the exclusion
    sets and every sampler receives them (:1484-1507); ``Random`` dispatches through the CLI (:1517-1526)."""
    import inspect
    from swat_b200 import retrieval, sample_retrieval as cli, shards
    assert list(inspect.signature(retrieval.sampling).parameters)[:6] == ["args", "logger", "model", "preprocess", "metrics", "dataset_root"]
    z, meta, cap, img, q, raw, prompts, paths, cmap = _case("bank_bf16")
    monkeypatch.chdir(tmp_path)
    os.makedirs("retrieved/semi-aves"); os.makedirs("data/semi-aves/prompts")
    pth = "retrieved/semi-aves/semi-aves_vitb32_openclip_laion400m_mined.pth"
    shards.save_mined_pth(pth, raw["caption_features"], raw["image_features"], raw["labels"], paths)
    torch.save({"alternates": prompts}, "data/semi-aves/prompts/semi-aves_vitb32_openclip_laion400m_prompt_tensors.pth")
    pickle.dump(cmap, open("cap.map", "wb"))
    k = int(z["k"])
    common = ["--dataset", "semi-aves", "--root", "retrieved", "--num_samples", str(k), "--bank_dtype", "bf16", "--data_dir", "data",
              "--caption_map_path", "cap.map", "--log_mode", "file"]
    # exclusion sets through the CLI: the oracle's ports of both producers + the verbatim sampler give the expectation
    raw_np = {kk: (v.numpy() if torch.is_tensor(v) else v) for kk, v in raw.items()}
    o_feats = so.transform_extracted_fea(raw_np)
    o_dd, _, _ = so.remove_near_duplicates2(o_feats)
    W = np.stack([prompts[c]["mean"].numpy() for c in prompts.keys()])          # features.prompt_sampler(..., 'mean') (:1489)
    o_zs, _ = so.zeroshot_clip_img_filter(o_feats, W)
    o_ms, o_nd, _ = so.verbatim_t2t_ranked_t2i_tshd_sampler({c: {"mean": v["mean"].numpy()} for c, v in prompts.items()}, k, 0.0, o_feats,
                                                            duplicates_dict={c: set(v) for c, v in o_dd.items()},
                                                            filtered_images_dict={c: set(v) for c, v in o_zs.items()})
    fn, n = cli.main(["--prefix", "EXCL", "--sampling_method", "T2T-rank-T2I-tshd", "--image_dedup", "--zeroshot_img_filter"] + common)
    counts = json.load(open("output/semi-aves_vitb32_openclip_laion400m_EXCL/EXCL_num_imgs_sampled.json"))
    assert counts == {c: int(v) for c, v in o_nd.items()} and n == sum(counts.values())
    got = [l.split(" ")[0] for l in open(fn).read().strip("\n").split("\n")] if n else []
    want = [p for fl in o_ms["file_list"] for p in fl]
    assert sorted(got) == sorted(want) and sum(a != b for a, b in zip(got, want)) <= 8
    banned = {p for v in o_dd.values() for p in v} | {p for v in o_zs.values() for p in v}
    assert not (set(got) & banned)
    # Random through the CLI, seeded like the reference (:1710): the golden run used seed 1234
    fn, n = cli.main(["--prefix", "RND", "--sampling_method", "Random", "--seed", "1234"] + common)
    ref = meta["random"]["plain"]
    row = {p: i for i, p in enumerate(paths)}
    assert [row[l.split(" ")[0]] for l in open(fn).read().strip("\n").split("\n")] == ref["rows"]
    assert json.load(open("output/semi-aves_vitb32_openclip_laion400m_RND/RND_num_imgs_sampled.json")) == ref["counts"]
    # the seam itself, called like the reference does
    retrieval.prompt_tensors_dict = {"alternates": prompts}
    args = Namespace(dataset="semi-aves", model_cfg="vitb32_openclip_laion400m", prompt_name="alternates", sampling_method="T2T-rank",
                     num_samples=k, sampling_threshold=0.0, zeroshot_img_filter=False, image_dedup=False, prefix="SEAM",
                     output_folder=str(tmp_path / "seam"), bank_dtype="bf16", caption_map_path="cap.map")
    os.makedirs(args.output_folder)
    path, ct = retrieval.sampling(args, logging.getLogger("t"), None, None, None, "retrieved/semi-aves", copy_to=None)
    assert ct == sum(meta["counts"]["part"]["t2t"].values()) and path.endswith("SEAM.txt")


def test_s1_primitives_on_gpu():
    from swat_b200 import retrieval
    z = np.load("tests/golden/primitives.npz")
    X, P, F = torch.from_numpy(z["X"]), torch.from_numpy(z["P"]), z["F"]
    np.testing.assert_allclose(retrieval.t2t_similarity(P, X), z["t2t_R3"], atol=2e-6)
    np.testing.assert_allclose(retrieval.cal_t2i_similarity(P[:1], X), z["t2t_R1"], atol=2e-6)
    one = retrieval.t2t_similarity(P[:1], X[:1])
    assert isinstance(one, list) and len(one) == 1
    for mode in ("min", "max", "mean"):
        np.testing.assert_allclose(retrieval.i2i_similarity_p2p([f for f in F], X, mode), z[f"p2p_{mode}"], atol=2e-6)
    with pytest.raises(ValueError):
        retrieval.i2i_similarity_p2p([f for f in F], X, "median")


def test_fewshot_samplers_match_reference_outputs(tmp_path):
    """The four few-shot samplers of the reference on the GPU: rank by the mean over 16 few-shot vectors
    (I2I-rank, I2T-rank) and T2T ranking with a max-over-few-shot predicate (>= 0.25 captions, >= 0.65 images)."""
    from swat_b200 import retrieval
    from tests.test_oracle_golden import _fewshot_case
    z, meta, cap, img, q, few, raw, prompts, fewshot, paths, cmap = _fewshot_case()
    k = int(z["k"]); labels = z["labels"]; C = q.shape[0]
    cmap_path = str(tmp_path / "cap.map"); pickle.dump(cmap, open(cmap_path, "wb"))
    raw_t = {"caption_features": torch.from_numpy(cap), "image_features": torch.from_numpy(img),
             "labels": torch.from_numpy(labels), "filepath": paths}
    prompts_t = {c: {"mean": torch.from_numpy(v["mean"])} for c, v in prompts.items()}
    args = Namespace(dataset="fewshot", output_folder=str(tmp_path / "out"), prefix="FS", bank_dtype="bf16", caption_map_path=cmap_path,
                     fewshot_features={c: [torch.from_numpy(x) for x in v] for c, v in fewshot.items()})
    feats = retrieval.transform_extracted_fea(raw_t)
    path_row = {p: i for i, p in enumerate(paths)}
    fmean = few.mean(axis=1)
    rank_scores = {"i2i_rank": (img @ few.reshape(-1, 512).T).reshape(len(img), C, -1).mean(-1),
                   "i2t_rank": (cap @ few.reshape(-1, 512).T).reshape(len(cap), C, -1).mean(-1),
                   "t2t_i2t": so.score_matrix(cap, q), "t2t_i2i": so.score_matrix(cap, q)}
    fns = {"i2i_rank": retrieval.i2i_ranked_sampler_p2p, "i2t_rank": retrieval.i2t_rank_sampler,
           "t2t_i2t": retrieval.t2t_rank_i2t_tshd_sampler, "t2t_i2i": retrieval.t2t_rank_i2i_tshd_sampler}
    for name, fn in fns.items():
        ms, nd = fn(args, logging.getLogger("t"), prompts_t, k, 0.0, feats)
        assert nd == meta["counts"][name], name
        ref_rows, S = z[f"{name}_rows"], rank_scores[name]
        pos = 0
        for files, labs in zip(ms["file_list"], ms["label_list"]):
            n = len(files); c = int(labs[0])
            assert_walk_equal([path_row[p] for p in files], ref_rows[pos:pos + n], lambda r, c=c: S[r, c], TIE_TOL, boundary_tol=1e-3,
                              what=f"{name} class {c}")
            pos += n
        assert torch.cat(ms["label_list"]).tolist() == z[f"{name}_labels"].tolist()


@pytest.mark.parametrize("name", ["bank_bf16", "bank_f32"])
def test_near_duplicate_removal_matches_reference(name):
    """remove_near_duplicates2 on the GPU (triangular Gram kernel) vs the reference's outputs, and its result
    used as the samplers' exclusion set."""
    from swat_b200 import retrieval
    z, meta, cap, img, q, raw, prompts, paths, cmap = _case(name)
    feats = retrieval.transform_extracted_fea(raw)
    dd, frac, avg = retrieval.remove_near_duplicates2(feats)
    ref = meta["near_dup"]
    np.testing.assert_allclose(frac, ref["fractions"], atol=1e-12)
    row = {p: i for i, p in enumerate(paths)}
    assert {k: sorted(row[p] for p in v) for k, v in dd.items() if v} == ref["dict"]
    # positional variant: flagged rows really have an earlier near-identical row of their class
    dp, _, _ = retrieval.remove_near_duplicates2(feats, positional=True)
    o_dd, _, _ = so.remove_near_duplicates2(so.transform_extracted_fea({k: (v.numpy() if torch.is_tensor(v) else v) for k, v in raw.items()}), positional=True)
    assert {k: sorted(v) for k, v in dp.items()} == {k: sorted(v) for k, v in o_dd.items()}
    # as exclusion set of the sampler: no excluded path may be sampled, counts follow the oracle
    import logging
    from argparse import Namespace
    args = Namespace(dataset="synthetic", output_folder="/tmp/swat_dedup_out", prefix="D", bank_dtype="bf16" if name == "bank_bf16" else "f32",
                     caption_map_path="/nonexistent")
    ms, nd = retrieval.t2t_ranked_sampler(args, logging.getLogger("t"), prompts, int(z["k"]), 0.0, feats, duplicates_dict=dp)
    sampled = {p for fl in ms["file_list"] for p in fl}
    assert not (sampled & {p for v in dp.values() for p in v})
    o_ms, o_nd, _ = so.verbatim_t2t_ranked_sampler({k: {"mean": v["mean"].numpy()} for k, v in prompts.items()}, int(z["k"]), 0.0,
                                                  so.transform_extracted_fea({k: (v.numpy() if torch.is_tensor(v) else v) for k, v in raw.items()}),
                                                  duplicates_dict={k: set(v) for k, v in o_dd.items()})
    assert nd == o_nd


@pytest.mark.parametrize("name", ["bank_bf16", "bank_f32"])
def test_zeroshot_filter_and_random_sampler(name):
    """zeroshot_clip_img_filter (GPU logits + argmax) and random_sampler (host shuffle, GPU T2I predicate) against the
    reference's outputs, including the bytes of the two diagnostic files of the random sampler."""
    import hashlib, logging, random, tempfile
    from argparse import Namespace
    from swat_b200 import retrieval
    z, meta, cap, img, q, raw, prompts, paths, cmap = _case(name)
    feats = retrieval.transform_extracted_fea(raw)
    class_ids = z["class_ids"]
    row = {p: i for i, p in enumerate(paths)}
    W = torch.zeros(int(class_ids.max()) + 1, 512)
    W[torch.from_numpy(class_ids)] = torch.as_tensor(q).float()
    head = torch.nn.Linear(512, W.shape[0], bias=False)
    with torch.no_grad():
        head.weight.copy_(W)
    root = tempfile.mkdtemp()
    for kk in feats.keys():
        os.makedirs(os.path.join(root, kk))
    zs = retrieval.zeroshot_clip_img_filter(None, None, root, pre_extracted_feats=feats, head=head)
    assert {k: sorted(row[p] for p in v) for k, v in zs.items() if v} == meta["zeroshot"]
    dd, _, _ = retrieval.remove_near_duplicates2(feats)
    tmp = tempfile.mkdtemp()
    cm = tmp + "/cap.map"
    with open(cm, "wb") as f:
        pickle.dump(cmap, f)
    for tag, thr, th, use_dups in (("plain", 0.0, False, False), ("t2i", 0.2, False, True), ("tailhead", 0.2, True, False)):
        random.seed(1234)
        args = Namespace(dataset="synthetic", output_folder=tmp, prefix="RND", caption_map_path=cm)
        ms, nd = retrieval.random_sampler(args, logging.getLogger("t"), prompts, int(z["k"]), thr, feats,
                                          duplicates_dict=dd if use_dups else None, tail_head=th)
        ref = meta["random"][tag]
        assert [row[p] for fl in ms["file_list"] for p in fl] == ref["rows"], tag
        assert nd == ref["counts"], tag
        np.testing.assert_allclose(torch.cat(ms["feature_list"]).double().sum(dim=1).numpy(), ref["featsum"], atol=1e-9)
        if name == "bank_bf16" or thr == 0.0:      # fp32 T2I scores print with 4 decimals; bf16-valued inputs make them exact
            assert hashlib.sha256(open(f"{tmp}/RND_sampled_list.txt", "rb").read()).hexdigest() == ref["sampled_sha"], tag
            assert hashlib.sha256(open(f"{tmp}/RND_filtered_list.txt", "rb").read()).hexdigest() == ref["filtered_sha"], tag


@pytest.mark.parametrize("name,dt", [("bank_bf16", "bf16"), ("bank_f32", "f32")])
def test_flat_shard_to_hbm_and_sampler(tmp_path, name, dt):
    """Loader row of SURVEY 8f: mined .pth -> flat shard -> HBM (whole shard and a rank's row range, small staging
    chunks so both pinned buffers cycle) -> same bytes, and the sampler on the loaded dict gives the reference rows."""
    from swat_b200 import retrieval, shards
    z, meta, cap, img, q, raw, prompts, paths, cmap = _case(name)
    pth = str(tmp_path / "mined.pth")
    shards.save_mined_pth(pth, raw["caption_features"], raw["image_features"], raw["labels"], paths)
    shards.convert_pth_to_flat(pth, str(tmp_path / "flat"), dt)
    fs = shards.FlatShard(str(tmp_path / "flat"))
    want_c = raw["caption_features"].to(torch.bfloat16 if dt == "bf16" else torch.float32)
    want_i = raw["image_features"].to(torch.bfloat16 if dt == "bf16" else torch.float32)
    for native in (True, False):           # the C-ABI loader (swat_bank_load) and the torch staging path it replaces
        c, i = fs.to_device("cuda:0", pinned_chunk_rows=100, native=native)
        assert torch.equal(c.cpu(), want_c) and torch.equal(i.cpu(), want_i)
        c, i = fs.to_device("cuda:0", rows=slice(37, 411), pinned_chunk_rows=64, native=native)
        assert torch.equal(c.cpu(), want_c[37:411]) and torch.equal(i.cpu(), want_i[37:411])
    print("GPUDirect Storage used:", fs.used_gds)
    from swat_b200 import _lib
    with pytest.raises(_lib.SwatError):
        _lib.bank_load(retrieval.get_context(0), str(tmp_path / "flat" / "caption.bin"), want_c.dtype, 0, 10 ** 7)      # beyond the file
    feats = retrieval.transform_extracted_fea(fs.as_mined_dict())
    args = Namespace(dataset="synthetic", output_folder=str(tmp_path / "out"), prefix="T2T", bank_dtype=dt, caption_map_path="/nonexistent")
    ms, nd = retrieval.t2t_ranked_t2i_tshd_sampler(args, logging.getLogger("t"), prompts, int(z["k"]), 0.0, feats)
    assert nd == meta["counts"]["part"]["t2t_t2i"]
    row = {p: r for r, p in enumerate(paths)}
    S = so.score_matrix(cap, q)
    ref_rows, pos = z["part_t2t_t2i_rows"], 0
    for files, labs in zip(ms["file_list"], ms["label_list"]):
        n = len(files)
        c = int(np.nonzero(z["class_ids"] == int(labs[0]))[0][0])
        assert_walk_equal([row[p] for p in files], ref_rows[pos:pos + n], lambda r, c=c: S[r, c], TIE_TOL, boundary_tol=1e-3,
                          what=f"{name} flat shard class {c}")
        pos += n
