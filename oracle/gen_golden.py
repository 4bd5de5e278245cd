"""Generate golden vectors by running the REFERENCE's own hot-path functions on CPU.

Run in the build container only (needs ``/root/reference``, which does not exist on the GPU box):

    python oracle/gen_golden.py            # rewrites tests/golden/*.npz and *.json

Recipe: stub the two absent third-party modules the hot path never touches, make
``Tensor.cuda`` the identity (``sample_retrieval.py:337-338, 399-400`` hard-code it), chdir to
``retrieval/`` (``extract_mined_feature.py:16`` opens ``../config.yml``), import the module and call
``transform_extracted_fea``, ``t2t_similarity``, ``cal_t2i_similarity``, ``i2i_similarity_p2p``,
``t2t_ranked_sampler`` and ``t2t_ranked_t2i_tshd_sampler`` unmodified.  Inputs are stored in the
fixtures (bf16 bit patterns as uint16, or raw fp32) so tests never depend on RNG reproducibility.
Test infrastructure, not product code.
"""
from __future__ import annotations

import hashlib
from collections import defaultdict
import json
import logging
import os
import pickle
import sys
import tempfile
import types
from argparse import Namespace

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from swat_b200 import synth  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(REPO, "tests", "golden")


def import_reference():
    for m in ("open_clip", "clip"):
        sys.modules.setdefault(m, types.ModuleType(m))
    torch.Tensor.cuda = lambda self, *a, **k: self
    os.chdir(os.path.join(REF, "retrieval"))
    sys.path.insert(0, os.path.join(REF, "retrieval"))
    sys.argv = ["sample_retrieval.py"]
    import sample_retrieval as sr
    return sr


def bf16_bits(t: torch.Tensor) -> np.ndarray:
    return t.contiguous().view(torch.int16).numpy().view(np.uint16)


def run_samplers(sr, raw, prompt_tensors, k, tmp, dataset, unpartitioned=False):
    """Returns dict of results for both samplers.  In the unpartitioned variant every class key
    aliases the whole bank (the reference functions accept this unchanged)."""
    paths = raw["filepath"]
    if unpartitioned:
        feats = {c: {"file_paths": paths, "feats": raw["image_features"],
                     "caption_feats": raw["caption_features"]} for c in prompt_tensors.keys()}
    else:
        feats = sr.transform_extracted_fea(raw)
    args = Namespace(dataset=dataset, output_folder=tmp, prefix="T2T")
    lg = logging.getLogger("golden")
    out = {}
    path_to_row = {p: i for i, p in enumerate(paths)}
    for name, fn in (("t2t", sr.t2t_ranked_sampler), ("t2t_t2i", sr.t2t_ranked_t2i_tshd_sampler), ("t2i", sr.t2i_ranked_sampler)):
        ms, nd = fn(args, lg, prompt_tensors, k, 0.0, feats)
        files = [p for fl in ms["file_list"] for p in fl]
        labels = torch.cat(ms["label_list"]).numpy() if ms["label_list"] else np.zeros(0, np.int64)
        rows = np.asarray([path_to_row[p] for p in files], dtype=np.int64)
        featsum = torch.cat(ms["feature_list"]).double().sum(dim=1).numpy() if ms["feature_list"] else np.zeros(0)
        if name in ("t2t", "t2i"):
            fl_name, sl_name = f"{tmp}/T2T_filtered_list.txt", f"{tmp}/T2T_sampled_list.txt"   # :763,768
        else:
            fl_name, sl_name = f"{tmp}/filtered_list.txt", f"{tmp}/sampled_list.txt"           # :817,822
        # rows of the reference's filtered_list (the rejected rows its walk met, in walk order): "..., <path>, <caption>"
        fl_text = open(fl_name).read()
        filtered_rows = np.asarray([path_to_row[l.split(", ")[-2]] for l in fl_text.split("\n")] if fl_text else [], dtype=np.int64)
        out[name] = dict(rows=rows, labels=labels, counts=nd, featsum=featsum, filtered_rows=filtered_rows,
                         filtered_sha=hashlib.sha256(open(fl_name, "rb").read()).hexdigest(),
                         sampled_sha=hashlib.sha256(open(sl_name, "rb").read()).hexdigest(),
                         sampled_head=open(sl_name).read().split("\n")[:3],
                         n_filtered=len(open(fl_name).read().split("\n")) if os.path.getsize(fl_name) else 0)
    return out, feats


def case_bank(sr, name, n_rows, C, k, seed, dtype, partitioned, rho, tie_block):
    qc, queries, coq = synth.make_queries(C, 1, seed=seed, dtype=dtype)
    cap, img, labels = synth.make_bank(n_rows, qc, seed=seed, dtype=dtype, rho=rho, tie_block=tie_block,
                                       partitioned=partitioned, dup_frac=0.01, chunk=1 << 20)
    class_ids = [3 * c + 1 for c in range(C)]           # non-contiguous ids: keys sort by int, not position
    paths, cmap = synth.make_paths(labels, class_ids=class_ids)
    tmp = tempfile.mkdtemp()
    with open(tmp + "/cap.map", "wb") as f:
        pickle.dump(cmap, f)
    # unpartitioned runs alias one bank under every class key: every path must resolve in the map
    sr.CAPTION_MAP_DICT[name] = tmp + "/cap.map"
    raw = {"caption_features": cap.float(), "image_features": img.float(),
           "labels": torch.tensor([class_ids[int(l)] for l in labels.tolist()]), "filepath": paths}
    prompts = {str(class_ids[c]): {"mean": qc[c].float()} for c in range(C)}
    res_p, feats = run_samplers(sr, raw, prompts, k, tmp, name, unpartitioned=False)
    path_to_row_tmp = {p: i for i, p in enumerate(paths)}
    res_u, _ = run_samplers(sr, raw, prompts, k, tmp, name, unpartitioned=True)
    # near-duplicate removal (remove_near_duplicates2 :237-275) on the regrouped dict
    dd, frac, avg = sr.remove_near_duplicates2(feats)
    dup_fixture = {"dict": {kk: sorted(path_to_row_tmp[p] for p in v) for kk, v in dd.items() if v}, "fractions": frac, "avg": avg}
    # zero-shot filter (zeroshot_clip_img_filter :278-329): head = class prompts at the rows of their class ids (:1489-1490)
    W = torch.zeros(max(class_ids) + 1, 512)
    for c in range(C):
        W[class_ids[c]] = qc[c].float()
    head = torch.nn.Linear(512, W.shape[0], bias=False)
    with torch.no_grad():
        head.weight.copy_(W)
    root = tempfile.mkdtemp()
    for kk in feats.keys():
        os.makedirs(os.path.join(root, kk))
    open(os.path.join(root, "not_a_class.txt"), "w").close()
    with torch.no_grad():
        zs = sr.zeroshot_clip_img_filter(None, None, root, pre_extracted_feats=feats, head=head)
    zs_fixture = {kk: sorted(path_to_row_tmp[p] for p in v) for kk, v in zs.items() if v}
    # random sampler (random_sampler :592-661), seeded like the CLI does (:1710)
    import random as _random
    rnd = {}
    for tag, thr, th, use_dups in (("plain", 0.0, False, False), ("t2i", 0.2, False, True), ("tailhead", 0.2, True, False)):
        _random.seed(1234)
        rargs = Namespace(dataset=name, output_folder=tmp, prefix="RND")
        ms, nd = sr.random_sampler(rargs, logging.getLogger("golden"), prompts, k, thr, feats,
                                   duplicates_dict=dd if use_dups else defaultdict(set), tail_head=th)
        rnd[tag] = dict(rows=[path_to_row_tmp[p] for fl in ms["file_list"] for p in fl], counts=nd,
                        featsum=[float(x) for x in torch.cat(ms["feature_list"]).double().sum(dim=1).tolist()],
                        sampled_sha=hashlib.sha256(open(f"{tmp}/RND_sampled_list.txt", "rb").read()).hexdigest(),
                        filtered_sha=hashlib.sha256(open(f"{tmp}/RND_filtered_list.txt", "rb").read()).hexdigest())
    # regroup fixture: key order + per-class original rows
    path_to_row = {p: i for i, p in enumerate(paths)}
    regroup_keys = list(feats.keys())
    regroup_rows = [np.asarray([path_to_row[p] for p in feats[kk]["file_paths"]], dtype=np.int64) for kk in regroup_keys]
    # primitive fixture: scores of class 0's prompt against the first 64 rows
    sim64 = np.asarray(sr.t2t_similarity(qc[0].float()[None, :], cap[:64].float()), dtype=np.float64)
    t2i64 = np.asarray(sr.cal_t2i_similarity(qc[0].float()[None, :], img[:64].float()), dtype=np.float64)
    arrays = dict(class_ids=np.asarray(class_ids), labels=labels.numpy(), k=np.int64(k),
                  sim64=sim64, t2i64=t2i64)
    if dtype == torch.bfloat16:
        arrays.update(cap_bf16=bf16_bits(cap), img_bf16=bf16_bits(img), q_bf16=bf16_bits(qc))
    else:
        arrays.update(cap_f32=cap.numpy(), img_f32=img.numpy(), q_f32=qc.numpy())
    for tag, res in (("part", res_p), ("unpart", res_u)):
        for m in ("t2t", "t2t_t2i", "t2i"):
            arrays[f"{tag}_{m}_rows"] = res[m]["rows"]
            arrays[f"{tag}_{m}_labels"] = res[m]["labels"]
            arrays[f"{tag}_{m}_featsum"] = res[m]["featsum"]
            if tag == "part":
                arrays[f"{tag}_{m}_filtered_rows"] = res[m]["filtered_rows"]
    for i, kk in enumerate(regroup_keys):
        arrays[f"regroup_rows_{i}"] = regroup_rows[i]
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **arrays)
    meta = dict(name=name, n_rows=n_rows, C=C, k=k, seed=seed, dtype=str(dtype), partitioned=partitioned,
                regroup_keys=regroup_keys, near_dup=dup_fixture, zeroshot=zs_fixture, random=rnd,
                counts={tag: {m: res[m]["counts"] for m in res} for tag, res in (("part", res_p), ("unpart", res_u))},
                diag={tag: {m: {x: res[m][x] for x in ("filtered_sha", "sampled_sha", "sampled_head", "n_filtered")}
                            for m in res} for tag, res in (("part", res_p), ("unpart", res_u))})
    with open(os.path.join(OUT, f"{name}.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print(name, "part", res_p["t2t"]["counts"], res_p["t2t_t2i"]["counts"])
    print(name, "unpart", res_u["t2t"]["counts"], res_u["t2t_t2i"]["counts"])


def case_primitives(sr):
    """R>1 prompt (mean over columns), p2p min/max/mean, N_c == 1, hand-built tie probe."""
    g = torch.Generator().manual_seed(1234)
    X = torch.nn.functional.normalize(torch.randn(257, 512, generator=g), dim=-1)
    P = torch.nn.functional.normalize(torch.randn(3, 512, generator=g), dim=-1)
    F = torch.nn.functional.normalize(torch.randn(16, 512, generator=g), dim=-1)
    arrays = dict(X=X.numpy(), P=P.numpy(), F=F.numpy())
    arrays["t2t_R3"] = np.asarray(sr.t2t_similarity(P, X), dtype=np.float64)
    arrays["t2i_R3"] = np.asarray(sr.cal_t2i_similarity(P, X), dtype=np.float64)
    arrays["t2t_R1"] = np.asarray(sr.t2t_similarity(P[:1], X), dtype=np.float64)
    arrays["t2t_single_row"] = np.asarray(sr.t2t_similarity(P[:1], X[:1]), dtype=np.float64)
    assert isinstance(sr.t2t_similarity(P[:1], X[:1]), list)
    for mode in ("min", "max", "mean"):
        arrays[f"p2p_{mode}"] = np.asarray(sr.i2i_similarity_p2p([f.numpy() for f in F], X, mode), dtype=np.float64)
    # tie probe: 9 rows, 1 class, rows 2/6/7 identical; row 3 best T2T but fails T2I
    q = torch.zeros(512); q[0] = 1.0
    def row(cos, j):
        v = torch.zeros(512); v[0] = cos; v[j] = (1 - cos * cos) ** 0.5; return v
    caps = torch.stack([row(0.10, 1), row(0.50, 2), row(0.25, 3), row(0.60, 4), row(0.40, 5),
                        row(0.30, 6), row(0.25, 3), row(0.25, 3), row(-0.20, 7)])
    imgs = torch.stack([row(0.30, 1), row(0.30, 2), row(0.30, 3), row(0.10, 4), row(0.26, 5),
                        row(0.25, 6), row(0.40, 3), row(0.20, 3), row(0.90, 7)])
    paths = [f"/r/0/{i}.jpg" for i in range(9)]
    tmp = tempfile.mkdtemp()
    with open(tmp + "/m.map", "wb") as f:
        pickle.dump({"0": {str(i): f"c{i}" for i in range(9)}}, f)
    sr.CAPTION_MAP_DICT["probe"] = tmp + "/m.map"
    feats = {"0": {"file_paths": paths, "feats": imgs, "caption_feats": caps}}
    args = Namespace(dataset="probe", output_folder=tmp, prefix="P")
    lg = logging.getLogger("golden")
    ms1, nd1 = sr.t2t_ranked_sampler(args, lg, {"0": {"mean": q}}, 4, 0.0, feats)
    ms2, nd2 = sr.t2t_ranked_t2i_tshd_sampler(args, lg, {"0": {"mean": q}}, 4, 0.0, feats)
    arrays["probe_caps"] = caps.numpy(); arrays["probe_imgs"] = imgs.numpy(); arrays["probe_q"] = q.numpy()
    arrays["probe_t2t_rows"] = np.asarray([int(p.split("/")[-1][:-4]) for p in ms1["file_list"][0]])
    arrays["probe_t2t_t2i_rows"] = np.asarray([int(p.split("/")[-1][:-4]) for p in ms2["file_list"][0]])
    print("probe", arrays["probe_t2t_rows"], arrays["probe_t2t_t2i_rows"])
    np.savez_compressed(os.path.join(OUT, "primitives.npz"), **arrays)


def case_fewshot(sr):
    """The four few-shot samplers (i2i_ranked_sampler_p2p, i2t_rank_sampler, t2t_rank_i2t_tshd_sampler,
    t2t_rank_i2i_tshd_sampler) on a partitioned bank whose relevant rows are close enough to the
    few-shot vectors for the 0.25 / 0.65 thresholds to split them."""
    C, N, k, seed = 5, 1200, 30, 21
    g = torch.Generator().manual_seed(seed)
    unit = lambda x: torch.nn.functional.normalize(x.float(), dim=-1)
    qc = unit(torch.randn(C, 512, generator=g)).to(torch.bfloat16)
    few = {c: [unit(qc[c].float() + 0.45 * unit(torch.randn(512, generator=g))).to(torch.bfloat16).float() for _ in range(16)] for c in range(C)}
    labels = torch.randint(0, C, (N,), generator=g)
    a = torch.rand(N, generator=g) * 0.9
    b = torch.rand(N, generator=g) * 0.95
    base = qc[labels].float()
    cap = unit(a[:, None] * base + torch.sqrt(1 - a * a)[:, None] * unit(torch.randn(N, 512, generator=g))).to(torch.bfloat16)
    img = unit(b[:, None] * base + torch.sqrt(1 - b * b)[:, None] * unit(torch.randn(N, 512, generator=g))).to(torch.bfloat16)
    cap[700:716] = cap[13]; img[700:716] = img[13]; labels[700:716] = labels[13]          # ties
    class_ids = list(range(C))                                                             # fewshot_fea is keyed by int(cls)
    paths, cmap = synth.make_paths(labels, class_ids=class_ids)
    tmp = tempfile.mkdtemp()
    with open(tmp + "/cap.map", "wb") as f:
        pickle.dump(cmap, f)
    sr.CAPTION_MAP_DICT["fewshot"] = tmp + "/cap.map"
    sr.get_fewshot_features = lambda dataset: {c: [x.numpy() for x in few[c]] for c in range(C)}
    raw = {"caption_features": cap.float(), "image_features": img.float(), "labels": labels, "filepath": paths}
    feats = sr.transform_extracted_fea(raw)
    prompts = {str(c): {"mean": qc[c].float()} for c in range(C)}
    args = Namespace(dataset="fewshot", output_folder=tmp, prefix="FS")
    lg = logging.getLogger("golden")
    path_to_row = {p: i for i, p in enumerate(paths)}
    arrays = dict(cap_bf16=bf16_bits(cap), img_bf16=bf16_bits(img), q_bf16=bf16_bits(qc), labels=labels.numpy(), k=np.int64(k),
                  few_bf16=bf16_bits(torch.stack([torch.stack(few[c]) for c in range(C)]).to(torch.bfloat16)))
    counts = {}
    for name, fn in (("i2i_rank", sr.i2i_ranked_sampler_p2p), ("i2t_rank", sr.i2t_rank_sampler),
                     ("t2t_i2t", sr.t2t_rank_i2t_tshd_sampler), ("t2t_i2i", sr.t2t_rank_i2i_tshd_sampler)):
        ms, nd = fn(args, lg, prompts, k, 0.0, feats)
        files = [p for fl in ms["file_list"] for p in fl]
        arrays[f"{name}_rows"] = np.asarray([path_to_row[p] for p in files], dtype=np.int64)
        arrays[f"{name}_labels"] = torch.cat(ms["label_list"]).numpy() if ms["label_list"] else np.zeros(0, np.int64)
        counts[name] = nd
        print("fewshot", name, nd)
    np.savez_compressed(os.path.join(OUT, "bank_fewshot.npz"), **arrays)
    with open(os.path.join(OUT, "bank_fewshot.json"), "w") as f:
        json.dump({"counts": counts, "C": C, "N": N, "k": k}, f, indent=1, sort_keys=True)


def main():
    os.makedirs(OUT, exist_ok=True)
    sr = import_reference()
    torch.set_num_threads(8)
    case_primitives(sr)
    case_fewshot(sr)
    case_bank(sr, "bank_bf16", n_rows=1536, C=6, k=40, seed=11, dtype=torch.bfloat16, partitioned=True,
              rho=0.5, tie_block=96)
    case_bank(sr, "bank_f32", n_rows=640, C=4, k=24, seed=12, dtype=torch.float32, partitioned=True,
              rho=0.6, tie_block=48)


if __name__ == "__main__":
    main()
