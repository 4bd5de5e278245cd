"""CPU-side checks: the C-ABI library builds/loads and exports every symbol the header declares, the
product path fails loudly without a GPU, and the host logic (regrouper, shard formats, split writer)
matches the oracle / the reference's golden vectors.  No kernel is launched here."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

from oracle import swat_oracle as so
from tests.golden_util import load_bank_case, make_paths

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_header_symbols():
    import __graft_entry__ as g
    g.build()
    from swat_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(REPO, "include", "swat_b200.h")).read()
    declared = set(re.findall(r"\b(swat_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/swat_b200.h but not exported"
    assert declared == set(_lib.EXPORTS)
    assert lib.swat_version() == 200
    deps = os.popen(f"ldd {_lib.LIB_PATH}").read()
    assert "libcuda.so" not in deps          # driver entry points are resolved at run time


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_product_path_fails_loudly_without_gpu():
    from swat_b200 import _lib, retrieval
    with pytest.raises(RuntimeError):
        _lib.Context(0)
    with pytest.raises(RuntimeError):
        retrieval.t2t_similarity(torch.zeros(1, 512), torch.zeros(4, 512))
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.swat_ctx_create(0, ctypes.byref(h)) == -3                      # SWAT_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.swat_last_error()


def test_product_package_never_imports_the_oracle():
    for root, _, files in os.walk(os.path.join(REPO, "swat_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert "oracle" not in src.replace("swat_oracle", "oracle") or "import oracle" not in src and "from oracle" not in src, f


@pytest.mark.parametrize("name", ["bank_bf16", "bank_f32"])
def test_regrouper_matches_reference(name):
    """transform_extracted_fea: key order (first appearance) and per-class row order of the reference."""
    from swat_b200 import retrieval
    z, meta, cap, img, q = load_bank_case(name)
    class_ids = z["class_ids"]; labels = z["labels"]
    paths, _ = make_paths(labels, class_ids)
    raw = {"caption_features": torch.from_numpy(cap), "image_features": torch.from_numpy(img),
           "labels": torch.from_numpy(class_ids[labels]), "filepath": paths}
    out = retrieval.transform_extracted_fea(raw)
    assert list(out.keys()) == meta["regroup_keys"]
    for i, kk in enumerate(out.keys()):
        assert out.rows_of(kk).tolist() == z[f"regroup_rows_{i}"].tolist()
        e = out[kk]
        assert e["file_paths"] == [paths[j] for j in z[f"regroup_rows_{i}"].tolist()]
        assert torch.equal(e["caption_feats"], torch.from_numpy(cap[z[f"regroup_rows_{i}"]]))
    ref = so.transform_extracted_fea({k: (v.numpy() if torch.is_tensor(v) else v) for k, v in raw.items()})
    assert list(ref.keys()) == list(out.keys())


def test_flat_shard_roundtrip(tmp_path):
    from swat_b200 import shards
    g = torch.Generator().manual_seed(0)
    cap = torch.nn.functional.normalize(torch.randn(300, 512, generator=g), dim=-1)
    img = torch.nn.functional.normalize(torch.randn(300, 512, generator=g), dim=-1)
    labels = torch.randint(0, 7, (300,), generator=g)
    paths = [f"/r/{int(l)}/{i}.jpg" for i, l in enumerate(labels)]
    pth = str(tmp_path / "ds_cfg_mined.pth")
    shards.save_mined_pth(pth, cap, img, labels, paths)
    d = shards.load_mined_pth(pth)
    assert torch.equal(d["caption_features"], cap) and d["filepath"] == paths
    for dt in ("f32", "bf16"):
        meta = shards.convert_pth_to_flat(pth, str(tmp_path / dt), dt)
        assert meta["n_rows"] == 300
        fs = shards.FlatShard(str(tmp_path / dt))
        want = cap if dt == "f32" else cap.to(torch.bfloat16)
        assert torch.equal(fs.caption(), want) and torch.equal(fs.caption(slice(10, 20)), want[10:20])
        assert torch.equal(fs.image(), img if dt == "f32" else img.to(torch.bfloat16))
        assert fs.labels().tolist() == labels.tolist() and fs.paths() == paths
        md = fs.as_mined_dict()
        assert set(md) == {"caption_features", "image_features", "labels", "filepath"}


def test_split_writer_matches_reference_format(tmp_path):
    """"<path> <label> 0\\n" per row (save_sample_file_list :1457-1462) and what MyDataset parses
    (utils/datasets/dataset_utils.py:141-154)."""
    from argparse import Namespace
    from swat_b200 import retrieval
    args = Namespace(output_folder=str(tmp_path), prefix="T2T500")
    files = ["/a/1/5.jpg", "/a/1/6.jpg", "/a/2/7.jpg"]
    fn = retrieval.save_sample_file_list(args, files, torch.tensor([1, 1, 2]), copy_to=str(tmp_path / "data"))
    text = open(fn).read()
    assert text == "".join(so.format_split_lines([files[:2], files[2:]], [np.array([1, 1]), np.array([2])]))
    assert open(tmp_path / "data" / "T2T500.txt").read() == text
    for line in text.strip("\n").split("\n"):
        path, label, src = line.split(" ")
        assert int(src) == 0 and int(label) in (1, 2) and path.endswith(".jpg")


def test_path_tables_follow_the_reference_layout(monkeypatch):
    from swat_b200 import config
    monkeypatch.setenv("SWAT_RETRIEVED_PATH", "/scratch/retrieved")
    assert config.CAPTION_MAP_DICT["semi-aves"] == "/scratch/retrieved/semi-aves/semi-aves_metadata-all-0.0-LAION400M.map"
    assert config.CAPTION_MAP_DICT["dtd"] == "/scratch/retrieved/dtd/dtd_metadata-random-0.0-LAION400M.map"
    assert config.MINED_DATASET_ROOT_DICT["imagenet"].endswith("imagenet_retrieved_LAION400M-all_synonyms-random")
    assert config.CAPTION_MAP_DICT.get("nope") is None


def test_bench_reference_arm_line_shape():
    """--impl reference prints one JSON line with the contract's keys (tiny sample so it runs in seconds)."""
    import subprocess, sys
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--classes", "8", "--k", "20", "--no-cpu"], capture_output=True, text=True, timeout=600,
                         env={**os.environ, "SWAT_BENCH_REF_ROWS": "3000"})
    line = json.loads(out.stdout.strip().split("\n")[-1])
    assert line["impl"] == "reference" and line["unit"] == "rows/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    # both arms must describe the same workload: the driver compares the `config` objects
    import argparse, importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(REPO, "bench.py"))
    bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
    ours = bench.config_dict(argparse.Namespace(rows=10_000_000, classes=8, k=20, t2t_only=False), 1)
    assert line["config"] == ours


def test_random_sampler_threshold0_is_host_only(tmp_path):
    """random_sampler with threshold 0 (the way sampling() calls it, :1517-1526) involves no scores at all: the mirror
    must reproduce the reference's shuffle + walk and the bytes of its two diagnostic files without touching a GPU."""
    import hashlib, logging, pickle, random
    from argparse import Namespace
    import torch
    from swat_b200 import retrieval
    from tests.golden_util import load_bank_case, make_paths
    for name in ("bank_bf16", "bank_f32"):
        z, meta, cap, img, q = load_bank_case(name)
        class_ids, labels = z["class_ids"], z["labels"]
        paths, cmap = make_paths(labels, class_ids)
        raw = {"caption_features": torch.from_numpy(cap), "image_features": torch.from_numpy(img),
               "labels": torch.from_numpy(class_ids[labels]), "filepath": paths}
        feats = retrieval.transform_extracted_fea(raw)
        prompts = {str(class_ids[c]): {"mean": torch.from_numpy(q[c])} for c in range(len(class_ids))}
        cm = str(tmp_path / f"{name}.map")
        with open(cm, "wb") as f:
            pickle.dump(cmap, f)
        random.seed(1234)
        args = Namespace(dataset="synthetic", output_folder=str(tmp_path / name), prefix="RND", caption_map_path=cm)
        ms, nd = retrieval.random_sampler(args, logging.getLogger("t"), prompts, int(z["k"]), 0.0, feats)
        ref = meta["random"]["plain"]
        row = {p: i for i, p in enumerate(paths)}
        assert [row[p] for fl in ms["file_list"] for p in fl] == ref["rows"] and nd == ref["counts"]
        assert hashlib.sha256(open(f"{args.output_folder}/RND_sampled_list.txt", "rb").read()).hexdigest() == ref["sampled_sha"]
        assert hashlib.sha256(open(f"{args.output_folder}/RND_filtered_list.txt", "rb").read()).hexdigest() == ref["filtered_sha"]


def test_filtered_list_walk_reproduces_reference_rows():
    """``filtered_list.txt`` (sample_retrieval.py:463-469, :521-527) is host logic over per-row scores: with the oracle's
    fp32 scores standing in for ``swat_score_rows`` it must list exactly the rows the reference rejected, in walk order."""
    import inspect
    import torch
    from oracle import swat_oracle as so
    from swat_b200 import retrieval
    from tests.golden_util import load_bank_case, make_paths
    assert list(inspect.signature(retrieval.sampling).parameters)[:6] == ["args", "logger", "model", "preprocess", "metrics", "dataset_root"]
    for name in ("bank_bf16", "bank_f32"):
        z, meta, cap, img, q = load_bank_case(name)
        class_ids, labels, k = z["class_ids"], z["labels"], int(z["k"])
        paths, _ = make_paths(labels, class_ids)
        classes = [str(c) for c in sorted(class_ids.tolist())]
        row_class = torch.from_numpy(labels.astype(np.int32))
        S, I = so.score_matrix(cap, q), so.score_matrix(img, q)
        own = lambda M: torch.from_numpy(M[np.arange(len(labels)), labels].astype(np.float32))
        row = {p: i for i, p in enumerate(paths)}
        for m, t2t_all, pred_all in (("t2t", own(S), None), ("t2t_t2i", own(S), own(I)), ("t2i", own(I), None)):
            lines = retrieval._filtered_lines(classes, row_class, paths, t2t_all, pred_all, None, k, 0.0, 0.25, None)
            got = [row[l.split(", ")[-2]] for l in lines]
            ref = z[f"part_{m}_filtered_rows"].tolist()
            assert len(got) == len(ref) == meta["diag"]["part"][m]["n_filtered"], (name, m)
            assert sorted(got) == sorted(ref) and sum(a != b for a, b in zip(got, ref)) <= 4, (name, m)


def test_unpartitioned_exclusion_runs_deeper_and_filters_on_the_host():
    from swat_b200 import retrieval
    sets = retrieval._exclusion_sets({"3": {"a", "b"}, 4: set()}, {"3": {"c"}, "5": {"d"}})
    assert sets == {"3": {"a", "b", "c"}, "5": {"d"}}
    bits, ex = retrieval._exclusion_bitmap(["p0", "a", "d", "c"], ["3", "5"], np.array([0, 0, 1, 1], np.int32), sets)
    assert ex.tolist() == [False, True, True, False]          # "c" is excluded for class 3 only, the row belongs to class 5
    assert int(bits[0]) == 0b0110
