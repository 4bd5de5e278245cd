"""CPU oracle for SWAT's retrieval hot path (score -> per-class top-k -> T2I filter walk).

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The product
path (``swat_b200``) never routes through this file and fails loudly without its CUDA library.

It restates, in numpy, the algorithm of ``/root/reference/retrieval/sample_retrieval.py``
(file:line citations below are into that file unless another file is named).  Parity pin:
``oracle/gen_golden.py`` imports the reference's own functions in the build container, runs them
on seeded synthetic inputs and commits the outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every function here against those vectors.  The reference
itself ships no numeric golden vectors or tests, so those generated fixtures
are the pin.

Two families of functions:

* ``verbatim_*``  -- the reference's per-class loop restated step for step (GEMV, Python
  ``sorted`` on zipped tuples, accept/walk with early break).  Slow by construction; this is what
  ``bench.py --impl reference`` times ("port" of the reference CPU path).
* ``topk_walk`` / ``score_matrix`` -- a vectorised restatement (one GEMM, exact tie-aware
  selection) used as the fast checker at >= 1 M rows.  ``tests/test_oracle_golden.py`` proves it
  equal to the verbatim family and to the reference's own outputs.
"""
from __future__ import annotations

from collections import defaultdict
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

REDUCE_NONE, REDUCE_MEAN, REDUCE_MAX, REDUCE_MIN = 0, 1, 2, 3
_REDUCE_NAMES = {"none": REDUCE_NONE, "mean": REDUCE_MEAN, "max": REDUCE_MAX, "min": REDUCE_MIN}


def reduce_code(reduce) -> int:
    if isinstance(reduce, str):
        return _REDUCE_NAMES[reduce.lower()]
    return int(reduce)


# --------------------------------------------------------------------------------------
# L1 primitives
# --------------------------------------------------------------------------------------
def similarity(class_prompt: np.ndarray, embeddings: np.ndarray) -> List[float]:
    """``t2t_similarity`` (:397-416) and ``cal_t2i_similarity`` (:335-353) -- same arithmetic.

    ``s = X @ q^T`` in fp32; if the prompt has R > 1 rows the R columns are averaged (:403-404,
    :341-342).  Returns a Python list of floats; a single row yields a 1-element list (:413-414).
    """
    q = np.asarray(class_prompt, dtype=np.float32)
    x = np.asarray(embeddings, dtype=np.float32)
    if q.ndim == 1:
        q = q[None, :]
    s = x @ q.T
    if s.shape[-1] > 1:
        s = s.mean(axis=-1, dtype=np.float32)
    result = np.squeeze(s).tolist()
    if isinstance(result, float):
        result = [result]
    return result


def similarity_p2p(fewshot_embedding: np.ndarray, embeddings: np.ndarray, mode: str) -> List[float]:
    """``i2i_similarity_p2p`` (:369-394): min / max / mean over the columns of ``X @ F^T``."""
    f = np.asarray(fewshot_embedding, dtype=np.float32)
    x = np.asarray(embeddings, dtype=np.float32)
    s = x @ f.T
    if mode == "min":
        sim = s.min(axis=-1)
    elif mode == "max":
        sim = s.max(axis=-1)
    elif mode == "mean":
        sim = s.mean(axis=-1, dtype=np.float32)
    else:
        raise ValueError("Invalid mode type.")
    result = np.squeeze(sim).tolist()
    if isinstance(result, float):
        result = [result]
    return result


# --------------------------------------------------------------------------------------
# loader / regrouper
# --------------------------------------------------------------------------------------
def transform_extracted_fea(pre_extracted_feats: dict) -> dict:
    """``transform_extracted_fea`` (:1387-1415): regroup rows by label.

    Keys are ``str(label)`` in first-appearance order; rows keep file order inside a class;
    classes with no rows are simply absent.  ``row_ids`` is an extra field (the original row
    index of every regrouped row) that the tests use to compare with un-regrouped kernels.
    """
    img = np.asarray(pre_extracted_feats["image_features"])
    cap = np.asarray(pre_extracted_feats["caption_features"])
    labels = np.asarray(pre_extracted_feats["labels"]).astype(np.int64)
    paths = pre_extracted_feats["filepath"]
    order = np.argsort(labels, kind="stable")
    sorted_labels = labels[order]
    uniq, first_pos = np.unique(labels, return_index=True)
    uniq = uniq[np.argsort(first_pos, kind="stable")]          # first-appearance order (:1401-1402)
    starts = np.searchsorted(sorted_labels, uniq, side="left")
    ends = np.searchsorted(sorted_labels, uniq, side="right")
    out = {}
    for lab, s, e in zip(uniq.tolist(), starts.tolist(), ends.tolist()):
        rows = order[s:e]
        out[str(lab)] = {
            "file_paths": [paths[i] for i in rows.tolist()],
            "feats": img[rows],
            "caption_feats": cap[rows],
            "row_ids": rows,
        }
    return out


# --------------------------------------------------------------------------------------
# exclusion-set producer
# --------------------------------------------------------------------------------------
def near_duplicate_positions(img_embeddings: np.ndarray, threshold: float = 0.9) -> np.ndarray:
    """Positions j with an earlier row i < j whose cosine exceeds the threshold
    (``np.triu(X @ X^T, k=1) > 0.9`` -> ``j_indices``, :254-259)."""
    x = np.asarray(img_embeddings, dtype=np.float32)
    sim = x @ x.T
    _, j = np.where(np.triu(sim, k=1) > threshold)
    return np.unique(j)


def remove_near_duplicates2(pre_extracted_feats: dict, threshold: float = 0.9, positional: bool = False):
    """``remove_near_duplicates2`` (:237-275), including its file-id-vs-position comparison (:262-267)."""
    classes = sorted(list(pre_extracted_feats.keys()), key=lambda x: int(x))
    dup = defaultdict(set)
    fractions = []
    for cls in classes:
        files = pre_extracted_feats[cls]["file_paths"]
        if files is None:
            continue
        to_remove = set(near_duplicate_positions(pre_extracted_feats[cls]["feats"], threshold).tolist())
        for pos, f in enumerate(files):
            key = pos if positional else int(f.split("/")[-1].split(".")[0])
            if key in to_remove:
                dup[cls].add(f)
        fractions.append(len(to_remove) / len(files))
    return dup, fractions, sum(fractions) / len(fractions)


def zeroshot_clip_img_filter(pre_extracted_feats: dict, head_weight: np.ndarray, classes=None, positional: bool = False):
    """``zeroshot_clip_img_filter`` (:278-329): rows whose zero-shot prediction ``argmax(feats @ W^T)`` (:299-300) is not
    their own class id; same file-id-vs-position comparison as the dedup (:315-318)."""
    if classes is None:
        classes = sorted(list(pre_extracted_feats.keys()), key=lambda x: int(x))
    W = np.asarray(head_weight, dtype=np.float32)
    out = defaultdict(set)
    fractions = []
    for cls in classes:
        files = pre_extracted_feats[cls]["file_paths"]
        logits = np.asarray(pre_extracted_feats[cls]["feats"], dtype=np.float32) @ W.T
        preds = np.argmax(logits, axis=1)
        to_remove = set(np.nonzero(preds != int(cls))[0].tolist())
        for pos, f in enumerate(files):
            key = pos if positional else int(f.split("/")[-1].split(".")[0])
            if key in to_remove:
                out[cls].add(f)
        fractions.append((len(files) - len(to_remove)) / len(files))
    return out, fractions


def verbatim_random_sampler(prompt_tensors, num_samples, threshold, pre_extracted_feats, duplicates_dict=None,
                            filtered_images_dict=None, tail_head=False, rng=None):
    """``random_sampler`` (:592-661): per class shuffle (Python ``random``), then the accept walk of ``add_to_split``
    (:439-482) with ``similarity`` = 1.0, or the T2I score when ``threshold != 0`` (:622-627).  Returns
    ``(file_list per class, counts, sampled strings, filtered strings)`` -- strings without captions."""
    import random as _random
    rng = rng or _random
    dups = duplicates_dict or defaultdict(set)
    filt = filtered_images_dict or defaultdict(set)
    classes = sorted(list(pre_extracted_feats.keys()), key=lambda x: int(x))
    files_out, counts, sampled, filtered = [], {}, [], []
    for cls in classes:
        file_list = pre_extracted_feats[cls]["file_paths"]
        if file_list is None:
            counts[cls] = 0
            continue
        sim = [1.0] * len(file_list)
        if threshold != 0:
            sim = similarity(np.asarray(prompt_tensors[cls]["mean"], dtype=np.float32)[None, :],
                             np.asarray(pre_extracted_feats[cls]["feats"], dtype=np.float32))
            sim = [sim] if isinstance(sim, float) else list(sim)
        zipped = list(zip(file_list, sim))
        rng.shuffle(zipped)
        if tail_head:
            if len(zipped) >= num_samples:
                threshold = 0                      # :638-639 -- sticks for every later class
        ct, acc = 0, []
        for fp, s_ in zipped:
            if ct == num_samples:
                break
            if s_ >= threshold and fp not in dups[str(cls)] and fp not in filt[str(cls)]:
                acc.append(fp); ct += 1
                sampled.append(f"{round(s_, 4)}/{threshold}, {fp}")
            else:
                filtered.append(f"{round(s_, 4)}/{threshold}, {fp}")
        if acc:
            files_out.append(acc)
        counts[cls] = ct
    return files_out, counts, sampled, filtered


# --------------------------------------------------------------------------------------
# accept / walk
# --------------------------------------------------------------------------------------
def check_caption(caption_map: dict, img_path: str) -> str:
    """``check_caption`` (:485-490)."""
    img_cls = img_path.split("/")[-2]
    img_id = img_path.split("/")[-1].split(".")[0]
    return caption_map[img_cls][img_id]


def walk_t2t(sorted_items: Sequence[Tuple[str, float, int]], cls: int, num_samples: int, threshold: float,
             duplicates: Optional[set] = None, filtered: Optional[set] = None,
             caption_map: Optional[dict] = None,
             filtered_list: Optional[list] = None, sampled_list: Optional[list] = None) -> List[int]:
    """``add_to_split`` (:439-482).  ``sorted_items`` = (path, t2t, row) in walk order.

    Stops *before* looking at the next item once ``num_samples`` were accepted (:451-452);
    accepts iff ``similarity >= threshold`` and path not in the two exclusion sets (:454-456).
    Returns accepted rows in walk order.
    """
    duplicates = duplicates or set()
    filtered = filtered or set()
    accepted: List[int] = []
    for path, sim, row in sorted_items:
        if len(accepted) == num_samples:
            break
        ok = (sim >= threshold) and (path not in duplicates) and (path not in filtered)
        if caption_map is not None:
            info = f"{round(sim, 4)}/{threshold}, {path}, {check_caption(caption_map, path)}"   # :463-469
            (sampled_list if ok else filtered_list).append(info)
        if ok:
            accepted.append(row)
    return accepted


def walk_t2t_t2i(sorted_items: Sequence[Tuple[str, float, int, float]], cls: int, num_samples: int,
                 threshold: float, t2i_threshold: float = 0.25,
                 duplicates: Optional[set] = None, filtered: Optional[set] = None,
                 caption_map: Optional[dict] = None,
                 filtered_list: Optional[list] = None, sampled_list: Optional[list] = None) -> List[int]:
    """``add_t2t_ranked_t2i_tshd_to_split`` (:492-540): same walk, predicate also needs
    ``t2i_sim >= t2i_threshold`` (:511-514)."""
    duplicates = duplicates or set()
    filtered = filtered or set()
    accepted: List[int] = []
    for path, sim, row, t2i in sorted_items:
        if len(accepted) == num_samples:
            break
        ok = (sim >= threshold) and (t2i >= t2i_threshold) and (path not in duplicates) and (path not in filtered)
        if caption_map is not None:
            info = (f"{round(sim, 4)}/{threshold}, {round(t2i, 4)}/{t2i_threshold}, {path}, "
                    f"{check_caption(caption_map, path)}")                                      # :519-525
            (sampled_list if ok else filtered_list).append(info)
        if ok:
            accepted.append(row)
    return accepted


# --------------------------------------------------------------------------------------
# samplers, reference-verbatim structure
# --------------------------------------------------------------------------------------
def verbatim_t2t_ranked_sampler(prompt_tensors: dict, num_samples: int, threshold: float,
                                pre_extracted_feats: dict,
                                duplicates_dict: Optional[dict] = None,
                                filtered_images_dict: Optional[dict] = None,
                                caption_map: Optional[dict] = None,
                                classes: Optional[Iterable[str]] = None,
                                rank_on_images: bool = False, rank_fewshot: Optional[dict] = None):
    """``t2t_ranked_sampler`` (:724-771).

    Per class in ascending int order (:734-735): GEMV scores (:752), Python ``sorted`` on the
    zipped tuples with ``reverse=True`` (:754; stable, so equal scores stay in ascending row
    order), then the accept walk.  Returns ``(mined_split, num_imgs_sampled_dict, diag)`` where
    ``mined_split`` has the reference's three lists plus ``row_list``/``score_list`` (per class,
    indices into that class's regrouped rows) and ``diag`` holds the two diagnostic string lists.
    """
    duplicates_dict = duplicates_dict if duplicates_dict is not None else defaultdict(set)
    filtered_images_dict = filtered_images_dict if filtered_images_dict is not None else defaultdict(set)
    classes = sorted(list(pre_extracted_feats.keys()) if classes is None else list(classes), key=lambda x: int(x))
    mined_split = {"feature_list": [], "label_list": [], "file_list": [], "row_list": [], "score_list": []}
    num_imgs_sampled_dict = {}
    filtered_list: List[str] = []
    sampled_list: List[str] = []
    for cls in classes:
        file_list = pre_extracted_feats[cls]["file_paths"]
        if file_list is None:
            continue
        img_embeddings = pre_extracted_feats[cls]["feats"]
        caption_embeddings = pre_extracted_feats[cls]["caption_feats"]
        class_prompt = np.asarray(prompt_tensors[cls]["mean"], dtype=np.float32)[None, :]       # :749-750
        # t2i_ranked_sampler (:1195-1243) is this function with cal_t2i_similarity on the image features (:1224);
        # i2i_ranked_sampler_p2p (:1016-1076) / i2t_rank_sampler (:1079-1133) rank by the mean over the few-shot vectors
        if rank_fewshot is not None:
            sim = similarity_p2p(np.stack(rank_fewshot[int(cls)]), img_embeddings if rank_on_images else caption_embeddings, "mean")
        else:
            sim = similarity(class_prompt, img_embeddings if rank_on_images else caption_embeddings)
        embedding_list = [img_embeddings[i] for i in range(len(img_embeddings))]                 # :753 (N row views)
        items = sorted(list(zip(file_list, sim, range(len(file_list)), embedding_list)), key=lambda x: x[1], reverse=True)
        items = [(p, s_, r) for p, s_, r, _ in items]
        acc = walk_t2t(items, int(cls), num_samples, threshold,
                       duplicates_dict.get(str(int(cls)), set()) if isinstance(duplicates_dict, dict) else set(),
                       filtered_images_dict.get(str(int(cls)), set()) if isinstance(filtered_images_dict, dict) else set(),
                       caption_map, filtered_list, sampled_list)
        num_imgs_sampled_dict[cls] = len(acc)
        if acc:                                                                                  # :472-480
            mined_split["feature_list"].append(np.stack([img_embeddings[i] for i in acc]))
            mined_split["label_list"].append(np.full(len(acc), int(cls), dtype=np.int64))
            mined_split["file_list"].append([file_list[i] for i in acc])
            mined_split["row_list"].append(np.asarray(acc, dtype=np.int64))
            mined_split["score_list"].append(np.asarray([sim[i] for i in acc], dtype=np.float32))
    return mined_split, num_imgs_sampled_dict, {"filtered_list": filtered_list, "sampled_list": sampled_list}


def verbatim_t2t_ranked_t2i_tshd_sampler(prompt_tensors: dict, num_samples: int, threshold: float,
                                         pre_extracted_feats: dict,
                                         duplicates_dict: Optional[dict] = None,
                                         filtered_images_dict: Optional[dict] = None,
                                         caption_map: Optional[dict] = None,
                                         t2i_threshold: float = 0.25,
                                         classes: Optional[Iterable[str]] = None,
                                         pred_fewshot: Optional[dict] = None, pred_on_captions: bool = False):
    """``t2t_ranked_t2i_tshd_sampler`` (:774-825): as above plus T2I scores (:806); tuples are
    sorted by T2T only (:807-808) and walked with the two-threshold predicate."""
    duplicates_dict = duplicates_dict if duplicates_dict is not None else defaultdict(set)
    filtered_images_dict = filtered_images_dict if filtered_images_dict is not None else defaultdict(set)
    classes = sorted(list(pre_extracted_feats.keys()) if classes is None else list(classes), key=lambda x: int(x))
    mined_split = {"feature_list": [], "label_list": [], "file_list": [], "row_list": [], "score_list": [],
                   "t2i_list": []}
    num_imgs_sampled_dict = {}
    filtered_list: List[str] = []
    sampled_list: List[str] = []
    for cls in classes:
        file_list = pre_extracted_feats[cls]["file_paths"]
        if file_list is None:
            continue
        img_embeddings = pre_extracted_feats[cls]["feats"]
        caption_embeddings = pre_extracted_feats[cls]["caption_feats"]
        class_prompt = np.asarray(prompt_tensors[cls]["mean"], dtype=np.float32)[None, :]
        sim = similarity(class_prompt, caption_embeddings)
        embedding_list = [img_embeddings[i] for i in range(len(img_embeddings))]                 # :805 (N row views)
        if pred_fewshot is not None:       # t2t_rank_i2t_tshd_sampler :869 (captions, 0.25) / t2t_rank_i2i_tshd_sampler :929 (images, 0.65)
            t2i = similarity_p2p(np.stack(pred_fewshot[int(cls)]), caption_embeddings if pred_on_captions else img_embeddings, "max")
        else:
            t2i = similarity(class_prompt, img_embeddings)
        items = sorted(list(zip(file_list, sim, range(len(file_list)), t2i, embedding_list)), key=lambda x: x[1], reverse=True)
        items = [(p, s_, r, t_) for p, s_, r, t_, _ in items]
        acc = walk_t2t_t2i(items, int(cls), num_samples, threshold, t2i_threshold,
                           duplicates_dict.get(str(int(cls)), set()) if isinstance(duplicates_dict, dict) else set(),
                           filtered_images_dict.get(str(int(cls)), set()) if isinstance(filtered_images_dict, dict) else set(),
                           caption_map, filtered_list, sampled_list)
        num_imgs_sampled_dict[cls] = len(acc)
        if acc:
            mined_split["feature_list"].append(np.stack([img_embeddings[i] for i in acc]))
            mined_split["label_list"].append(np.full(len(acc), int(cls), dtype=np.int64))
            mined_split["file_list"].append([file_list[i] for i in acc])
            mined_split["row_list"].append(np.asarray(acc, dtype=np.int64))
            mined_split["score_list"].append(np.asarray([sim[i] for i in acc], dtype=np.float32))
            mined_split["t2i_list"].append(np.asarray([t2i[i] for i in acc], dtype=np.float32))
    return mined_split, num_imgs_sampled_dict, {"filtered_list": filtered_list, "sampled_list": sampled_list}


def format_split_lines(file_lists: Sequence[Sequence[str]], label_lists: Sequence[np.ndarray]) -> List[str]:
    """``save_sample_file_list`` (:1457-1462) line format: ``"<path> <label> 0\\n"``, class-major."""
    lines = []
    for files, labels in zip(file_lists, label_lists):
        for p, l in zip(files, np.asarray(labels).tolist()):
            lines.append(f"{p} {l} {0}\n")
    return lines


# --------------------------------------------------------------------------------------
# vectorised restatement (fast checker for >= 1 M rows)
# --------------------------------------------------------------------------------------
def score_matrix(bank: np.ndarray, queries: np.ndarray, class_of_query: Optional[np.ndarray] = None,
                 n_classes: Optional[int] = None, reduce="none") -> np.ndarray:
    """``[N, C]`` fp32 class scores: ``bank @ Q^T`` then the per-class reduce over the query
    columns of each class (none: :749-752; mean: :403-404; min/max/mean: :377-385)."""
    r = reduce_code(reduce)
    s = np.asarray(bank, dtype=np.float32) @ np.asarray(queries, dtype=np.float32).T
    if r == REDUCE_NONE:
        return s
    coq = np.asarray(class_of_query)
    C = int(n_classes if n_classes is not None else coq.max() + 1)
    out = np.empty((s.shape[0], C), dtype=np.float32)
    for c in range(C):
        cols = np.nonzero(coq == c)[0]
        if cols.size == 0:
            out[:, c] = -np.inf
        elif r == REDUCE_MEAN:
            out[:, c] = s[:, cols].mean(axis=-1, dtype=np.float32)
        elif r == REDUCE_MAX:
            out[:, c] = s[:, cols].max(axis=-1)
        else:
            out[:, c] = s[:, cols].min(axis=-1)
    return out


def select_walk(t2t: np.ndarray, k: int, threshold: float,
                t2i: Optional[np.ndarray] = None, t2i_threshold: float = 0.25,
                eligible: Optional[np.ndarray] = None) -> np.ndarray:
    """Rows one class accepts, in walk order, without materialising the sort.

    The reference sorts every row by T2T descending with a stable sort (:754, :807) and accepts
    rows passing the predicate until ``k`` are taken (:451-456, :507-514).  That is: among the
    predicate-passing rows take the ``k`` largest by (T2T descending, row index ascending).
    """
    t2t = np.asarray(t2t, dtype=np.float32)
    ok = t2t >= np.float32(threshold) if not np.isneginf(threshold) else np.ones_like(t2t, dtype=bool)
    if t2i is not None:
        ok &= np.asarray(t2i, dtype=np.float32) >= np.float32(t2i_threshold)
    if eligible is not None:
        ok &= eligible
    idx = np.nonzero(ok)[0]
    if idx.size == 0 or k <= 0:
        return np.empty(0, dtype=np.int64)
    sc = t2t[idx]
    if idx.size > k:
        kth = np.partition(sc, idx.size - k)[idx.size - k]      # k-th largest value
        keep = sc >= kth                                        # keeps every tie at the boundary
        idx, sc = idx[keep], sc[keep]
    order = np.lexsort((idx, -sc.astype(np.float64)))           # score desc, row asc
    return idx[order][:k].astype(np.int64)


def topk_walk(t2t_bank: np.ndarray, queries: np.ndarray, k: int, threshold: float = 0.0,
              t2i_bank: Optional[np.ndarray] = None, t2i_threshold: float = 0.25,
              class_of_query: Optional[np.ndarray] = None, n_classes: Optional[int] = None, reduce="none",
              row_labels: Optional[np.ndarray] = None, exclude: Optional[np.ndarray] = None,
              row_chunk: int = 1 << 18):
    """Unpartitioned generalisation: every class scans the whole bank, or, with
    ``row_labels`` (dense class index per row, -1 = none), only its own rows (the reference's
    partitioned case).  Returns ``rows [C,k] int64 (-1 padded), t2t [C,k], t2i [C,k] or None,
    counts [C]``; order inside a class is the reference's walk order.
    """
    r = reduce_code(reduce)
    Q = np.asarray(queries, dtype=np.float32)
    if class_of_query is None:
        class_of_query = np.arange(Q.shape[0])
    C = int(n_classes if n_classes is not None else np.max(class_of_query) + 1)
    N = t2t_bank.shape[0]
    # running candidate pools per class keep memory bounded for large N
    pools: List[List[Tuple[np.ndarray, np.ndarray, Optional[np.ndarray]]]] = [[] for _ in range(C)]
    for s0 in range(0, N, row_chunk):
        s1 = min(N, s0 + row_chunk)
        S = score_matrix(np.asarray(t2t_bank[s0:s1], dtype=np.float32), Q, class_of_query, C, r)
        I = None
        if t2i_bank is not None:
            I = score_matrix(np.asarray(t2i_bank[s0:s1], dtype=np.float32), Q, class_of_query, C, r)
        ex = None if exclude is None else ~np.asarray(exclude[s0:s1], dtype=bool)
        lab = None if row_labels is None else np.asarray(row_labels[s0:s1])
        for c in range(C):
            el = ex
            if lab is not None:
                m = lab == c
                el = m if el is None else (el & m)
            sel = select_walk(S[:, c], k, threshold, None if I is None else I[:, c], t2i_threshold, el)
            if sel.size:
                pools[c].append((sel + s0, S[sel, c], None if I is None else I[sel, c]))
    rows = np.full((C, k), -1, dtype=np.int64)
    out_s = np.zeros((C, k), dtype=np.float32)
    out_i = None if t2i_bank is None else np.zeros((C, k), dtype=np.float32)
    counts = np.zeros(C, dtype=np.int32)
    for c in range(C):
        if not pools[c]:
            continue
        idx = np.concatenate([p[0] for p in pools[c]])
        sc = np.concatenate([p[1] for p in pools[c]])
        order = np.lexsort((idx, -sc.astype(np.float64)))[:k]
        n = order.size
        rows[c, :n] = idx[order]
        out_s[c, :n] = sc[order]
        if out_i is not None:
            ti = np.concatenate([p[2] for p in pools[c]])
            out_i[c, :n] = ti[order]
        counts[c] = n
    return rows, out_s, out_i, counts


def merge_topk(rows: np.ndarray, scores: np.ndarray, counts: np.ndarray, k: int):
    """Merge per-shard results ``[G,C,k]`` into ``[C,k]`` under (score desc, global row asc) --
    the associative top-k that SURVEY.md 8(e) shards on."""
    G, C, _ = rows.shape
    out_r = np.full((C, k), -1, dtype=np.int64)
    out_s = np.zeros((C, k), dtype=np.float32)
    out_c = np.zeros(C, dtype=np.int32)
    for c in range(C):
        idx = np.concatenate([rows[g, c, :counts[g, c]] for g in range(G)])
        sc = np.concatenate([scores[g, c, :counts[g, c]] for g in range(G)])
        order = np.lexsort((idx, -sc.astype(np.float64)))[:k]
        out_r[c, :order.size] = idx[order]
        out_s[c, :order.size] = sc[order]
        out_c[c] = order.size
    return out_r, out_s, out_c
