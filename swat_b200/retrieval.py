"""Host-side mirror of the reference's retrieval hot path, same names / arguments / return
structures as ``/root/reference/retrieval/sample_retrieval.py`` (file:line below), backed by the
C-ABI CUDA library.  A SWAT maintainer switches with::

    from swat_b200.retrieval import (t2t_similarity, cal_t2i_similarity, transform_extracted_fea,
                                     t2t_ranked_sampler, t2t_ranked_t2i_tshd_sampler)

What changes underneath: instead of one GEMV + Python ``sorted`` + walk per class, all classes are
scored against the whole bank in one tcgen05 scan with the selection fused in, and only the
<= C*k winners ever come back to the host.  No CPU fallback: without the library or a GPU every
function raises.
"""
from __future__ import annotations

import json
import os
import pickle
import shutil
from collections import defaultdict
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib

_CTX: Dict[int, "_lib.Context"] = {}


def get_context(device: int = 0) -> "_lib.Context":
    if device not in _CTX:
        _CTX[device] = _lib.Context(device)
    return _CTX[device]


# ---------------------------------------------------------------------------------------------
# S1 primitives (kept for drop-in compatibility; they materialise N scores, the samplers do not)
# ---------------------------------------------------------------------------------------------
def _dense(class_prompt: torch.Tensor, embeddings: torch.Tensor, reduce: str) -> List[float]:
    ctx = get_context()
    q = torch.as_tensor(class_prompt).detach().float()
    if q.dim() == 1:
        q = q[None, :]
    x = torch.as_tensor(embeddings).detach()
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    if x.dim() == 1:
        x = x[None, :]
    R = q.shape[0]
    qs = _lib.Queries(ctx, q, np.zeros(R, np.int32), 1, reduce if R > 1 else "none")
    s = _lib.scores_dense(ctx, qs, x.contiguous().cuda(ctx.device))
    result = s.squeeze().cpu().tolist()
    if isinstance(result, float):          # single row -> 1-element list (:413-414)
        result = [result]
    qs.close()
    return result


def t2t_similarity(class_prompt, caption_embeddings) -> List[float]:
    """``t2t_similarity`` (:397-416): ``X @ q^T``, mean over prompt rows when R > 1."""
    return _dense(class_prompt, caption_embeddings, "mean")


def cal_t2i_similarity(class_prompt, img_embeddings) -> List[float]:
    """``cal_t2i_similarity`` (:335-353)."""
    return _dense(class_prompt, img_embeddings, "mean")


def i2i_similarity_p2p(fewshot_embedding, img_embeddings, mode: str) -> List[float]:
    """``i2i_similarity_p2p`` (:369-394): min / max / mean over the few-shot columns."""
    if mode not in ("min", "max", "mean"):
        raise ValueError("Invalid mode type.")
    f = torch.from_numpy(np.stack(fewshot_embedding)) if not torch.is_tensor(fewshot_embedding) else fewshot_embedding
    return _dense(f, img_embeddings, mode)


# ---------------------------------------------------------------------------------------------
# loader / regrouper
# ---------------------------------------------------------------------------------------------
class RegroupedFeats(dict):
    """What ``transform_extracted_fea`` returns: ``{str(label): {'file_paths', 'feats',
    'caption_feats'}}`` with keys in first-appearance order -- built lazily per class -- plus the
    flat row-major tensors the fused sampler consumes directly (``flat``)."""

    def __init__(self, raw: dict):
        super().__init__()
        self.flat = raw
        labels = torch.as_tensor(raw["labels"]).cpu().long()
        self.labels = labels
        order = torch.argsort(labels, stable=True)
        sorted_labels = labels[order]
        uniq, counts = torch.unique_consecutive(sorted_labels, return_counts=True)
        ends = torch.cumsum(counts, 0)
        starts = ends - counts
        first_row = order[starts]                       # first appearance of each label (stable sort)
        appear = torch.argsort(first_row)
        self._rows = {}
        for i in appear.tolist():                       # first-appearance order (:1401-1402)
            key = str(int(uniq[i]))
            self._rows[key] = order[starts[i]:ends[i]]
            dict.__setitem__(self, key, None)

    def rows_of(self, key: str) -> torch.Tensor:
        return self._rows[key]

    def __getitem__(self, key):
        v = dict.__getitem__(self, key)
        if v is None:
            rows = self._rows[key]
            paths = self.flat["filepath"]
            v = {"file_paths": [paths[i] for i in rows.tolist()],
                 "feats": torch.as_tensor(self.flat["image_features"])[rows],
                 "caption_feats": torch.as_tensor(self.flat["caption_features"])[rows]}
            dict.__setitem__(self, key, v)
        return v

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def values(self):
        return [self[k] for k in self.keys()]


def transform_extracted_fea(pre_extracted_feats: dict) -> RegroupedFeats:
    """``transform_extracted_fea`` (:1387-1415) without the per-row Python loop: one stable argsort
    of the labels; per-class entries are materialised only when somebody indexes them."""
    out = RegroupedFeats(pre_extracted_feats)
    print("len(collection_dict):", len(out))             # the reference prints this (:1408)
    return out


# ---------------------------------------------------------------------------------------------
# exclusion-set producer
# ---------------------------------------------------------------------------------------------
def remove_near_duplicates2(pre_extracted_feats, threshold: float = 0.9, positional: bool = False, device: int = 0):
    """``remove_near_duplicates2`` (:237-275): per class, image rows whose cosine to an EARLIER row of the
    class exceeds 0.9 are duplicates.  Returns ``(duplicate_images_dict, dup_images_fraction,
    avg_dup_images_fraction)`` like the reference.

    The reference then keeps the files whose *file id* (``<id>.jpg``) equals one of the duplicate
    *positions* (:262-267) -- it compares names with indices.  That is reproduced by default because the
    resulting ``duplicates_dict`` is what its samplers consume; ``positional=True`` marks the duplicate
    rows themselves."""
    ctx = get_context(device)
    classes = sorted(list(pre_extracted_feats.keys()), key=lambda x: int(x))
    if isinstance(pre_extracted_feats, RegroupedFeats):
        img = torch.as_tensor(pre_extracted_feats.flat["image_features"])
        rows = [pre_extracted_feats.rows_of(c) for c in classes]
        files = [[pre_extracted_feats.flat["filepath"][i] for i in r.tolist()] for r in rows]
        order = torch.cat(rows) if rows else torch.zeros(0, dtype=torch.int64)
    else:
        feats = [torch.as_tensor(pre_extracted_feats[c]["feats"]) for c in classes]
        files = [pre_extracted_feats[c]["file_paths"] for c in classes]
        img = torch.cat(feats)
        order = torch.arange(img.shape[0], dtype=torch.int64)
    if img.dtype not in (torch.float32, torch.bfloat16):
        img = img.float()
    sizes = torch.tensor([len(f) for f in files], dtype=torch.int64)
    class_start = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(sizes, 0)])
    dup = _lib.near_duplicates(ctx, img.contiguous().cuda(device), order, class_start, threshold).cpu().numpy().astype(bool)
    duplicate_images_dict = defaultdict(set)
    dup_images_fraction = []
    for ci, cls in enumerate(classes):
        s0, s1 = int(class_start[ci]), int(class_start[ci + 1])
        if s1 == s0:
            continue
        to_remove = set(np.nonzero(dup[s0:s1])[0].tolist())                  # positions inside the class (:259)
        for pos, f in enumerate(files[ci]):
            key = pos if positional else int(f.split('/')[-1].split('.')[0])   # :264-266
            if key in to_remove:
                duplicate_images_dict[cls].add(f)
        dup_images_fraction.append(len(to_remove) / len(files[ci]))
    avg = sum(dup_images_fraction) / len(dup_images_fraction) if dup_images_fraction else 0.0
    return duplicate_images_dict, dup_images_fraction, avg


# ---------------------------------------------------------------------------------------------
# samplers
# ---------------------------------------------------------------------------------------------
def _head_weight(head) -> torch.Tensor:
    """``[C,512]`` weights of the zero-shot head: ``MyLinear`` (``.linear.weight``, utils/models.py:47-68), an
    ``nn.Linear`` or a plain tensor.  The reference builds it with ``bias=False`` (:1490)."""
    if torch.is_tensor(head):
        return head.detach().float().cpu()
    lin = getattr(head, "linear", head)
    bias = getattr(lin, "bias", None)
    if bias is not None and bool((bias.detach() != 0).any()):
        raise NotImplementedError("zero-shot head with a non-zero bias")
    return lin.weight.detach().float().cpu()


def zeroshot_clip_img_filter(model=None, preprocess=None, root_folder=None, pre_extracted_feats=None, head=None,
                             positional: bool = False, device: int = 0):
    """``zeroshot_clip_img_filter`` (:278-329): classify every mined image with the zero-shot head and collect the
    ones not predicted as their own class into ``filtered_images_dict``.  The logits ``feats @ W^T`` (:299) and the
    argmax (:300) run on the GPU for all classes at once (``swat_zeroshot_predict``).

    Like ``remove_near_duplicates2`` the reference compares the integer *file id* of every file with the set of
    mispredicted *positions* (:315-318); that is reproduced unless ``positional=True``.  ``model`` and ``preprocess``
    are unused (as in the reference, which requires pre-extracted features, :293-297)."""
    if pre_extracted_feats is None:
        raise ValueError("Pre-extracted features are required for zeroshot filtering.")        # :297
    if root_folder is not None and os.path.isdir(root_folder):
        classes = [c for c in os.listdir(root_folder) if os.path.isdir(os.path.join(root_folder, c))]   # :282-284
    else:
        classes = list(pre_extracted_feats.keys())
    classes = sorted(classes, key=lambda x: int(x))                                              # :285
    W = _head_weight(head)
    _, img, paths, row_class = _flatten(pre_extracted_feats, classes)
    ctx = get_context(device)
    qs = _lib.Queries(ctx, W)
    if img.dtype not in (torch.float32, torch.bfloat16):
        img = img.float()
    pred = _lib.zeroshot_predict(ctx, qs, img.contiguous().cuda(device)).cpu()
    qs.close()
    filtered_images_dict = defaultdict(set)
    fractions = []
    for ci, cls in enumerate(classes):
        files = pre_extracted_feats[cls]["file_paths"]
        n = len(files)
        p = pred if row_class is None else pred[row_class == ci]      # file order inside the class is kept
        to_remove = set((p != int(cls)).nonzero().flatten().tolist())                            # :303-309
        for pos, f in enumerate(files):
            key = pos if positional else int(f.split("/")[-1].split(".")[0])                     # :315-318
            if key in to_remove:
                filtered_images_dict[cls].add(f)
        fractions.append((n - len(to_remove)) / n)
    print(f"Average unique images: {round(sum(fractions) / len(fractions), 4)}")                 # :326
    return filtered_images_dict


def _flatten(pre_extracted_feats, classes: List[str]):
    """Return (caption [N,512], image [N,512], paths, row_class int32 [N] or None)."""
    if isinstance(pre_extracted_feats, RegroupedFeats):
        raw = pre_extracted_feats.flat
        labels = pre_extracted_feats.labels
        lut = {int(c): i for i, c in enumerate(classes)}
        uniq = torch.unique(labels)
        table = torch.full((int(uniq.max()) + 1 if uniq.numel() else 1,), -1, dtype=torch.int32)
        neg = uniq[uniq < 0]
        if neg.numel():
            raise ValueError("negative labels are not supported")
        for u in uniq.tolist():
            table[u] = lut.get(int(u), -1)
        return (torch.as_tensor(raw["caption_features"]), torch.as_tensor(raw["image_features"]), raw["filepath"],
                table[labels].contiguous())
    first = pre_extracted_feats[classes[0]]
    aliased = all(pre_extracted_feats[c]["caption_feats"] is first["caption_feats"] and
                  pre_extracted_feats[c]["file_paths"] is first["file_paths"] for c in classes)
    if aliased:                                           # unpartitioned: every class scans the whole bank
        return first["caption_feats"], first["feats"], first["file_paths"], None
    caps, imgs, paths, rc = [], [], [], []
    for i, c in enumerate(classes):                        # the reference's regrouped dict: concatenate
        e = pre_extracted_feats[c]
        if e["file_paths"] is None:
            continue
        caps.append(torch.as_tensor(e["caption_feats"])); imgs.append(torch.as_tensor(e["feats"]))
        paths.extend(e["file_paths"])
        rc.append(torch.full((len(e["file_paths"]),), i, dtype=torch.int32))
    return torch.cat(caps), torch.cat(imgs), paths, torch.cat(rc)


def _exclusion_sets(duplicates_dict, filtered_images_dict):
    sets = {}
    for d in (duplicates_dict, filtered_images_dict):
        if d:
            for k, v in d.items():
                if v:
                    sets.setdefault(str(k), set()).update(v)
    return sets


def _exclusion_bitmap(paths, classes, row_class, sets):
    """Row-level bitmap for the partitioned layout (a row belongs to one class, so "excluded for its class" is a row
    property): bit set = never accept (``add_to_split`` :454-456)."""
    index = {p: i for i, p in enumerate(paths)}
    ex = np.zeros(len(paths), dtype=bool)
    for k, v in sets.items():
        for p in v:
            i = index.get(p)
            if i is not None and classes[int(row_class[i])] == k:
                ex[i] = True
    bits = np.packbits(ex, bitorder="little")
    bits = np.concatenate([bits, np.zeros((-len(bits)) % 4, np.uint8)]).view(np.int32)
    return torch.from_numpy(bits.copy()), ex


def check_caption(caption_map, img_path):
    """``check_caption`` (:485-490)."""
    img_cls = img_path.split('/')[-2]
    img_id = img_path.split('/')[-1].split('.')[0]
    return caption_map[img_cls][img_id]


def _fewshot_queries(fewshot_fea, classes):
    """``fewshot_fea[int(cls)]`` = list of few-shot image embeddings (``get_fewshot_features`` :997-1014)
    -> stacked query rows, class_of_query."""
    rows, coq = [], []
    for i, c in enumerate(classes):
        f = fewshot_fea[int(c)]
        f = torch.stack([torch.as_tensor(x).detach().float().cpu().reshape(-1) for x in f])
        rows.append(f)
        coq.extend([i] * f.shape[0])
    return torch.cat(rows), torch.tensor(coq, dtype=torch.int32)


def _load_caption_map(args):
    cmap_path = getattr(args, "caption_map_path", None)
    if cmap_path is None:
        try:
            from .config import CAPTION_MAP_DICT
            cmap_path = CAPTION_MAP_DICT.get(args.dataset)
        except Exception:
            cmap_path = None
    if cmap_path and os.path.exists(cmap_path):
        with open(cmap_path, "rb") as f:                                              # :730-732
            return pickle.load(f)
    return None


def _score_own_class(ctx, qs, bank, row_class, device, chunk=1 << 20):
    """Per-row canonical score against the row's own class (``swat_score_rows``); host banks go through the device in chunks."""
    if bank.is_cuda:
        return _lib.score_rows(ctx, qs, bank.contiguous(), row_class.to(bank.device)).cpu()
    out = []
    for s0 in range(0, bank.shape[0], chunk):
        out.append(_lib.score_rows(ctx, qs, bank[s0:s0 + chunk].contiguous().cuda(device), row_class[s0:s0 + chunk].cuda(device)).cpu())
    return torch.cat(out) if out else torch.zeros(0)


def _filtered_lines(classes, row_class, paths, t2t_all, pred_all, excluded, num_samples, threshold, t2i_threshold, caption_map):
    """The reference's ``filtered_list`` (:463-469, :521-527) for partitioned data: per class, walk the class's rows in
    (score desc, row asc) order until ``num_samples`` are accepted and report every rejected row met on the way --
    all remaining rows when the class never fills."""
    rc = row_class.numpy()
    order = np.argsort(rc, kind="stable")
    starts = np.searchsorted(rc[order], np.arange(len(classes) + 1))
    t2t_np = t2t_all.numpy()
    pred_np = None if pred_all is None else pred_all.numpy()
    lines: List[str] = []
    for i, _cls in enumerate(classes):
        rows = order[starts[i]:starts[i + 1]]
        if rows.size == 0:
            continue
        s = t2t_np[rows]
        o = np.lexsort((rows, -s.astype(np.float64)))
        rows, s = rows[o], s[o]
        ok = s >= threshold
        if pred_np is not None:
            ok &= pred_np[rows] >= t2i_threshold
        if excluded is not None:
            ok &= ~excluded[rows]
        filled = np.nonzero(np.cumsum(ok) == num_samples)[0]
        end = int(filled[0]) + 1 if filled.size else rows.size
        for j in np.nonzero(~ok[:end])[0].tolist():
            r = int(rows[j]); p = paths[r]
            caption = check_caption(caption_map, p) if caption_map is not None else ""
            if pred_np is not None:
                lines.append(f"{round(float(s[j]), 4)}/{threshold}, {round(float(pred_np[r]), 4)}/{t2i_threshold}, {p}, {caption}")
            else:
                lines.append(f"{round(float(s[j]), 4)}/{threshold}, {p}, {caption}")
    return lines


def _run_sampler(args, logger, prompt_tensors, num_samples, threshold, pre_extracted_feats, duplicates_dict,
                 filtered_images_dict, with_t2i: bool, t2i_threshold: float = 0.25, rank_on_images: bool = False,
                 rank_fewshot=None, pred_fewshot=None, pred_on_captions: bool = False, file_prefix: Optional[bool] = None):
    """One engine for every ranked sampler of the reference:

    rank stage   -- class prompt (``['mean']``) or the mean over a class's few-shot vectors (``rank_fewshot``),
                    against the caption bank or the image bank;
    predicate    -- none, the same prompt against the image bank (T2T-rank-T2I-tshd), or the max over the
                    few-shot vectors (``pred_fewshot``) against the image / caption bank, ``>= t2i_threshold``.
    """
    classes = sorted(list(pre_extracted_feats.keys()), key=lambda x: int(x))          # :734-735
    empty = set()
    if not isinstance(pre_extracted_feats, RegroupedFeats):
        empty = {c for c in classes if pre_extracted_feats[c]["file_paths"] is None}   # :743-745: skipped, no dict entry
    cap, img, paths, row_class = _flatten(pre_extracted_feats, classes)
    feat_bank = img                                          # feature_list always carries the image features (:753, :1225)
    bank_dtype = getattr(args, "bank_dtype", None)
    if bank_dtype in ("bf16", torch.bfloat16):
        cap = cap.to(torch.bfloat16); img = img.to(torch.bfloat16)
    elif cap.dtype not in (torch.float32, torch.bfloat16):
        cap = cap.float(); img = img.float()
    rank_bank = img if rank_on_images else cap               # T2I-rank / I2I-rank score the image bank (:1224, :1049)
    pred_bank = cap if pred_on_captions else img             # what the accept predicate is evaluated on
    device = int(getattr(args, "device_index", 0))
    ctx = get_context(device)
    if rank_fewshot is not None:                             # I2I-rank / I2T-rank: mean over the few-shot columns (:1049, :1112)
        fq, fcoq = _fewshot_queries(rank_fewshot, classes)
        qs = _lib.Queries(ctx, fq, fcoq, len(classes), "mean")
    else:
        q = torch.stack([torch.as_tensor(prompt_tensors[c]["mean"]).detach().float().cpu().reshape(-1) for c in classes])  # :749
        qs = _lib.Queries(ctx, q)
    sets = _exclusion_sets(duplicates_dict, filtered_images_dict)
    exclude, excluded_rows, k_run = None, None, int(num_samples)
    if sets and row_class is not None:
        exclude, excluded_rows = _exclusion_bitmap(paths, classes, row_class, sets)
    elif sets:
        # unpartitioned (every class scans the whole bank): a path excluded for one class stays eligible for the others,
        # so walk deep enough to cover the largest exclusion set and drop the excluded paths on the host
        k_run = num_samples + max(len(v) for v in sets.values())
        if k_run > 4096:
            raise NotImplementedError(f"per-class exclusion sets of up to {k_run - num_samples} files on an unpartitioned bank "
                                      "exceed the walk depth; regroup the features by label (transform_extracted_fea)")
    qs_pred = None
    if pred_fewshot is None:
        t2i_bank = img if with_t2i else None
        if rank_bank.is_cuda:
            res = _lib.topk(ctx, qs, rank_bank.contiguous(), k_run, threshold,
                            t2i_bank=None if t2i_bank is None else t2i_bank.contiguous(), t2i_threshold=t2i_threshold,
                            row_class=None if row_class is None else row_class.cuda(device),
                            exclude=None if exclude is None else exclude.cuda(device))
            scores, rows, t2i, counts = [None if x is None else x.cpu() for x in res]
        else:
            scores, rows, t2i, counts = _lib.topk_host(ctx, qs, rank_bank.contiguous(), k_run, threshold,
                                                      t2i_bank=None if t2i_bank is None else t2i_bank.contiguous(),
                                                      t2i_threshold=t2i_threshold, row_class=row_class, exclude=exclude)
    else:
        # predicate with its own query set: max over the class's few-shot vectors (:869, :929), on the
        # caption bank (I2T) or the image bank (I2I).  Composed from the job / walk entry points; the
        # over-fetch deepens until every class is provably exact (candidate list not truncated, or k accepted).
        pq, pcoq = _fewshot_queries(pred_fewshot, classes)
        qs_pred = _lib.Queries(ctx, pq, pcoq, len(classes), "max")
        d_rank = rank_bank.contiguous().cuda(device)
        d_pred = pred_bank.contiguous().cuda(device)
        d_rc = None if row_class is None else row_class.cuda(device)
        d_ex = None if exclude is None else exclude.cuda(device)
        eps = _lib.scan_eps(qs, d_rank.dtype)
        kf = min(4096, max(1024, 2 * k_run))
        while True:
            job = _lib.Job(ctx, qs, kf, threshold - eps)
            job.scan(d_rank, row_class=d_rc, exclude=d_ex)
            sc, rw, cn, tr = job.select()
            over = job.overflowed()
            job.close()
            if over:
                ctx.set_option("cand_cap", 8 * kf + 16384); ctx.set_option("list_entries", 64 << 20)
                continue
            o_s, o_r, o_t, o_c, _, o_i = _lib.rescore_walk(ctx, qs, d_rank, sc, rw, cn, tr, k_run, threshold, aux_bank=d_pred,
                                                          aux_threshold=t2i_threshold, eps=eps, aux_queries=qs_pred)
            if int(o_i.sum()) == 0:
                break
            if kf >= 4096:
                raise _lib.SwatError(-5, "few-shot predicate walk not provably exact at k_fetch=4096 "
                                         "(more than 4096 eligible rows in a class and fewer than k pass)")
            kf = min(4096, kf * 2)
        ctx.set_option("cand_cap", 0); ctx.set_option("list_entries", 0)
        scores, rows, t2i, counts = o_s.cpu(), o_r.cpu(), o_t.cpu(), o_c.cpu()
        with_t2i = True
    caption_map = _load_caption_map(args)
    mined_split = {"feature_list": [], "label_list": [], "file_list": []}
    num_imgs_sampled_dict = {}
    sampled_list: List[str] = []
    img = feat_bank
    img_host = img if not img.is_cuda else None
    for i, cls in enumerate(classes):
        if cls in empty:
            logger.info(f"class {cls} has no images. Continue")                       # :744
            continue
        n = int(counts[i])
        r = rows[i, :n]
        sc = scores[i, :n].tolist()
        ti = t2i[i, :n].tolist() if t2i is not None else None
        if k_run != num_samples:                             # host-side exclusion (unpartitioned): drop, keep the first num_samples
            bad = sets.get(str(cls), ())
            keep = [j for j, row in enumerate(r.tolist()) if paths[row] not in bad][:num_samples]
            r = r[keep]; sc = [sc[j] for j in keep]; ti = None if ti is None else [ti[j] for j in keep]
            n = len(keep)
        num_imgs_sampled_dict[cls] = n
        if n == 0:
            continue                                                                  # :472-474
        files = [paths[j] for j in r.tolist()]
        feats = (img_host[r] if img_host is not None else img[r.to(img.device)].cpu()).float()
        mined_split["feature_list"].append(feats)
        mined_split["label_list"].append(torch.full((n,), int(cls), dtype=torch.int64))
        mined_split["file_list"].append(files)
        for j, p in enumerate(files):
            caption = check_caption(caption_map, p) if caption_map is not None else ""
            if with_t2i:
                sampled_list.append(f"{round(sc[j], 4)}/{threshold}, {round(ti[j], 4)}/{t2i_threshold}, {p}, {caption}")
            else:
                sampled_list.append(f"{round(sc[j], 4)}/{threshold}, {p}, {caption}")
    prefixed = (not with_t2i) if file_prefix is None else file_prefix                 # :763,768,1236,1241 vs :817,822
    prefix = f"{args.prefix}_" if prefixed else ""
    os.makedirs(args.output_folder, exist_ok=True)
    # filtered_list (:761-764, :815-818): the rejected rows the walk met.  Partitioned data only (per class it lists up to
    # every row of the class); args.filtered_list = False skips this diagnostic and its extra pass over the bank.
    if row_class is not None and getattr(args, "filtered_list", True):
        t2t_all = _score_own_class(ctx, qs, rank_bank, row_class, device)
        pred_all = None
        if with_t2i:
            pred_all = _score_own_class(ctx, qs_pred if qs_pred is not None else qs, pred_bank, row_class, device)
        filtered_list = _filtered_lines(classes, row_class, paths, t2t_all, pred_all, excluded_rows, num_samples, threshold,
                                        t2i_threshold, caption_map)
        logger.info(f"len(filtered_list): {len(filtered_list)}")
        with open(f"{args.output_folder}/{prefix}filtered_list.txt", "w") as f:
            f.write("\n".join(filtered_list))
    if qs_pred is not None:
        qs_pred.close()
    qs.close()
    logger.info(f"len(sampled_list): {len(sampled_list)}")
    with open(f"{args.output_folder}/{prefix}sampled_list.txt", "w") as f:
        f.write("\n".join(sampled_list))
    return mined_split, num_imgs_sampled_dict


def t2t_ranked_sampler(args, logger, prompt_tensors, num_samples, threshold, pre_extracted_feats,
                       duplicates_dict: defaultdict = defaultdict(set), filtered_images_dict: defaultdict = defaultdict(set)):
    """``t2t_ranked_sampler`` (:724-771): per class the top ``num_samples`` rows by T2T cosine with
    ``similarity >= threshold``, ties by lowest row."""
    return _run_sampler(args, logger, prompt_tensors, num_samples, threshold, pre_extracted_feats, duplicates_dict,
                        filtered_images_dict, with_t2i=False)


def t2t_ranked_t2i_tshd_sampler(args, logger, prompt_tensors, num_samples, threshold, pre_extracted_feats,
                                duplicates_dict: defaultdict = defaultdict(set),
                                filtered_images_dict: defaultdict = defaultdict(set)):
    """``t2t_ranked_t2i_tshd_sampler`` (:774-825): walk rows in T2T-descending order, accept iff
    ``T2T >= threshold`` and ``T2I >= 0.25`` (:511-514), stop at ``num_samples``."""
    return _run_sampler(args, logger, prompt_tensors, num_samples, threshold, pre_extracted_feats, duplicates_dict,
                        filtered_images_dict, with_t2i=True, t2i_threshold=0.25)


def t2i_ranked_sampler(args, logger, prompt_tensors, num_samples, threshold, pre_extracted_feats,
                       duplicates_dict: defaultdict = defaultdict(set), filtered_images_dict: defaultdict = defaultdict(set)):
    """``t2i_ranked_sampler`` (:1195-1243): the same ranked walk on the IMAGE features
    (``cal_t2i_similarity`` :1224) -- the same kernel with the banks swapped."""
    return _run_sampler(args, logger, prompt_tensors, num_samples, threshold, pre_extracted_feats, duplicates_dict,
                        filtered_images_dict, with_t2i=False, rank_on_images=True)


def random_sampler(args, logger, prompt_tensors, num_samples, threshold, pre_extracted_feats,
                   duplicates_dict=None, filtered_images_dict=None, tail_head=False):
    """``random_sampler`` (:592-661): per class, shuffle the rows with Python's ``random`` (seeded by the caller,
    :1710) and accept the first ``num_samples`` that pass ``similarity >= threshold`` and the exclusion sets
    (``add_to_split`` :439-482).  ``similarity`` is 1.0, or the T2I score of the class prompt when ``threshold != 0``
    (:622-627; computed on the GPU).  No ranking is involved, so the walk itself stays on the host, as in the
    reference; the RNG is consumed exactly as ``random.shuffle(path_sim_zip)`` does."""
    import random
    dups = duplicates_dict if duplicates_dict is not None else defaultdict(set)
    filt = filtered_images_dict if filtered_images_dict is not None else defaultdict(set)
    caption_map = None
    cmap_path = getattr(args, "caption_map_path", None)
    if cmap_path is None:
        try:
            from .config import CAPTION_MAP_DICT
            cmap_path = CAPTION_MAP_DICT.get(args.dataset)
        except Exception:
            cmap_path = None
    if cmap_path and os.path.exists(cmap_path):
        with open(cmap_path, "rb") as f:                                                         # :599-601
            caption_map = pickle.load(f)
    classes = sorted(list(pre_extracted_feats.keys()), key=lambda x: int(x))                     # :603-604
    mined_split = {"feature_list": [], "label_list": [], "file_list": []}
    num_imgs_sampled_dict = {}
    filtered_list: List[str] = []
    sampled_list: List[str] = []
    tail_ct = 0
    for cls in classes:
        file_list = pre_extracted_feats[cls]["file_paths"]
        if file_list is None:
            num_imgs_sampled_dict[cls] = 0                                                       # :614-617
            continue
        img_embeddings = pre_extracted_feats[cls]["feats"]
        similarity = [1.0 for _ in range(len(file_list))]
        if threshold != 0:
            class_prompt = torch.as_tensor(prompt_tensors[cls]["mean"]).unsqueeze(0)             # :623-626
            similarity = cal_t2i_similarity(class_prompt, img_embeddings)
        order = list(range(len(file_list)))
        random.shuffle(order)                                                                    # :633 (same draws as shuffling the zip)
        if tail_head:
            if len(order) < num_samples:
                tail_ct += 1
            else:
                threshold = 0                                                                    # :636-641 -- sticks for later classes
        ct, acc = 0, []
        for i in order:                                                                          # add_to_split :450-469
            if ct == num_samples:
                break
            fp, sim = file_list[i], similarity[i]
            caption = check_caption(caption_map, fp) if caption_map is not None else ""
            info = f"{round(sim, 4)}/{threshold}, {fp}, {caption}"
            if sim >= threshold and fp not in dups[str(cls)] and fp not in filt[str(cls)]:
                acc.append(i)
                ct += 1
                sampled_list.append(info)
            else:
                filtered_list.append(info)
        if acc:
            idx = torch.tensor(acc, dtype=torch.int64)
            mined_split["feature_list"].append(torch.as_tensor(img_embeddings)[idx].float())
            mined_split["label_list"].append(torch.full((len(acc),), int(cls), dtype=torch.int64))
            mined_split["file_list"].append([file_list[i] for i in acc])
        num_imgs_sampled_dict[cls] = ct
    if tail_head:
        print(f"Number of tail classes: {tail_ct}")
    os.makedirs(args.output_folder, exist_ok=True)
    logger.info(f"len(filtered_list): {len(filtered_list)}")
    with open(f"{args.output_folder}/{args.prefix}_filtered_list.txt", "w") as f:                # :649-651
        f.write("\n".join(filtered_list))
    logger.info(f"len(sampled_list): {len(sampled_list)}")
    with open(f"{args.output_folder}/{args.prefix}_sampled_list.txt", "w") as f:                 # :654-656
        f.write("\n".join(sampled_list))
    return mined_split, num_imgs_sampled_dict


def _load_fewshot(args):
    fs = getattr(args, "fewshot_features", None)
    if fs is not None:
        return fs
    fn = getattr(args, "fewshot_path", None) or \
        f"../data/{args.dataset}/pre_extracted/{args.dataset}_probing_vitb32_openclip_laion400m_1_train_features.pth"
    fea = torch.load(fn, map_location="cpu", weights_only=False)                      # get_fewshot_features :997-1014
    out: Dict[int, list] = {}
    for f, l in zip(fea["image_features"], fea["labels"].tolist()):
        out.setdefault(int(l), []).append(f)
    return out


def i2i_ranked_sampler_p2p(args, logger, prompt_tensors, num_samples, threshold, pre_extracted_feats,
                           duplicates_dict: defaultdict = defaultdict(set), filtered_images_dict: defaultdict = defaultdict(set)):
    """``i2i_ranked_sampler_p2p`` (:1016-1076): rank by the MEAN cosine between a row's image feature and
    the class's 16 few-shot image features (``i2i_similarity_p2p(..., 'mean')`` :1049)."""
    return _run_sampler(args, logger, prompt_tensors, num_samples, threshold, pre_extracted_feats, duplicates_dict,
                        filtered_images_dict, with_t2i=False, rank_on_images=True, rank_fewshot=_load_fewshot(args))


def i2t_rank_sampler(args, logger, prompt_tensors, num_samples, threshold, pre_extracted_feats,
                     duplicates_dict: defaultdict = defaultdict(set), filtered_images_dict: defaultdict = defaultdict(set)):
    """``i2t_rank_sampler`` (:1079-1133): the same ranking against the CAPTION features (:1112)."""
    return _run_sampler(args, logger, prompt_tensors, num_samples, threshold, pre_extracted_feats, duplicates_dict,
                        filtered_images_dict, with_t2i=False, rank_on_images=False, rank_fewshot=_load_fewshot(args))


def t2t_rank_i2t_tshd_sampler(args, logger, prompt_tensors, num_samples, threshold, pre_extracted_feats,
                              duplicates_dict: defaultdict = defaultdict(set), filtered_images_dict: defaultdict = defaultdict(set)):
    """``t2t_rank_i2t_tshd_sampler`` (:831-890): T2T ranking; accept iff the MAX cosine between the row's
    caption feature and the class's few-shot image features is >= 0.25 (:869, default threshold :499)."""
    return _run_sampler(args, logger, prompt_tensors, num_samples, threshold, pre_extracted_feats, duplicates_dict,
                        filtered_images_dict, with_t2i=True, t2i_threshold=0.25, pred_fewshot=_load_fewshot(args),
                        pred_on_captions=True, file_prefix=False)


def t2t_rank_i2i_tshd_sampler(args, logger, prompt_tensors, num_samples, threshold, pre_extracted_feats,
                              duplicates_dict: defaultdict = defaultdict(set), filtered_images_dict: defaultdict = defaultdict(set)):
    """``t2t_rank_i2i_tshd_sampler`` (:893-953): T2T ranking; accept iff the MAX cosine between the row's
    image feature and the class's few-shot image features is >= 0.65 (:929, :939)."""
    return _run_sampler(args, logger, prompt_tensors, num_samples, threshold, pre_extracted_feats, duplicates_dict,
                        filtered_images_dict, with_t2i=True, t2i_threshold=0.65, pred_fewshot=_load_fewshot(args),
                        pred_on_captions=False, file_prefix=False)


# ---------------------------------------------------------------------------------------------
# output writer + orchestration tail
# ---------------------------------------------------------------------------------------------
def save_sample_file_list(args, final_file_list, label_tensor, logger=None, copy_to: Optional[str] = None):
    """``save_sample_file_list`` (:1457-1469): ``"<path> <label> 0\\n"`` per accepted row."""
    fn = f"{args.output_folder}/{args.prefix}.txt"
    labels = torch.as_tensor(label_tensor).tolist()
    with open(fn, "w") as f:
        f.write("".join(f"{p} {l} {0}\n" for p, l in zip(final_file_list, labels)))
    if logger:
        logger.info(f"Saved file_list to: {fn}")
    if copy_to:
        os.makedirs(copy_to, exist_ok=True)
        shutil.copy(fn, copy_to)
        if logger:
            logger.info(f"Copied file to: {copy_to}")
    return fn


# Module globals of the reference script that ``sampling`` reads (set under ``__main__`` there, :1736-1740): the CLI
# (swat_b200/sample_retrieval.py) or the embedding application assigns them before calling ``sampling``.
prompt_tensors_dict: Dict[str, dict] = {}


def _zeroshot_head_weight(prompt_tensors) -> torch.Tensor:
    """``features.prompt_sampler(prompt_tensors, sample_by='mean')`` (utils/features.py:12-24): the class prompts stacked
    in dict order -> weights of the zero-shot head ``MyLinear(weights=..., bias=False)`` (:1489-1490)."""
    return torch.stack([torch.as_tensor(prompt_tensors[k]["mean"]).detach().float().cpu().reshape(-1) for k in prompt_tensors.keys()], dim=0)


def sampling(args, logger, model=None, preprocess=None, metrics=None, dataset_root=None, *, prompt_tensors=None,
             pre_extracted_feats=None, copy_to: Optional[str] = "default"):
    """``sampling(args, logger, model, preprocess, metrics, dataset_root)`` (:1471-1670), same positional signature and
    return value ``(file_list_path, sample_ct)``: load + regroup the mined features, optional zero-shot filtering
    (:1484-1497) and near-duplicate removal (:1499-1507) feeding the exclusion sets of every sampler, dispatch on
    ``args.sampling_method`` (:1517-1617), write ``{prefix}.txt`` and ``{prefix}_num_imgs_sampled.json``.

    ``model`` / ``preprocess`` / ``metrics`` are accepted for signature compatibility and unused on this path (the
    reference passes them on to code that requires pre-extracted features anyway, :293-297).  The prompt tensors come
    from the module global ``prompt_tensors_dict[args.prompt_name]`` like in the reference (:1489, :1519), or from the
    keyword ``prompt_tensors``.  Keyword extensions: ``pre_extracted_feats`` (skip the load), ``copy_to`` (where the
    split txt is copied; default ``../data/{dataset}/`` :1466, ``None`` = no copy)."""
    if prompt_tensors is None:
        if getattr(args, "prompt_name", None) not in prompt_tensors_dict:
            raise KeyError(f"prompt tensors for prompt_name={getattr(args, 'prompt_name', None)!r} not set: assign "
                           "swat_b200.retrieval.prompt_tensors_dict (the reference's module global) or pass prompt_tensors=")
        prompt_tensors = prompt_tensors_dict[args.prompt_name]
    if pre_extracted_feats is None:
        fn = f"{dataset_root}/{args.dataset}_{args.model_cfg}_mined.pth"
        if not os.path.exists(fn):
            logger.info(f"Error: Pre-extracted features not found. {fn}")
            raise NotImplementedError                                               # :1478-1479
        from .shards import load_mined_pth
        pre_extracted_feats = load_mined_pth(fn)
        logger.info(f"Loaded pre-extracted mined features from: {fn}")
    feats = transform_extracted_fea(pre_extracted_feats) if "labels" in pre_extracted_feats else pre_extracted_feats
    # ---------- zero-shot CLIP image filtering (:1484-1497)
    if not getattr(args, "zeroshot_img_filter", False):
        logger.info("No zeroshot image filtering!")
        filtered_images_dict = defaultdict(set)
    else:
        logger.info("Doing zeroshot image filtering!")
        filtered_images_dict = zeroshot_clip_img_filter(model=model, preprocess=preprocess, root_folder=getattr(args, "root_folder", None),
                                                        pre_extracted_feats=feats, head=_zeroshot_head_weight(prompt_tensors))
    # ---------- image de-duplication (:1499-1507)
    if not getattr(args, "image_dedup", False):
        logger.info("No image deduplication!")
        duplicate_images_dict = defaultdict(set)
    else:
        logger.info("Doing image deduplication!")
        duplicate_images_dict, dup_images_fraction, avg_dup_images_fraction = remove_near_duplicates2(feats)
        logger.info(f"dup_images_fraction: {dup_images_fraction}")
        logger.info(f"avg_dup_images_fraction: {avg_dup_images_fraction}")
    logger.info(f"Sampling method: {args.sampling_method}, sampling number: {args.num_samples}, "
                f"sampling threshold: {args.sampling_threshold}")
    kw = dict(duplicates_dict=duplicate_images_dict, filtered_images_dict=filtered_images_dict)
    m = args.sampling_method
    if m == "Random":                                                                # :1517-1526
        mined_split, num_imgs_sampled_dict = random_sampler(args, logger, prompt_tensors, args.num_samples, 0.0, feats, tail_head=False, **kw)
    elif m == "T2T-rank":                                                            # :1571-1579
        mined_split, num_imgs_sampled_dict = t2t_ranked_sampler(args, logger, prompt_tensors, args.num_samples, 0.0, feats, **kw)
    elif m == "T2T-rank-T2I-tshd":                                                   # :1581-1589
        mined_split, num_imgs_sampled_dict = t2t_ranked_t2i_tshd_sampler(args, logger, prompt_tensors, args.num_samples, 0.0, feats, **kw)
    elif m == "T2I-rank":                                                            # :1610-1617
        mined_split, num_imgs_sampled_dict = t2i_ranked_sampler(args, logger, prompt_tensors, args.num_samples, 0.0, feats, **kw)
    elif m == "I2I-rank":                                                            # :1538-1551
        mined_split, num_imgs_sampled_dict = i2i_ranked_sampler_p2p(args, logger, prompt_tensors, args.num_samples, 0.0, feats, **kw)
    elif m == "I2T-rank":                                                            # :1553-1560
        mined_split, num_imgs_sampled_dict = i2t_rank_sampler(args, logger, prompt_tensors, args.num_samples, 0.0, feats, **kw)
    elif m == "T2T-rank-I2T-tshd":                                                   # :1591-1598
        mined_split, num_imgs_sampled_dict = t2t_rank_i2t_tshd_sampler(args, logger, prompt_tensors, args.num_samples, 0.0, feats, **kw)
    elif m == "T2T-rank-I2I-tshd":                                                   # :1600-1607
        mined_split, num_imgs_sampled_dict = t2t_rank_i2i_tshd_sampler(args, logger, prompt_tensors, args.num_samples, 0.0, feats, **kw)
    else:
        raise NotImplementedError(f"sampling method {m} is outside the accelerated hot path")
    logger.info(f"len(file_list): {len(mined_split['file_list'])}")
    final_file_list = [p for fl in mined_split["file_list"] for p in fl]
    logger.info(f"len(final_file_list): {len(final_file_list)}")
    labels_tensor = torch.cat(mined_split["label_list"], dim=0) if mined_split["label_list"] else torch.zeros(0, dtype=torch.int64)
    if mined_split["feature_list"]:
        feature_tensor = torch.cat(mined_split["feature_list"], dim=0)
        logger.info(f"feature_tensor.shape: {feature_tensor.shape}")
    logger.info(f"labels_tensor.shape: {labels_tensor.shape}")
    if copy_to == "default":
        copy_to = f"../data/{args.dataset}/"                                         # :1466
    file_list_path = save_sample_file_list(args, final_file_list, labels_tensor, logger, copy_to)
    with open(f"{args.output_folder}/{args.prefix}_num_imgs_sampled.json", "w") as f:
        json.dump(num_imgs_sampled_dict, f, indent=4)                               # :1666-1668
    return file_list_path, len(final_file_list)
