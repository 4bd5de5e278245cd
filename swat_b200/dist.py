"""Row-sharded multi-GPU retrieval: one process per GPU, each rank scans its own
contiguous range of bank rows, then ONE gather of the per-class candidate lists and a merge.

The reference has no distributed code at all (single process, single GPU,
``sample_retrieval.py:1737``); the sharding follows from the algorithm: rows are independent and the
per-class top-k under (score desc, row asc) is associative.  For the T2I walk
(``add_t2t_ranked_t2i_tshd_to_split`` :492-540) the ranks exchange *candidates* (T2T top-k_fetch with
their T2I score), not locally walked results: the accept walk runs once, globally, in the merge, and
the merge proves exactness against each truncated shard's frontier.
"""
from __future__ import annotations

import math
import os
from typing import Callable, Optional, Tuple

import torch

from . import _lib


def shard_range(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced split: rank r owns rows [start, end)."""
    base, rem = divmod(int(n_rows), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def local_candidates(ctx, queries, t2t_bank: torch.Tensor, k_fetch: int, t2t_threshold: float = 0.0,
                     t2i_bank: Optional[torch.Tensor] = None, row_offset: int = 0,
                     row_class: Optional[torch.Tensor] = None, exclude: Optional[torch.Tensor] = None,
                     class_depth: Optional[torch.Tensor] = None):
    """This shard's T2T top-``k_fetch`` per class (global row ids) with the T2I score of every
    candidate.  Returns ``(scores, rows, t2i | None, counts, truncated)`` on the device."""
    dbg = os.environ.get("SWAT_DEBUG")
    cap, lists = None, None
    cache = ctx.__dict__.setdefault("_job_cache", {})
    for _ in range(8):
        key = (id(queries), int(k_fetch), float(t2t_threshold), cap, lists)
        job = cache.get(key)
        if job is None:                       # job buffers (survivor lists: ~100s of MB) are reused across calls
            if len(cache) >= 4:
                cache.pop(next(iter(cache))).close()
            job = cache[key] = _lib.Job(ctx, queries, k_fetch, t2t_threshold)
        else:
            job.reset()
        job.set_class_depth(class_depth)
        job.scan(t2t_bank, row_base=0, row_class=row_class, exclude=exclude)
        scores, rows, counts, trunc = job.select()
        over = job.overflowed()
        if dbg:
            print(f"[swat dist] scan+select k_fetch={k_fetch} overflow={over}", flush=True)
        if not over:
            break
        if over & 1:
            cap = (cap or (2 * k_fetch + 4096)) * 4
            ctx.set_option("cand_cap", cap)
        if over & 2:
            lists = (lists or (4 << 20)) * 4
            ctx.set_option("list_entries", lists)
    else:
        raise _lib.SwatError(-4, "candidate buffers keep overflowing")
    if cap is not None:
        ctx.set_option("cand_cap", 0)
    if lists is not None:
        ctx.set_option("list_entries", 0)
    t2i = None
    if t2i_bank is not None:
        # threshold -inf and k == k_fetch: every candidate is kept in order, we only want its T2I score
        scores, rows, t2i, counts, _ = _lib.t2i_walk(ctx, queries, t2i_bank, scores, rows, counts, None, k_fetch,
                                                     float("-inf"), img_row_base=0)
    rows = torch.where(rows >= 0, rows + int(row_offset), rows)
    if dbg:
        torch.cuda.synchronize()
        print("[swat dist] local candidates ready", flush=True)
    return scores, rows, t2i, counts, trunc


def pack(scores, rows, t2i, counts, trunc) -> torch.Tensor:
    """One int32 buffer per rank so the exchange is a single collective."""
    parts = [rows.contiguous().view(torch.int32).flatten(), scores.contiguous().view(torch.int32).flatten()]
    if t2i is not None:
        parts.append(t2i.contiguous().view(torch.int32).flatten())
    parts += [counts.to(torch.int32).flatten(), trunc.to(torch.int32).flatten()]
    return torch.cat(parts)


def unpack(buf: torch.Tensor, world: int, n_classes: int, k_fetch: int, with_t2i: bool):
    buf = buf.view(world, -1)
    n = n_classes * k_fetch
    o = 0
    rows = buf[:, o:o + 2 * n].contiguous().view(torch.int64).view(world, n_classes, k_fetch); o += 2 * n
    scores = buf[:, o:o + n].contiguous().view(torch.float32).view(world, n_classes, k_fetch); o += n
    t2i = None
    if with_t2i:
        t2i = buf[:, o:o + n].contiguous().view(torch.float32).view(world, n_classes, k_fetch); o += n
    counts = buf[:, o:o + n_classes].contiguous(); o += n_classes
    trunc = buf[:, o:o + n_classes].contiguous()
    return scores, rows, t2i, counts, trunc


def gather_merge(local, k: int, t2i_threshold: float, world: int, ctx=None, group=None,
                 merge_fn: Optional[Callable] = None):
    """Single all-gather of the packed candidate lists (NCCL over NVLink on GPUs, gloo in the CPU
    tests), then the merge walk.  ``merge_fn(scores, rows, t2i, counts, trunc, k, thr)`` replaces the
    CUDA merge in the CPU tests.  Returns ``(scores, rows, t2i | None, counts, incomplete)``."""
    import torch.distributed as dist
    scores, rows, t2i, counts, trunc = local
    n_classes, k_fetch = scores.shape
    mine = pack(scores, rows, t2i, counts, trunc)
    if world > 1:
        out = torch.empty(world * mine.numel(), dtype=torch.int32, device=mine.device)
        dist.all_gather_into_tensor(out, mine, group=group)
        if os.environ.get("SWAT_DEBUG"):
            torch.cuda.synchronize()
            print("[swat dist] all_gather done", flush=True)
    else:
        out = mine
    g_scores, g_rows, g_t2i, g_counts, g_trunc = unpack(out, world, n_classes, k_fetch, t2i is not None)
    thr = t2i_threshold if t2i is not None else float("-inf")
    if merge_fn is not None:
        return merge_fn(g_scores, g_rows, g_t2i, g_counts, g_trunc, k, thr)
    return _lib.merge_topk(ctx, g_scores, g_rows, g_counts, aux=g_t2i, truncated=g_trunc, k_out=k, aux_threshold=thr)


def topk_sharded(ctx, queries, t2t_bank: torch.Tensor, k: int, t2t_threshold: float = 0.0,
                 t2i_bank: Optional[torch.Tensor] = None, t2i_threshold: float = 0.25, row_offset: int = 0,
                 world: int = 1, group=None, k_fetch: Optional[int] = None, max_k_fetch: int = 4096):
    """Whole multi-GPU pipeline for this rank's shard.  Every rank returns the merged result.
    Classes whose walk is not provably exact are escalated collectively (4x deeper over-fetch for
    those classes only, finally the exact in-pass predicate)."""
    base = k if t2i_bank is None else max(1024, 2 * k)
    if k_fetch is None:
        # T2T only: the merged top-k never reaches below a shard's k-th candidate, k suffices.
        # T2I walk: over-fetch so that k candidates pass the predicate above every shard's frontier;
        # classes whose walk needed more depth before start deeper (per class, remembered on `queries`).
        k_fetch, depth = base, None
        hint = queries.__dict__.get("_depth_hint")
        if t2i_bank is not None and hint is not None and int(hint.max()) > base:
            depth = torch.clamp(hint, min=base, max=max_k_fetch).to(torch.int32)
            k_fetch = int(depth.max())
    else:
        depth = None
    k_fetch = max(1, min(int(k_fetch), max_k_fetch))
    local = local_candidates(ctx, queries, t2t_bank, k_fetch, t2t_threshold, t2i_bank, row_offset, class_depth=depth)
    res = gather_merge(local, k, t2i_threshold, world, ctx=ctx, group=group)
    bad = res[4].nonzero().flatten().tolist()          # identical on every rank: the merge input is the all-gather
    if not bad:
        return res
    out_s, out_r, out_t, out_c, _ = res
    sub = queries.subset(bad)
    was = k_fetch if depth is None else int(depth[bad].min())
    if was < max_k_fetch:
        # targeted escalation: only the classes that are not provably exact are re-scanned, 2x deeper
        nxt = min(max_k_fetch, 2 * was)
        r2 = topk_sharded(ctx, sub, t2t_bank, k, t2t_threshold, t2i_bank, t2i_threshold, row_offset, world, group,
                          k_fetch=nxt, max_k_fetch=max_k_fetch)
        reached = max(nxt, int(sub.__dict__.get("_reached", nxt)))
        queries.__dict__["_reached"] = reached
        if queries.__dict__.get("_depth_hint") is None:
            queries.__dict__["_depth_hint"] = torch.zeros(queries.n_classes, dtype=torch.int32)
        queries.__dict__["_depth_hint"][bad] = torch.maximum(queries.__dict__["_depth_hint"][bad], torch.tensor(reached, dtype=torch.int32))
    else:
        # The walk reaches below the deepest over-fetch of some shard (few rows pass T2I): every shard
        # computes its exact local top-k of predicate-passing rows (swat_topk falls back to the in-pass
        # predicate where needed); top-k of passing rows is associative, so a plain merge finishes it.
        s, r, t, c = _lib.topk(ctx, sub, t2t_bank, k, t2t_threshold, t2i_bank=t2i_bank, t2i_threshold=t2i_threshold,
                               row_offset=row_offset)
        r2 = gather_merge((s, r, t, c, torch.zeros_like(c)), k, float("-inf"), world, ctx=ctx, group=group)
    idx = torch.tensor(bad, device=out_s.device)
    out_s[idx], out_r[idx], out_c[idx] = r2[0], r2[1], r2[3]
    if out_t is not None:
        out_t[idx] = r2[2]
    return out_s, out_r, out_t, out_c, torch.zeros_like(out_c)
