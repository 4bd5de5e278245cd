"""Row-sharded multi-GPU retrieval: one process per GPU, each rank scans its own
contiguous range of bank rows, then ONE gather of the per-class walk results and a merge.

The reference has no distributed code at all (single process, single GPU,
``sample_retrieval.py:1737``); the sharding follows from the algorithm: rows are independent and the
per-class top-k of predicate-passing rows under (score desc, row asc) is associative.  Every rank
walks its own candidates (exact re-score, accept predicate; ``add_t2t_ranked_t2i_tshd_to_split``
:492-540) and ships at most k accepted rows per class plus one float, its *limit*: the score down to
which its list is proven complete (``-inf`` when it found k rows or saw every eligible row).  The merge
keeps the k best of the union and is exact iff all of them lie above every shard's limit -- a shard
with few passing rows never has to scan to the bottom for the global answer.
"""
from __future__ import annotations

import os
from typing import Callable, Optional, Tuple

import torch

from . import _lib

MAX_K_FETCH = 4096


def shard_range(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced split: rank r owns rows [start, end)."""
    base, rem = divmod(int(n_rows), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def packed_layout(n_classes: int, k: int, with_t2i: bool) -> dict:
    """int32 offsets of one rank's exchange buffer: ``rows`` (int64, first so it stays 8-byte aligned), ``scores``,
    ``t2i`` (optional), ``counts``, ``limit`` (float32 bits), one ``flags`` word (the job's overflow bits) and padding
    to an even length, so that rank r's slice of the all-gathered buffer starts on an 8-byte boundary too."""
    n = int(n_classes) * int(k)
    lay = {"rows": 0, "scores": 2 * n}
    o = 3 * n
    lay["t2i"] = o if with_t2i else None
    o += n if with_t2i else 0
    lay["counts"] = o; o += n_classes
    lay["limit"] = o; o += n_classes
    lay["flags"] = o; o += 1
    lay["len"] = o + (o & 1)
    lay["n_classes"], lay["k"] = int(n_classes), int(k)
    return lay


class PackedResults:
    """One rank's exchange buffer with typed views into it: the walk kernel writes its outputs straight into the
    views, the buffer goes into the all-gather as it is, and the merge reads the gathered buffer through a
    per-shard stride -- no pack or unpack copies."""

    def __init__(self, n_classes: int, k: int, with_t2i: bool, device, buf: Optional[torch.Tensor] = None):
        self.lay = lay = packed_layout(n_classes, k, with_t2i)
        n = n_classes * k
        self.buf = torch.zeros(lay["len"], dtype=torch.int32, device=device) if buf is None else buf
        b = self.buf
        self.rows = b[0:2 * n].view(torch.int64).view(n_classes, k)
        self.scores = b[lay["scores"]:lay["scores"] + n].view(torch.float32).view(n_classes, k)
        self.t2i = b[lay["t2i"]:lay["t2i"] + n].view(torch.float32).view(n_classes, k) if with_t2i else None
        self.counts = b[lay["counts"]:lay["counts"] + n_classes]
        self.limit = b[lay["limit"]:lay["limit"] + n_classes].view(torch.float32)
        self.flags = b[lay["flags"]:lay["flags"] + 1]


def default_k_fetch(k: int, with_t2i: bool, eps: float) -> int:
    """First over-fetch of a walk (mirrors the library's choice for resident banks): without a predicate the
    candidates must reach 2 eps below the k-th score; with one, deep enough for k rows to pass."""
    wide = eps > 1e-3
    if not with_t2i:
        kf = k + (max(1024, k) if wide else max(64, k // 8))
    else:
        kf = max(2 * k, 1024) + (1024 if wide else 0)
    return max(1, min((kf + 31) // 32 * 32, MAX_K_FETCH))


def _get_job(ctx, queries, k_fetch, threshold, cap, lists):
    cache = queries.__dict__.setdefault("_job_cache", {})        # closed with the query set (Queries.close)
    key = (int(k_fetch), float(threshold), cap, lists)
    job = cache.get(key)
    if job is None:                       # job buffers (survivor lists: ~100s of MB) are reused across calls
        if len(cache) >= 3:
            cache.pop(next(iter(cache))).close()
        job = cache[key] = _lib.Job(ctx, queries, k_fetch, threshold)
    else:
        job.reset()
    return job


def local_walk(ctx, queries, t2t_bank: torch.Tensor, k: int, k_fetch: int, t2t_threshold: float = 0.0,
               t2i_bank: Optional[torch.Tensor] = None, t2i_threshold: float = 0.25, row_offset: int = 0,
               row_class: Optional[torch.Tensor] = None, exclude: Optional[torch.Tensor] = None,
               class_depth: Optional[torch.Tensor] = None, packed: Optional[PackedResults] = None, check: bool = True):
    """This shard's walk: scan -> T2T top-``k_fetch`` per class -> exact re-score (+ T2I predicate) -> the first
    ``k`` accepted rows (global ids) and the shard's limit.  Returns ``(scores, rows, t2i | None, counts, limit)``
    on the device.

    ``packed``: write the results into that exchange buffer.  ``check=False`` (with ``packed``): do not
    synchronise to test the job's overflow bits; they are copied into ``packed.flags`` on the stream and
    travel with the results, so every rank sees every rank's bits after the exchange."""
    dbg = os.environ.get("SWAT_DEBUG")
    cap, lists = None, None
    eps = _lib.scan_eps(queries, t2t_bank.dtype)
    for _ in range(8):
        job = _get_job(ctx, queries, k_fetch, t2t_threshold - eps, cap, lists)
        job.set_class_depth(class_depth)
        timed = ctx.__dict__.get("_time_scans")       # bench.py: CUDA events around the scan, on the launching stream
        if timed is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        job.scan(t2t_bank, row_base=0, row_class=row_class, exclude=exclude)
        if timed is not None:
            e1.record()
            timed.append((e0, e1))
        scores, rows, counts, trunc = job.select(row_offset)
        if packed is not None and not check:
            job.export_flags(packed.flags)
            break
        over = job.overflowed()
        if dbg:
            print(f"[swat dist] scan+select k_fetch={k_fetch} overflow={over}", flush=True)
        if not over:
            if packed is not None:
                packed.flags.zero_()
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
    out = None
    if packed is not None:
        t2i_view = packed.t2i if packed.t2i is not None else torch.empty(queries.n_classes, k, dtype=torch.float32, device=t2t_bank.device)
        out = (packed.scores, packed.rows, t2i_view, packed.counts, packed.limit)
    o_s, o_r, o_t, o_c, o_l, _ = _lib.rescore_walk(ctx, queries, t2t_bank, scores, rows, counts, trunc, k, t2t_threshold,
                                                  aux_bank=t2i_bank, aux_threshold=t2i_threshold, bank_row_base=row_offset,
                                                  eps=eps, out=out)
    if dbg:
        torch.cuda.synchronize()
        print("[swat dist] local walk ready", flush=True)
    return o_s, o_r, (o_t if t2i_bank is not None else None), o_c, o_l


def pack(scores, rows, t2i, counts, limit, flags: int = 0) -> torch.Tensor:
    """One int32 buffer per rank (layout: ``packed_layout``) so the exchange is a single collective."""
    Cn, k = scores.shape
    p = PackedResults(Cn, k, t2i is not None, scores.device)
    p.rows.copy_(rows); p.scores.copy_(scores)
    if t2i is not None:
        p.t2i.copy_(t2i)
    p.counts.copy_(counts.to(torch.int32)); p.limit.copy_(limit.to(torch.float32))
    p.flags.fill_(int(flags))
    return p.buf


def unpack(buf: torch.Tensor, world: int, n_classes: int, k: int, with_t2i: bool):
    """Dense ``[world, ...]`` copies of the gathered arrays (CPU tests, generic merge functions)."""
    buf = buf.view(world, -1)
    parts = [PackedResults(n_classes, k, with_t2i, buf.device, buf=buf[w].contiguous()) for w in range(world)]
    st = lambda name: torch.stack([getattr(p, name) for p in parts])
    return st("scores"), st("rows"), (st("t2i") if with_t2i else None), st("counts"), st("limit")


def unpack_flags(buf: torch.Tensor, world: int, n_classes: int, k: int, with_t2i: bool) -> torch.Tensor:
    lay = packed_layout(n_classes, k, with_t2i)
    return buf.view(world, -1)[:, lay["flags"]]


def gather_packed(mine: torch.Tensor, world: int, group=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Single all-gather of the packed walk results (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
    if world <= 1:
        return mine
    import torch.distributed as dist
    if out is None:
        out = torch.empty(world * mine.numel(), dtype=torch.int32, device=mine.device)
    dist.all_gather_into_tensor(out, mine, group=group)
    if os.environ.get("SWAT_DEBUG"):
        torch.cuda.synchronize()
        print("[swat dist] all_gather done", flush=True)
    return out


def merge_packed(gathered: torch.Tensor, lay: dict, world: int, k: int, ctx=None, merge_fn: Optional[Callable] = None):
    """Merge over an all-gathered packed buffer.  On the GPU the merge kernel reads the buffer in place
    (per-shard stride = one rank's buffer); ``merge_fn`` (CPU tests) gets dense copies."""
    Cn, kin, with_t2i = lay["n_classes"], lay["k"], lay["t2i"] is not None
    if merge_fn is not None:
        return merge_fn(*unpack(gathered, world, Cn, kin, with_t2i), k)
    first = PackedResults(Cn, kin, with_t2i, gathered.device, buf=gathered[:lay["len"]])
    return _lib.merge_topk(ctx, first.scores, first.rows, first.counts, aux=first.t2i, limit=first.limit, k_out=k,
                           n_shards=world, shard_stride_bytes=lay["len"] * 4)


def gather_merge(local, k: int, world: int, ctx=None, group=None, merge_fn: Optional[Callable] = None):
    """Exchange + merge for walk results ``(scores, rows, t2i | None, counts, limit)``.
    ``merge_fn(scores, rows, t2i, counts, limit, k)`` replaces the CUDA merge in the CPU tests.
    Returns ``(scores, rows, t2i | None, counts, incomplete)``."""
    scores, rows, t2i, counts, limit = local
    n_classes, k_in = scores.shape
    out = gather_packed(pack(scores, rows, t2i, counts, limit), world, group)
    return merge_packed(out, packed_layout(n_classes, k_in, t2i is not None), world, k, ctx, merge_fn)


def _sub_row_class(row_class: Optional[torch.Tensor], classes, n_classes: int) -> Optional[torch.Tensor]:
    """row_class renumbered to a sub-query set: class ``classes[i]`` becomes ``i``, every other row -1."""
    if row_class is None:
        return None
    remap = torch.full((n_classes + 1,), -1, dtype=torch.int32, device=row_class.device)
    remap[torch.tensor(list(classes), device=row_class.device)] = torch.arange(len(classes), dtype=torch.int32, device=row_class.device)
    idx = torch.where(row_class >= 0, row_class, torch.full_like(row_class, n_classes)).long()
    return remap[idx].contiguous()


def topk_sharded(ctx, queries, t2t_bank: torch.Tensor, k: int, t2t_threshold: float = 0.0,
                 t2i_bank: Optional[torch.Tensor] = None, t2i_threshold: float = 0.25, row_offset: int = 0,
                 world: int = 1, group=None, k_fetch: Optional[int] = None, max_k_fetch: int = MAX_K_FETCH,
                 row_class: Optional[torch.Tensor] = None, exclude: Optional[torch.Tensor] = None):
    """Whole multi-GPU pipeline for this rank's shard (``row_class`` / ``exclude`` cover this rank's rows).  Every
    rank returns the merged result ``(scores, rows, t2i | None, counts, incomplete)``.  Classes the merge cannot prove
    exact are escalated collectively (2x deeper over-fetch for those classes only, finally every shard's own exact
    top-k through ``swat_topk``)."""
    eps = _lib.scan_eps(queries, t2t_bank.dtype)
    base = default_k_fetch(k, t2i_bank is not None, eps)
    if k_fetch is None:
        # classes whose walk needed more depth before start deeper (per class, remembered on `queries`)
        k_fetch, depth = base, None
        hint = queries.__dict__.get("_depth_hint")
        if hint is not None and int(hint.max()) > base:
            depth = torch.clamp(hint, min=base, max=max_k_fetch).to(torch.int32)
            k_fetch = int(depth.max())
    else:
        depth = None
    k_fetch = max(k, min(int(k_fetch), max_k_fetch))
    # Optimistic pass: nothing synchronises before the exchange.  Kernels write into the packed buffer, the overflow
    # bits travel with it, and ONE read-back at the end fetches every rank's bits and the merge's `incomplete` flags.
    Cn, with_t2i = queries.n_classes, t2i_bank is not None
    bufs = ctx.__dict__.setdefault("_packed_cache", {})
    key = (Cn, k, with_t2i, world)
    if key not in bufs:
        if len(bufs) >= 4:
            bufs.pop(next(iter(bufs)))
        p = PackedResults(Cn, k, with_t2i, t2t_bank.device)
        bufs[key] = (p, torch.empty(world * p.lay["len"], dtype=torch.int32, device=t2t_bank.device) if world > 1 else None)
    packed, gbuf = bufs[key]
    local_walk(ctx, queries, t2t_bank, k, k_fetch, t2t_threshold, t2i_bank, t2i_threshold, row_offset, row_class, exclude,
               class_depth=depth, packed=packed, check=False)
    gathered = gather_packed(packed.buf, world, group, out=gbuf)
    res = merge_packed(gathered, packed.lay, world, k, ctx=ctx)
    status = torch.cat([unpack_flags(gathered, world, Cn, k, with_t2i), res[4]]).tolist()   # the step's only host sync
    if any(status[:world]):
        # some rank's candidate buffers overflowed (identical view on every rank): redo with the checked local stage,
        # which grows that rank's buffers and rescans before anything is exchanged
        local = local_walk(ctx, queries, t2t_bank, k, k_fetch, t2t_threshold, t2i_bank, t2i_threshold, row_offset, row_class,
                           exclude, class_depth=depth)
        res = gather_merge(local, k, world, ctx=ctx, group=group)
        status = [0] * world + res[4].tolist()
    bad = [c for c, v in enumerate(status[world:]) if v]          # identical on every rank: the merge input is the all-gather
    if not bad:
        return res
    out_s, out_r, out_t, out_c, _ = res
    sub = queries.subset(bad)
    sub_rc = _sub_row_class(row_class, bad, Cn)
    was = k_fetch if depth is None else int(depth[bad].min())
    if was < max_k_fetch:
        # targeted escalation: only the classes that are not provably exact are re-scanned, 2x deeper
        nxt = min(max_k_fetch, 2 * was)
        r2 = topk_sharded(ctx, sub, t2t_bank, k, t2t_threshold, t2i_bank, t2i_threshold, row_offset, world, group,
                          k_fetch=nxt, max_k_fetch=max_k_fetch, row_class=sub_rc, exclude=exclude)
        reached = max(nxt, int(sub.__dict__.get("_reached", nxt)))
        queries.__dict__["_reached"] = reached
        if queries.__dict__.get("_depth_hint") is None:
            queries.__dict__["_depth_hint"] = torch.zeros(queries.n_classes, dtype=torch.int32)
        queries.__dict__["_depth_hint"][bad] = torch.maximum(queries.__dict__["_depth_hint"][bad], torch.tensor(reached, dtype=torch.int32))
    else:
        # The walk reaches below the deepest over-fetch of some shard (few rows pass the predicate): every shard
        # computes its exact local top-k of passing rows (swat_topk: bank-swap pass, in-pass predicate where needed);
        # top-k of passing rows is associative, so a plain merge finishes it.
        s, r, t, c = _lib.topk(ctx, sub, t2t_bank, k, t2t_threshold, t2i_bank=t2i_bank, t2i_threshold=t2i_threshold,
                               row_offset=row_offset, row_class=sub_rc, exclude=exclude)
        lim = torch.full((len(bad),), float("-inf"), dtype=torch.float32, device=s.device)
        r2 = gather_merge((s, r, t, c, lim), k, world, ctx=ctx, group=group)
    idx = torch.tensor(bad, device=out_s.device)
    out_s[idx], out_r[idx], out_c[idx] = r2[0], r2[1], r2[3]
    if out_t is not None:
        out_t[idx] = r2[2]
    return out_s, out_r, out_t, out_c, torch.zeros_like(out_c)
