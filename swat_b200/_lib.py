"""ctypes binding of ``libswat_b200.so`` (C-ABI in ``include/swat_b200.h``).

PyTorch is used for device memory and streams only; every compute call goes through the C-ABI with
raw pointers.  There is no CPU fallback: a missing library or a missing CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libswat_b200.so")

BF16, F32 = 0, 1
REDUCE = {"none": 0, "mean": 1, "max": 2, "min": 3}
ENGINE = {"auto": 0, "tc": 1, "simt": 2}
DIM = 512

EXPORTS = [
    "swat_version", "swat_last_error", "swat_ctx_create", "swat_ctx_destroy", "swat_ctx_set_option",
    "swat_ctx_launch_count", "swat_bank_load", "swat_queries_create", "swat_queries_destroy", "swat_job_create", "swat_job_reset",
    "swat_job_set_class_depth", "swat_job_scan", "swat_job_select", "swat_job_export_flags", "swat_job_status", "swat_job_destroy", "swat_scan_eps", "swat_rescore_walk", "swat_merge_topk",
    "swat_scores_dense", "swat_score_rows", "swat_zeroshot_predict", "swat_near_duplicates", "swat_topk", "swat_topk_host", "swat_ctx_last_timing",
]


class SwatError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"swat_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    """Load the shared library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python swat_b200/csrc/build.py` "
                           "(swat_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    lib.swat_version.restype = i32
    lib.swat_last_error.restype = C.c_char_p
    lib.swat_ctx_launch_count.restype = i64
    lib.swat_ctx_launch_count.argtypes = [vp]
    sig = {
        "swat_ctx_create": [i32, C.POINTER(vp)],
        "swat_ctx_destroy": [vp],
        "swat_ctx_set_option": [vp, C.c_char_p, i64],
        "swat_ctx_last_timing": [vp, C.POINTER(C.c_double)],
        "swat_bank_load": [vp, C.c_char_p, i32, i64, i64, vp, i64, C.POINTER(i32), vp],
        "swat_queries_create": [vp, vp, i32, vp, i32, i32, C.POINTER(vp)],
        "swat_queries_destroy": [vp],
        "swat_job_create": [vp, vp, i32, f32, C.POINTER(vp)],
        "swat_job_reset": [vp, vp],
        "swat_job_set_class_depth": [vp, vp, vp],
        "swat_job_scan": [vp, vp, i32, i64, i64, vp, f32, vp, vp, i32, vp],
        "swat_job_select": [vp, i64, vp, vp, vp, vp, vp],
        "swat_job_export_flags": [vp, vp, vp],
        "swat_job_status": [vp, C.POINTER(i32)],
        "swat_job_destroy": [vp],
        "swat_scan_eps": [vp, i32, i32, C.POINTER(f32)],
        "swat_rescore_walk": [vp, vp, vp, vp, vp, i32, i64, i64, vp, vp, vp, vp, i32, i32, f32, f32, f32, vp, vp, vp, vp, vp, vp, vp],
        "swat_merge_topk": [vp, vp, vp, vp, vp, vp, i32, i64, i32, i32, i32, f32, vp, vp, vp, vp, vp, vp],
        "swat_scores_dense": [vp, vp, vp, i32, i64, vp, i32, vp],
        "swat_score_rows": [vp, vp, vp, i32, i64, vp, vp, vp],
        "swat_zeroshot_predict": [vp, vp, vp, i32, i64, vp, i32, vp],
        "swat_near_duplicates": [vp, vp, i32, i64, vp, vp, i32, i32, f32, vp, vp],
        "swat_topk": [vp, vp, vp, vp, i32, i64, i64, i32, f32, f32, vp, vp, vp, vp, vp, vp, vp],
        "swat_topk_host": [vp, vp, vp, vp, i32, i64, i64, i32, f32, f32, vp, vp, vp, vp, vp, vp],
    }
    for name, args in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = i32
    _lib = lib
    return lib


def _check(rc: int):
    if rc != 0:
        raise SwatError(rc, load().swat_last_error().decode("utf-8", "replace"))


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float32:
        return F32
    raise TypeError(f"bank dtype must be bfloat16 or float32, got {t.dtype}")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _bank_ok(t: torch.Tensor, name: str, cuda: bool):
    if t.dim() != 2 or t.shape[1] != DIM:
        raise ValueError(f"{name} must be [N, {DIM}], got {tuple(t.shape)}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    if cuda != t.is_cuda:
        raise ValueError(f"{name} must live on {'a CUDA device' if cuda else 'the host'}")


class Context:
    """One per device (``swat_ctx``)."""

    def __init__(self, device: int = 0, **options):
        lib = load()
        if not torch.cuda.is_available():
            raise RuntimeError("swat_b200 needs a CUDA device (sm_100); there is no CPU fallback")
        self.device = int(device)
        self._h = C.c_void_p()
        _check(lib.swat_ctx_create(self.device, C.byref(self._h)))
        for k, v in options.items():
            self.set_option(k, v)

    def set_option(self, name: str, value: int):
        _check(load().swat_ctx_set_option(self._h, name.encode(), int(value)))

    @property
    def launch_count(self) -> int:
        return int(load().swat_ctx_launch_count(self._h))

    def last_timing(self) -> dict:
        out = (C.c_double * 8)()
        _check(load().swat_ctx_last_timing(self._h, out))
        keys = ["scan_ms", "select_ms", "t2i_ms", "total_ms", "scan_launches", "h2d_bytes", "d2h_bytes", "escalations"]
        return dict(zip(keys, [float(x) for x in out]))

    def close(self):
        if self._h:
            load().swat_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Queries:
    """Prompt tensors on the device (``swat_queries``): ``queries [Q,512]`` fp32 (any device; copied
    to the host), ``class_of_query [Q]`` non-decreasing dense class index, ``reduce`` in
    none|mean|max|min."""

    def __init__(self, ctx: Context, queries, class_of_query=None, n_classes: Optional[int] = None, reduce="none"):
        q = torch.as_tensor(queries).detach().to("cpu", torch.float32).contiguous()
        if q.dim() != 2 or q.shape[1] != DIM:
            raise ValueError(f"queries must be [Q, {DIM}]")
        self.ctx = ctx
        self.n_queries = int(q.shape[0])
        self.host_queries = q                      # kept for building sub-query sets (targeted escalation)
        coq = None
        if class_of_query is not None:
            coq = torch.as_tensor(class_of_query).detach().to("cpu", torch.int32).contiguous()
            if coq.numel() != self.n_queries:
                raise ValueError("class_of_query must have one entry per query")
        self.n_classes = int(n_classes if n_classes is not None else (self.n_queries if coq is None else int(coq.max()) + 1))
        self.reduce = REDUCE[reduce] if isinstance(reduce, str) else int(reduce)
        self.host_class_of_query = coq if coq is not None else torch.arange(self.n_queries, dtype=torch.int32)
        self._h = C.c_void_p()
        _check(load().swat_queries_create(ctx._h, _ptr(q), self.n_queries, _ptr(coq), self.n_classes, self.reduce,
                                          C.byref(self._h)))

    def subset(self, classes: Sequence[int]) -> "Queries":
        """Query set restricted to ``classes`` (renumbered 0..len-1, same order); cached per class tuple."""
        key = tuple(int(c) for c in classes)
        cache = self.__dict__.setdefault("_subsets", {})
        if key not in cache:
            coq = self.host_class_of_query
            sel = torch.cat([(coq == c).nonzero().flatten() for c in key])
            new_coq = torch.repeat_interleave(torch.arange(len(key), dtype=torch.int32),
                                              torch.tensor([int((coq == c).sum()) for c in key]))
            if len(cache) >= 4:
                cache.pop(next(iter(cache))).close()
            cache[key] = Queries(self.ctx, self.host_queries[sel], new_coq, len(key), self.reduce)
        return cache[key]

    def close(self):
        for sub in self.__dict__.get("_subsets", {}).values():
            sub.close()
        self.__dict__["_subsets"] = {}
        for job in self.__dict__.get("_job_cache", {}).values():      # swat_b200.dist keeps its streaming jobs here
            job.close()
        self.__dict__["_job_cache"] = {}
        if self._h:
            load().swat_queries_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Job:
    """Streaming per-class top-``k_fetch`` (``swat_job``)."""

    def __init__(self, ctx: Context, queries: Queries, k_fetch: int, t2t_threshold: float = 0.0):
        self.ctx, self.queries, self.k_fetch = ctx, queries, int(k_fetch)
        self._h = C.c_void_p()
        _check(load().swat_job_create(ctx._h, queries._h, self.k_fetch, float(t2t_threshold), C.byref(self._h)))

    def reset(self):
        _check(load().swat_job_reset(self._h, _stream(self.ctx.device)))

    def set_class_depth(self, depth=None):
        """Per-class depth (host int32 ``[C]``, each in ``[1, k_fetch]``); ``None`` = uniform ``k_fetch``."""
        d = None if depth is None else torch.as_tensor(depth).detach().to("cpu", torch.int32).contiguous()
        _check(load().swat_job_set_class_depth(self._h, _ptr(d), _stream(self.ctx.device)))

    def scan(self, bank: torch.Tensor, row_base: int = 0, t2i_bank: Optional[torch.Tensor] = None,
             t2i_threshold: float = 0.25, row_class: Optional[torch.Tensor] = None,
             exclude: Optional[torch.Tensor] = None, engine="auto"):
        _bank_ok(bank, "bank", True)
        if t2i_bank is not None:
            _bank_ok(t2i_bank, "t2i_bank", True)
            if t2i_bank.shape != bank.shape or t2i_bank.dtype != bank.dtype:
                raise ValueError("t2i_bank must match bank in shape and dtype")
        if row_class is not None and (row_class.dtype != torch.int32 or row_class.numel() != bank.shape[0] or not row_class.is_cuda):
            raise ValueError("row_class must be a CUDA int32 tensor with one entry per row")
        if exclude is not None and (exclude.dtype != torch.int32 or exclude.numel() * 32 < bank.shape[0] or not exclude.is_cuda):
            raise ValueError("exclude must be a CUDA int32 bitmap covering every row")
        _check(load().swat_job_scan(self._h, _ptr(bank), _dtype_code(bank), int(bank.shape[0]), int(row_base), _ptr(t2i_bank),
                                    float(t2i_threshold), _ptr(row_class), _ptr(exclude), ENGINE[engine],
                                    _stream(self.ctx.device)))

    def select(self, row_offset: int = 0, out=None):
        """Sorted top-k_fetch per class; ``out = (scores, rows, counts, trunc)`` writes into caller tensors
        (e.g. slices of a packed exchange buffer)."""
        dev = torch.device("cuda", self.ctx.device)
        Cn, kf = self.queries.n_classes, self.k_fetch
        if out is None:
            scores = torch.empty(Cn, kf, dtype=torch.float32, device=dev)
            rows = torch.empty(Cn, kf, dtype=torch.int64, device=dev)
            counts = torch.empty(Cn, dtype=torch.int32, device=dev)
            trunc = torch.empty(Cn, dtype=torch.int32, device=dev)
        else:
            scores, rows, counts, trunc = out
            _out_ok(scores, (Cn, kf), torch.float32); _out_ok(rows, (Cn, kf), torch.int64)
            _out_ok(counts, (Cn,), torch.int32); _out_ok(trunc, (Cn,), torch.int32)
        _check(load().swat_job_select(self._h, int(row_offset), _ptr(scores), _ptr(rows), _ptr(counts), _ptr(trunc),
                                      _stream(self.ctx.device)))
        return scores, rows, counts, trunc

    def export_flags(self, out: torch.Tensor):
        """Stream-ordered copy of the overflow word into ``out`` (one int32 on the device): no host sync."""
        _out_ok(out, (1,), torch.int32)
        _check(load().swat_job_export_flags(self._h, _ptr(out), _stream(self.ctx.device)))

    def overflowed(self) -> int:
        """0 = fine; bit0 = class candidate buffers ("cand_cap"), bit1 = survivor lists ("list_entries")."""
        o = C.c_int32(0)
        _check(load().swat_job_status(self._h, C.byref(o)))
        return int(o.value)

    def close(self):
        if self._h:
            load().swat_job_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def bank_load(ctx: Context, path: str, dtype: torch.dtype, row_begin: int, row_end: int, chunk_rows: int = 0):
    """Rows ``[row_begin, row_end)`` of a flat shard file -> a new ``[n,512]`` device tensor (``swat_bank_load``).
    Returns ``(tensor, used_gds)``."""
    dev = torch.device("cuda", ctx.device)
    out = torch.empty(int(row_end) - int(row_begin), DIM, dtype=dtype, device=dev)
    gds = C.c_int32(0)
    with torch.cuda.device(dev):
        _check(load().swat_bank_load(ctx._h, os.fsencode(path), _dtype_code(out), int(row_begin), int(row_end), _ptr(out), int(chunk_rows),
                                     C.byref(gds), _stream(ctx.device)))
    return out, bool(gds.value)


def scores_dense(ctx: Context, queries: Queries, bank: torch.Tensor, engine="auto") -> torch.Tensor:
    """``[N, C]`` class scores (S1 compatibility / tests)."""
    _bank_ok(bank, "bank", True)
    out = torch.empty(bank.shape[0], queries.n_classes, dtype=torch.float32, device=bank.device)
    _check(load().swat_scores_dense(ctx._h, queries._h, _ptr(bank), _dtype_code(bank), int(bank.shape[0]), _ptr(out),
                                    ENGINE[engine], _stream(ctx.device)))
    return out


def score_rows(ctx: Context, queries: Queries, bank: torch.Tensor, row_class: torch.Tensor) -> torch.Tensor:
    """``[N]`` canonical score of every row against the queries of its own class (``row_class`` int32, -1 = none)."""
    _bank_ok(bank, "bank", True)
    if row_class.dtype != torch.int32 or row_class.numel() != bank.shape[0] or not row_class.is_cuda:
        raise ValueError("row_class must be a CUDA int32 tensor with one entry per row")
    out = torch.empty(bank.shape[0], dtype=torch.float32, device=bank.device)
    _check(load().swat_score_rows(ctx._h, queries._h, _ptr(bank), _dtype_code(bank), int(bank.shape[0]), _ptr(row_class), _ptr(out),
                                  _stream(ctx.device)))
    return out


def zeroshot_predict(ctx: Context, queries: Queries, bank: torch.Tensor, engine="auto") -> torch.Tensor:
    """``[N]`` int32: argmax over the class scores of every row (the zero-shot head's prediction)."""
    _bank_ok(bank, "bank", True)
    out = torch.empty(bank.shape[0], dtype=torch.int32, device=bank.device)
    _check(load().swat_zeroshot_predict(ctx._h, queries._h, _ptr(bank), _dtype_code(bank), int(bank.shape[0]), _ptr(out),
                                        ENGINE[engine], _stream(ctx.device)))
    return out


def near_duplicates(ctx: Context, bank: torch.Tensor, order: torch.Tensor, class_start: torch.Tensor, threshold: float = 0.9) -> torch.Tensor:
    """``dup[p] = 1`` iff row ``order[p]`` has an earlier row of its class with cosine > threshold."""
    _bank_ok(bank, "bank", True)
    dev = bank.device
    order = order.to(dev, torch.int64).contiguous()
    cs = class_start.to(torch.int32).cpu()
    n_classes = int(cs.numel()) - 1
    max_rows = int((cs[1:] - cs[:-1]).max()) if n_classes > 0 else 0
    d_cs = cs.to(dev)
    dup = torch.zeros(order.numel(), dtype=torch.uint8, device=dev)
    _check(load().swat_near_duplicates(ctx._h, _ptr(bank), _dtype_code(bank), int(bank.shape[0]), _ptr(order), _ptr(d_cs), n_classes,
                                       max_rows, float(threshold), _ptr(dup), _stream(ctx.device)))
    return dup


def topk(ctx: Context, queries: Queries, t2t_bank: torch.Tensor, k: int, t2t_threshold: float = 0.0,
         t2i_bank: Optional[torch.Tensor] = None, t2i_threshold: float = 0.25, row_class: Optional[torch.Tensor] = None,
         exclude: Optional[torch.Tensor] = None, row_offset: int = 0, out=None):
    """Whole pipeline on HBM-resident banks.  Returns ``(scores [C,k], rows [C,k] int64, t2i [C,k] | None,
    counts [C] int32)`` on the device."""
    _bank_ok(t2t_bank, "t2t_bank", True)
    if t2i_bank is not None:
        _bank_ok(t2i_bank, "t2i_bank", True)
        if t2i_bank.shape != t2t_bank.shape or t2i_bank.dtype != t2t_bank.dtype:
            raise ValueError("t2i_bank must match t2t_bank in shape and dtype")
    dev = t2t_bank.device
    Cn = queries.n_classes
    if out is None:
        scores = torch.empty(Cn, k, dtype=torch.float32, device=dev)
        rows = torch.empty(Cn, k, dtype=torch.int64, device=dev)
        t2i = torch.empty(Cn, k, dtype=torch.float32, device=dev) if t2i_bank is not None else None
        counts = torch.empty(Cn, dtype=torch.int32, device=dev)
    else:
        scores, rows, t2i, counts = out
    _check(load().swat_topk(ctx._h, queries._h, _ptr(t2t_bank), _ptr(t2i_bank), _dtype_code(t2t_bank), int(t2t_bank.shape[0]),
                            int(row_offset), int(k), float(t2t_threshold), float(t2i_threshold), _ptr(row_class), _ptr(exclude),
                            _ptr(scores), _ptr(rows), _ptr(t2i), _ptr(counts), _stream(ctx.device)))
    return scores, rows, t2i, counts


def topk_host(ctx: Context, queries: Queries, t2t_bank: torch.Tensor, k: int, t2t_threshold: float = 0.0,
              t2i_bank: Optional[torch.Tensor] = None, t2i_threshold: float = 0.25, row_class: Optional[torch.Tensor] = None,
              exclude: Optional[torch.Tensor] = None, row_offset: int = 0, out=None):
    """Whole pipeline on HOST banks (the reference's ``torch.load``-ed CPU tensors).  Pinned tensors
    stream at full PCIe rate.  Returns host tensors."""
    _bank_ok(t2t_bank, "t2t_bank", False)
    if t2i_bank is not None:
        _bank_ok(t2i_bank, "t2i_bank", False)
    Cn = queries.n_classes
    if out is None:
        scores = torch.empty(Cn, k, dtype=torch.float32)
        rows = torch.empty(Cn, k, dtype=torch.int64)
        t2i = torch.empty(Cn, k, dtype=torch.float32) if t2i_bank is not None else None
        counts = torch.empty(Cn, dtype=torch.int32)
    else:
        scores, rows, t2i, counts = out
    _check(load().swat_topk_host(ctx._h, queries._h, _ptr(t2t_bank), _ptr(t2i_bank), _dtype_code(t2t_bank), int(t2t_bank.shape[0]),
                                 int(row_offset), int(k), float(t2t_threshold), float(t2i_threshold), _ptr(row_class),
                                 _ptr(exclude), _ptr(scores), _ptr(rows), _ptr(t2i), _ptr(counts)))
    return scores, rows, t2i, counts


def _out_ok(t: torch.Tensor, shape, dtype):
    if tuple(t.shape) != tuple(shape) or t.dtype != dtype or not t.is_contiguous() or not t.is_cuda:
        raise ValueError(f"output tensor must be a contiguous CUDA {dtype} tensor of shape {tuple(shape)}")


def scan_eps(queries: Queries, dtype, engine="auto") -> float:
    """Bound on |score a scan ranks a row by - canonical score| for banks of ``dtype`` (torch dtype or code)."""
    code = dtype if isinstance(dtype, int) else (BF16 if dtype == torch.bfloat16 else F32)
    e = C.c_float(0.0)
    _check(load().swat_scan_eps(queries._h, code, ENGINE[engine], C.byref(e)))
    return float(e.value)


def rescore_walk(ctx: Context, queries: Queries, t2t_bank: torch.Tensor, cand_scores, cand_rows, cand_counts, truncated, k: int,
                 t2t_threshold: float = 0.0, aux_bank: Optional[torch.Tensor] = None, aux_threshold: float = 0.25,
                 bank_row_base: int = 0, eps: Optional[float] = None, out=None, aux_queries: Optional[Queries] = None):
    """Exact re-score of the candidates (canonical fp32 dot against ``t2t_bank`` and, when given, ``aux_bank``) +
    accept walk.  Returns ``(scores, rows, aux, counts, limit, incomplete)``; ``out = (scores, rows, aux, counts,
    limit)`` writes into caller tensors (e.g. views of a packed exchange buffer)."""
    _bank_ok(t2t_bank, "t2t_bank", True)
    if aux_bank is not None:
        _bank_ok(aux_bank, "aux_bank", True)
        if aux_bank.shape != t2t_bank.shape or aux_bank.dtype != t2t_bank.dtype:
            raise ValueError("aux_bank must match t2t_bank in shape and dtype")
    dev = t2t_bank.device
    Cn, kf = cand_scores.shape
    if eps is None:
        eps = scan_eps(queries, t2t_bank.dtype)
    if out is None:
        o_s = torch.empty(Cn, k, dtype=torch.float32, device=dev)
        o_r = torch.empty(Cn, k, dtype=torch.int64, device=dev)
        o_t = torch.empty(Cn, k, dtype=torch.float32, device=dev)
        o_c = torch.empty(Cn, dtype=torch.int32, device=dev)
        o_l = torch.empty(Cn, dtype=torch.float32, device=dev)
    else:
        o_s, o_r, o_t, o_c, o_l = out
        _out_ok(o_s, (Cn, k), torch.float32); _out_ok(o_r, (Cn, k), torch.int64)
        _out_ok(o_t, (Cn, k), torch.float32); _out_ok(o_c, (Cn,), torch.int32); _out_ok(o_l, (Cn,), torch.float32)
    o_i = torch.empty(Cn, dtype=torch.int32, device=dev)
    _check(load().swat_rescore_walk(ctx._h, queries._h, None if aux_queries is None else aux_queries._h, _ptr(t2t_bank), _ptr(aux_bank), _dtype_code(t2t_bank), int(t2t_bank.shape[0]),
                                    int(bank_row_base), _ptr(cand_scores), _ptr(cand_rows), _ptr(cand_counts), _ptr(truncated), int(kf),
                                    int(k), float(t2t_threshold), float(aux_threshold), float(eps), _ptr(o_s), _ptr(o_r), _ptr(o_t),
                                    _ptr(o_c), _ptr(o_l), _ptr(o_i), _stream(ctx.device)))
    return o_s, o_r, o_t, o_c, o_l, o_i


def merge_topk(ctx: Context, scores: torch.Tensor, rows: torch.Tensor, counts: torch.Tensor, aux: Optional[torch.Tensor] = None,
               limit: Optional[torch.Tensor] = None, k_out: Optional[int] = None, aux_threshold: float = float("-inf"),
               n_shards: Optional[int] = None, shard_stride_bytes: int = 0):
    """Merge gathered per-shard walk results ``[G,C,k_in]`` (canonical scores, rows global) into ``[C,k_out]``: the
    best ``k_out`` entries with ``aux >= aux_threshold``.  ``limit [G,C]`` float32: shard g vouches only for rows
    scoring above ``limit[g,c]`` (``-inf`` = complete).  Returns ``(scores, rows, aux | None, counts, incomplete)``;
    ``incomplete[c] == 1`` means the result reaches down to some shard's limit (re-run the shards deeper).

    With ``shard_stride_bytes > 0`` the arguments are shard 0's ``[C,k_in]`` / ``[C]`` slices of an
    all-gathered packed buffer and shard g of every array lies ``g * shard_stride_bytes`` further on
    (``n_shards`` then gives G); nothing is copied."""
    if shard_stride_bytes:
        if n_shards is None:
            raise ValueError("n_shards is required with shard_stride_bytes")
        G, (Cn, k_in) = int(n_shards), scores.shape
    else:
        G, Cn, k_in = scores.shape
        scores, rows, counts = scores.contiguous(), rows.contiguous(), counts.contiguous()
        aux = None if aux is None else aux.contiguous()
        limit = None if limit is None else limit.to(torch.float32).contiguous()
    k_out = int(k_in if k_out is None else k_out)
    dev = scores.device
    o_s = torch.empty(Cn, k_out, dtype=torch.float32, device=dev)
    o_r = torch.empty(Cn, k_out, dtype=torch.int64, device=dev)
    o_a = torch.empty(Cn, k_out, dtype=torch.float32, device=dev) if aux is not None else None
    o_c = torch.empty(Cn, dtype=torch.int32, device=dev)
    o_i = torch.empty(Cn, dtype=torch.int32, device=dev)
    _check(load().swat_merge_topk(ctx._h, _ptr(scores), _ptr(rows), _ptr(aux), _ptr(counts), _ptr(limit), int(G),
                                  int(shard_stride_bytes), int(Cn), int(k_in), k_out, float(aux_threshold), _ptr(o_s), _ptr(o_r),
                                  _ptr(o_a), _ptr(o_c), _ptr(o_i), _stream(ctx.device)))
    return o_s, o_r, o_a, o_c, o_i
