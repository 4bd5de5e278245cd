"""Chunked fp32 torch-on-GPU restatement of the oracle's ``topk_walk`` (test infrastructure).

The CPU oracle (``oracle/swat_oracle.py``) finishes 1 M x 200 in seconds but not 10-50 M rows x 1000 classes.  This
restatement does the same arithmetic -- fp32 ``bank @ Q^T`` (cuBLAS SGEMM, TF32 off), per-class reduce over a class's
queries, accept predicate ``T2T >= thr and T2I >= t2i_thr`` (sample_retrieval.py:511-514), first k under
(score desc, row asc) -- with plain torch ops on the device, none of the product's kernels.  ``tests/test_gpu_configs.py``
first proves it equal to ``so.topk_walk`` on 1 M rows, then uses it as the checker at BASELINE.json's full sizes.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch


def _class_scores(x: torch.Tensor, Qt: torch.Tensor, starts: Optional[torch.Tensor], sizes: Optional[torch.Tensor], reduce: str):
    """[n, C] class scores of a row chunk: x [n,512] fp32, Qt [512,Q] fp32."""
    S = x @ Qt
    if reduce == "none":
        return S
    C = starts.numel()
    out = torch.empty(S.shape[0], C, dtype=torch.float32, device=S.device)
    uniq = torch.unique(sizes).tolist()
    for R in uniq:                               # classes with the same group size are reduced together
        cls = (sizes == R).nonzero().flatten()
        cols = (starts[cls][:, None] + torch.arange(R, device=S.device)[None, :]).reshape(-1)
        G = S[:, cols].view(S.shape[0], cls.numel(), R)
        if reduce == "mean":
            v = G.sum(-1) / float(R) if R > 1 else G[..., 0]
        elif reduce == "max":
            v = G.max(-1).values
        else:
            v = G.min(-1).values
        out[:, cls] = v
    return out


@torch.no_grad()
def restate_topk_walk(t2t_bank: torch.Tensor, queries: torch.Tensor, k: int, threshold: float = 0.0,
                      t2i_bank: Optional[torch.Tensor] = None, t2i_threshold: float = 0.25,
                      class_of_query: Optional[torch.Tensor] = None, n_classes: Optional[int] = None, reduce: str = "none",
                      row_labels: Optional[torch.Tensor] = None, chunk: int = 1 << 18):
    """Same contract as ``so.topk_walk``: ``rows [C,k] int64 (-1 padded), t2t [C,k], t2i [C,k] | None, counts [C]``
    (numpy).  Inputs live on the GPU; ``queries`` is fp32 ``[Q,512]``."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        dev = t2t_bank.device
        Q = queries.to(dev, torch.float32)
        Qt = Q.t().contiguous()
        nq = Q.shape[0]
        if class_of_query is None:
            coq = torch.arange(nq, device=dev)
        else:
            coq = torch.as_tensor(class_of_query).to(dev).long()
        C = int(n_classes if n_classes is not None else int(coq.max()) + 1)
        starts = sizes = None
        if reduce != "none":
            sizes = torch.bincount(coq, minlength=C)
            starts = torch.cumsum(sizes, 0) - sizes
        N = t2t_bank.shape[0]
        pool_s = [[] for _ in range(C)]
        pool_r = [[] for _ in range(C)]
        pool_i = [[] for _ in range(C)]
        NEG = torch.tensor(float("-inf"), device=dev)
        for s0 in range(0, N, chunk):
            s1 = min(N, s0 + chunk)
            S = _class_scores(t2t_bank[s0:s1].float(), Qt, starts, sizes, reduce)
            ok = S >= threshold
            I = None
            if t2i_bank is not None:
                I = _class_scores(t2i_bank[s0:s1].float(), Qt, starts, sizes, reduce)
                ok &= I >= t2i_threshold
            if row_labels is not None:
                lab = row_labels[s0:s1].to(dev).long()
                ok &= lab[:, None] == torch.arange(C, device=dev)[None, :]
            M = torch.where(ok, S, NEG)
            kk = min(k, s1 - s0)
            kth = torch.topk(M, kk, dim=0).values[-1]                    # [C] k-th best eligible score of the chunk (or -inf)
            keep = ok & (M >= kth[None, :])                              # every tie of the k-th score stays in
            r, c = keep.nonzero(as_tuple=True)
            sc = S[r, c]
            ic = I[r, c] if I is not None else None
            order = torch.argsort(c, stable=True)
            r, c, sc = r[order], c[order], sc[order]
            if ic is not None:
                ic = ic[order]
            bounds = torch.searchsorted(c, torch.arange(C + 1, device=dev)).tolist()
            r = (r + s0).cpu().numpy(); sc = sc.cpu().numpy(); ic = None if ic is None else ic.cpu().numpy()
            for ci in range(C):
                a, b = bounds[ci], bounds[ci + 1]
                if b > a:
                    pool_r[ci].append(r[a:b]); pool_s[ci].append(sc[a:b])
                    if ic is not None:
                        pool_i[ci].append(ic[a:b])
        rows = np.full((C, k), -1, dtype=np.int64)
        out_s = np.zeros((C, k), dtype=np.float32)
        out_i = None if t2i_bank is None else np.zeros((C, k), dtype=np.float32)
        counts = np.zeros(C, dtype=np.int32)
        for ci in range(C):
            if not pool_r[ci]:
                continue
            idx = np.concatenate(pool_r[ci]); sc = np.concatenate(pool_s[ci])
            order = np.lexsort((idx, -sc.astype(np.float64)))[:k]
            n = order.size
            rows[ci, :n] = idx[order]; out_s[ci, :n] = sc[order]; counts[ci] = n
            if out_i is not None:
                out_i[ci, :n] = np.concatenate(pool_i[ci])[order]
        return rows, out_s, out_i, counts
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def compare_walks(got, ref, tie_tol: float, boundary_tol: float = 1e-3, what: str = "", aux_thr: Optional[float] = None,
                  flip_tol: float = 2e-6, thr: Optional[float] = None):
    """``got`` = (scores, rows, t2i | None, counts) from the product (device or host tensors), ``ref`` = (rows, scores,
    t2i | None, counts) from an oracle (numpy).  The north star's parity rule per class:

    * same counts, scores within 1e-3, each side descending;
    * rows on one side only are either at the k-th boundary (score within ``boundary_tol`` of the last accepted score)
      or *predicate flips*: their T2I score (or, with ``thr``, their T2T score) sits within ``flip_tol`` of the
      threshold, so two fp32 summation orders legitimately disagree on ``t2i >= 0.25`` (each flip also moves one row
      across the k-th boundary);
    * the rows both sides accepted appear in the same order up to swaps among scores that agree to ``tie_tol``.

    Returns (interior swaps, boundary differences, predicate flips, positions compared)."""
    g_s, g_r, g_t, g_c = [None if x is None else (x.cpu().numpy() if torch.is_tensor(x) else np.asarray(x)) for x in got]
    r_r, r_s, r_t, r_c = ref
    r_c = np.asarray(r_c)
    swaps = boundary = flips = total = 0
    for c in range(g_r.shape[0]):
        n, m = int(g_c[c]), int(r_c[c])
        total += n
        assert np.all(g_r[c, n:] == -1), f"{what} class {c}: padding"
        a, b = g_r[c, :n], r_r[c, :m]
        assert np.all(np.diff(g_s[c, :n]) <= 0), f"{what} class {c}: scores not descending"
        if n == m and np.array_equal(a, b):
            np.testing.assert_allclose(g_s[c, :n], r_s[c, :m], atol=1e-3, err_msg=f"{what} class {c} scores")
            continue
        sa = dict(zip(a.tolist(), g_s[c, :n].tolist())); sb = dict(zip(b.tolist(), r_s[c, :m].tolist()))
        only_a, only_b = set(sa) - set(sb), set(sb) - set(sa)
        flipped = set()
        if aux_thr is not None and g_t is not None and r_t is not None:
            ta = dict(zip(a.tolist(), g_t[c, :n].tolist())); tb = dict(zip(b.tolist(), r_t[c, :m].tolist()))
            flipped = {r for r in only_a if abs(ta[r] - aux_thr) <= flip_tol} | {r for r in only_b if abs(tb[r] - aux_thr) <= flip_tol}
        if thr is not None:            # same for the T2T threshold: a score within flip_tol of it may be accepted by one side only
            flipped |= {r for r in only_a if abs(sa[r] - thr) <= flip_tol} | {r for r in only_b if abs(sb[r] - thr) <= flip_tol}
        assert abs(n - m) <= len(flipped), f"{what} class {c}: count {n} != {m}"
        last = min(float(g_s[c, n - 1]) if n else 1.0, float(r_s[c, m - 1]) if m else 1.0)
        for r in (only_a | only_b) - flipped:
            s_ = sa.get(r, sb.get(r))
            assert abs(s_ - last) <= boundary_tol, f"{what} class {c}: row {r} (score {s_}) differs away from the k-th boundary ({last})"
        boundary += (len((only_a | only_b) - flipped) + 1) // 2
        flips += len(flipped)
        common = set(sa) & set(sb)
        a2 = [r for r in a.tolist() if r in common]; b2 = [r for r in b.tolist() if r in common]
        for x, y in zip(a2, b2):
            if x != y:
                swaps += 1
                assert abs(sa[x] - sb[y]) <= tie_tol and abs(sa[x] - sa[y]) <= tie_tol, \
                    f"{what} class {c}: rows {x} / {y} out of order beyond a near-tie ({sa[x]} vs {sb[y]})"
        for r in common:
            assert abs(sa[r] - sb[r]) <= 1e-3, f"{what} class {c}: score of row {r}"
    return swaps, boundary, flips, total
