"""Seeded synthetic LAION-like feature banks and prompt tensors.

Pure-noise unit vectors never pass ``T2I >= 0.25`` and half of them fail ``T2T >= 0``, so the
banks carry structure: a fraction ``rho`` of the rows is "relevant" to one class with relevance
``a ~ Beta(2, 5)`` (caption side) and ``b ~ clip(0.27 + 0.08 N(0,1))`` (image side); 0.1 % of the
rows duplicate an earlier row and one block of identical rows forces ties at the k-th boundary.
Rows and queries are L2-normalised in fp32 and *then* rounded to the bank dtype; oracles are fed
the rounded values upcast to fp32 so quantisation is not counted as error.

The layout mirrors what ``extract_mined_feature.py:166-168,208,277-279`` writes:
``caption_features [N,512]``, ``image_features [N,512]``, ``labels [N]``, ``filepath`` list.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

DIM = 512


def _unit(x: torch.Tensor) -> torch.Tensor:
    return torch.nn.functional.normalize(x.float(), dim=-1)


def make_queries(n_classes: int, syn_per_class=1, seed: int = 0, device="cpu", dtype=torch.float32,
                 syn_noise: float = 0.3):
    """Class vectors ``q_c`` and synonym vectors ``normalise(q_c + syn_noise * u)``.

    ``syn_per_class`` is an int or a list of per-class group sizes.  Returns
    ``(class_vecs [C,512], queries [Q,512], class_of_query [Q] int32)``; with one synonym per
    class the queries *are* the class vectors (the reference's ``['mean']`` prompt,
    ``sample_retrieval.py:749-750``).
    """
    g = torch.Generator(device="cpu").manual_seed(int(seed) * 7919 + 17)
    qc = _unit(torch.randn(n_classes, DIM, generator=g))
    if isinstance(syn_per_class, int):
        sizes = [syn_per_class] * n_classes
    else:
        sizes = list(syn_per_class)
    if all(s == 1 for s in sizes):
        queries = qc.clone()
        coq = torch.arange(n_classes, dtype=torch.int32)
    else:
        coq = torch.repeat_interleave(torch.arange(n_classes, dtype=torch.int32), torch.tensor(sizes))
        u = _unit(torch.randn(coq.numel(), DIM, generator=g))
        queries = _unit(qc[coq.long()] + syn_noise * u)
    qc = qc.to(dtype).to(device)
    queries = queries.to(dtype).to(device)
    return qc, queries, coq.to(device)


def make_bank(n_rows: int, class_vecs: torch.Tensor, seed: int = 0, device="cpu", dtype=torch.bfloat16,
              rho: float = 0.05, dup_frac: float = 0.001, tie_block: int = 1000,
              partitioned: bool = False, zipf_s: float = 1.0, chunk: int = 1 << 20,
              with_images: bool = True, row_offset: int = 0):
    """Generate ``caption [N,512]``, ``image [N,512] or None`` and ``labels [N] int64``.

    ``labels`` is the class a relevant row was drawn for (or, when ``partitioned``, the class
    folder of every row, Zipf-distributed); kernels ignore it in unpartitioned mode.
    Generation is chunked and seeded per chunk (``seed, row_offset + chunk_start``) so a shard
    generated on rank r equals the matching slice of the single-GPU bank whenever ``row_offset`` is
    a multiple of ``chunk`` -- the shard-count-invariance tests rely on it.  ``chunk`` must then be
    the same on every rank.
    """
    dev = torch.device(device)
    C = class_vecs.shape[0]
    qc = class_vecs.to(dev).float()
    cap = torch.empty(n_rows, DIM, dtype=dtype, device=dev)
    img = torch.empty(n_rows, DIM, dtype=dtype, device=dev) if with_images else None
    labels = torch.empty(n_rows, dtype=torch.int64, device=dev)
    if partitioned:
        w = 1.0 / torch.arange(1, C + 1, dtype=torch.float64) ** zipf_s
        w = (w / w.sum()).to(dev)
    for s0 in range(0, n_rows, chunk):
        s1 = min(n_rows, s0 + chunk)
        n = s1 - s0
        g = torch.Generator(device=dev).manual_seed((int(seed) * 1000003 + (row_offset + s0)) % (2 ** 63 - 1))
        if partitioned:
            lab = torch.multinomial(w.float(), n, replacement=True, generator=g)
        else:
            lab = torch.randint(0, C, (n,), generator=g, device=dev)
        relevant = torch.rand(n, generator=g, device=dev) < rho
        # Beta(2,5) via order statistics of uniforms is awkward on GPU; use the Gamma ratio with
        # torch's sampler-free construction: Beta(2,5) == 2nd smallest of 6 uniforms.
        u6 = torch.rand(n, 6, generator=g, device=dev)
        a = torch.sort(u6, dim=-1).values[:, 1]
        a = torch.where(relevant, a, torch.zeros_like(a))
        b = (0.27 + 0.08 * torch.randn(n, generator=g, device=dev)).clamp_(0.0, 1.0)
        b = torch.where(relevant, b, torch.zeros_like(b))
        base = qc[lab]
        u = _unit(torch.randn(n, DIM, generator=g, device=dev))
        x = _unit(a[:, None] * base + torch.sqrt(1.0 - a * a)[:, None] * u)
        # duplicates of an earlier row in the same chunk (exact ties, lowest index must win)
        n_dup = int(n * dup_frac)
        if n_dup > 0 and n > 2:
            dst = torch.randint(1, n, (n_dup,), generator=g, device=dev)
            src = (torch.rand(n_dup, generator=g, device=dev) * dst.float()).long().clamp_(0, n - 1)
            src = torch.minimum(src, dst - 1)
        else:
            dst = src = None
        if with_images:
            v = _unit(torch.randn(n, DIM, generator=g, device=dev))
            y = _unit(b[:, None] * base + torch.sqrt(1.0 - b * b)[:, None] * v)
        if dst is not None:
            x[dst] = x[src]
            lab[dst] = lab[src]
            if with_images:
                y[dst] = y[src]
        cap[s0:s1] = x.to(dtype)
        if with_images:
            img[s0:s1] = y.to(dtype)
        labels[s0:s1] = lab
    # one block of identical relevant rows (ties at and around the k-th boundary)
    if tie_block > 0 and n_rows >= 4 * tie_block and row_offset == 0:
        t0 = n_rows // 3
        g = torch.Generator(device=dev).manual_seed(int(seed) * 31 + 5)
        c = int(torch.randint(0, C, (1,), generator=g, device=dev).item())
        u = _unit(torch.randn(DIM, generator=g, device=dev))
        a0, b0 = 0.45, 0.30
        row = _unit(a0 * qc[c] + math.sqrt(1 - a0 * a0) * u)
        cap[t0:t0 + tie_block] = row.to(dtype)
        if with_images:
            v = _unit(torch.randn(DIM, generator=g, device=dev))
            irow = _unit(b0 * qc[c] + math.sqrt(1 - b0 * b0) * v)
            img[t0:t0 + tie_block] = irow.to(dtype)
        labels[t0:t0 + tie_block] = c
    return cap, img, labels


def make_paths(labels, root: str = "/scratch/retrieved/synthetic", class_ids=None):
    """``filepath`` entries in the reference's ``<root>/<cls>/<id>.jpg`` form
    (``utils/datasets/dataset_utils.py:303-308``) and the matching caption map
    ``{cls: {img_id: caption}}`` (``retrieval/process_meta_map.py:5-47``)."""
    lab = labels.tolist() if hasattr(labels, "tolist") else list(labels)
    paths, cmap = [], {}
    for i, l in enumerate(lab):
        cid = str(int(l) if class_ids is None else class_ids[int(l)])
        paths.append(f"{root}/{cid}/{i}.jpg")
        cmap.setdefault(cid, {})[str(i)] = f"synthetic caption {i}"
    return paths, cmap


def plant_needles(cap: torch.Tensor, class_vecs: torch.Tensor, classes, per_class: int, seed: int = 0,
                  img: Optional[torch.Tensor] = None):
    """Overwrite ``per_class`` random rows per listed class with rows whose T2T cosine to the class
    vector is a known, strictly decreasing ladder in (0.90, 0.99) -- above everything the generator
    produces -- so the expected top of each class is known in closed form at any bank size.
    Returns ``{class: rows (LongTensor, in expected rank order)}``."""
    dev = cap.device
    g = torch.Generator(device="cpu").manual_seed(int(seed) * 104729 + 3)
    n = cap.shape[0]
    total = per_class * len(classes)
    rows = torch.randperm(n, generator=g)[:total].view(len(classes), per_class)
    out = {}
    for ci, c in enumerate(classes):
        q = class_vecs[c].float().to(dev)
        r = rows[ci].to(dev)
        cosines = torch.linspace(0.99, 0.90, per_class, device=dev)
        u = torch.randn(per_class, DIM, generator=g).to(dev)
        u = _unit(u - (u @ q)[:, None] * q[None, :])
        x = cosines[:, None] * q[None, :] + torch.sqrt(1 - cosines ** 2)[:, None] * u
        cap[r] = _unit(x).to(cap.dtype)
        if img is not None:
            y = 0.5 * q[None, :] + math.sqrt(0.75) * u
            img[r] = _unit(y).to(img.dtype)
        out[int(c)] = r.cpu()
    return out
