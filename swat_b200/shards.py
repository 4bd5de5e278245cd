"""Feature-shard loading for the retrieval hot path.

Two on-disk layouts:

* the reference's ``{root}/{dataset}/{dataset}_{model_cfg}_mined.pth`` written by
  ``extract_mined_feature.py:277-279`` -- a ``torch.save`` dict ``{caption_features [N,512],
  image_features [N,512], labels [N], filepath list[str]}`` (fields built at ``:166-168, :208``) --
  read here with ``mmap=True`` so the two matrices are zero-copy host views;
* a flat shard directory (``meta.json`` + raw row-major ``caption.bin`` / ``image.bin`` /
  ``labels.i64`` + ``paths.txt``) that ``np.memmap``s in O(1) and streams straight into pinned
  staging buffers; at LAION scale the pickled N-string list of the ``.pth`` does not scale.

No compute happens here: these are the inputs of ``swat_topk_host`` / ``swat_topk``.
"""
from __future__ import annotations

import json
import os
import warnings
from typing import Optional

import numpy as np
import torch

DIM = 512
_NP = {"bf16": np.uint16, "f32": np.float32}
_TORCH = {"bf16": torch.bfloat16, "f32": torch.float32}


def load_mined_pth(path: str, mmap: bool = True) -> dict:
    """``torch.load`` of the reference's mined-feature file (``sample_retrieval.py:1473-1476``).
    Tensors saved from CUDA are mapped to the host."""
    try:
        return torch.load(path, map_location="cpu", mmap=mmap, weights_only=False)
    except (RuntimeError, ValueError):
        return torch.load(path, map_location="cpu", weights_only=False)     # legacy (non-zip) files cannot be mmapped


def save_mined_pth(path: str, caption_features, image_features, labels, filepath):
    """Write the reference layout (``extract_mined_feature.py:277-279``)."""
    torch.save({"caption_features": caption_features, "labels": labels, "filepath": list(filepath),
                "image_features": image_features}, path)


def write_flat_shard(out_dir: str, caption_features: torch.Tensor, image_features: Optional[torch.Tensor], labels,
                     filepath, dtype: str = "bf16", row_offset: int = 0):
    """Convert to the flat layout.  ``dtype`` = ``bf16`` (round-to-nearest-even, the tcgen05 path)
    or ``f32`` (bit-exact copy of the reference's fp32 features)."""
    os.makedirs(out_dir, exist_ok=True)
    n = int(caption_features.shape[0])

    def dump(t, name):
        t = torch.as_tensor(t).detach().cpu().to(_TORCH[dtype]).contiguous()
        arr = t.view(torch.int16).numpy() if dtype == "bf16" else t.numpy()
        arr.tofile(os.path.join(out_dir, name))

    dump(caption_features, "caption.bin")
    if image_features is not None:
        dump(image_features, "image.bin")
    np.asarray(torch.as_tensor(labels).cpu(), dtype=np.int64).tofile(os.path.join(out_dir, "labels.i64"))
    with open(os.path.join(out_dir, "paths.txt"), "w") as f:
        f.write("\n".join(filepath))
    meta = {"n_rows": n, "dim": DIM, "dtype": dtype, "row_offset": int(row_offset), "has_images": image_features is not None,
            "format": "swat_b200 flat shard v1"}
    with open(os.path.join(out_dir, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    return meta


def convert_pth_to_flat(pth_path: str, out_dir: str, dtype: str = "bf16"):
    d = load_mined_pth(pth_path)
    return write_flat_shard(out_dir, d["caption_features"], d.get("image_features"), d["labels"], d["filepath"], dtype)


class FlatShard:
    """Memory-mapped flat shard; ``as_mined_dict()`` gives the reference's dict layout with zero-copy
    host tensors (bf16 banks are exposed as ``torch.bfloat16`` views of the mapped bytes)."""

    def __init__(self, path: str):
        self.path = path
        self.meta = json.load(open(os.path.join(path, "meta.json")))
        self.n_rows, self.dtype = int(self.meta["n_rows"]), self.meta["dtype"]
        self.row_offset = int(self.meta.get("row_offset", 0))
        shape = (self.n_rows, DIM)
        self._cap = np.memmap(os.path.join(path, "caption.bin"), dtype=_NP[self.dtype], mode="r", shape=shape)
        self._img = (np.memmap(os.path.join(path, "image.bin"), dtype=_NP[self.dtype], mode="r", shape=shape)
                     if self.meta.get("has_images") else None)
        self._labels = np.memmap(os.path.join(path, "labels.i64"), dtype=np.int64, mode="r", shape=(self.n_rows,))
        self._paths = None

    def _tensor(self, m, rows: Optional[slice] = None):
        a = m if rows is None else m[rows]
        with warnings.catch_warnings():       # the mapping is read-only by design; the tensors are only ever read
            warnings.simplefilter("ignore", UserWarning)
            t = torch.from_numpy(np.ascontiguousarray(a) if rows is not None else np.asarray(a))
        return t.view(torch.bfloat16) if self.dtype == "bf16" else t

    def caption(self, rows: Optional[slice] = None) -> torch.Tensor:
        return self._tensor(self._cap, rows)

    def image(self, rows: Optional[slice] = None) -> Optional[torch.Tensor]:
        return None if self._img is None else self._tensor(self._img, rows)

    def labels(self) -> torch.Tensor:
        return torch.from_numpy(np.asarray(self._labels))

    def paths(self):
        if self._paths is None:
            with open(os.path.join(self.path, "paths.txt")) as f:
                self._paths = f.read().split("\n")
        return self._paths

    def as_mined_dict(self) -> dict:
        return {"caption_features": self.caption(), "image_features": self.image(), "labels": self.labels(),
                "filepath": self.paths()}

    def to_device(self, device, rows: Optional[slice] = None, pinned_chunk_rows: int = 1 << 16, native: bool = True):
        """Rows -> HBM (a rank of a sharded run passes its own row range).

        ``native`` (default): the C-ABI loader ``swat_bank_load`` -- ``pread`` into two pinned staging buffers with the read
        of chunk i+1 overlapping the H2D copy of chunk i, or GPUDirect Storage with ``SWAT_GDS=1``; ``self.used_gds`` tells
        which.  ``native=False`` does the staged copy with torch tensors (kept as the loader's cross-check)."""
        rows = rows or slice(0, self.n_rows)
        if native:
            from . import _lib
            from .retrieval import get_context
            dev = torch.device(device)
            ctx = get_context(dev.index if dev.index is not None else torch.cuda.current_device())
            tdt = _TORCH[self.dtype]
            out_c, g1 = _lib.bank_load(ctx, os.path.join(self.path, "caption.bin"), tdt, rows.start, rows.stop, pinned_chunk_rows)
            out_i, g2 = (None, False)
            if self._img is not None:
                out_i, g2 = _lib.bank_load(ctx, os.path.join(self.path, "image.bin"), tdt, rows.start, rows.stop, pinned_chunk_rows)
            self.used_gds = bool(g1 or g2)
            return out_c, out_i
        n = rows.stop - rows.start
        tdt = _TORCH[self.dtype]
        out_c = torch.empty(n, DIM, dtype=tdt, device=device)
        out_i = torch.empty(n, DIM, dtype=tdt, device=device) if self._img is not None else None
        chunk = min(pinned_chunk_rows, max(n, 1))
        stages = [torch.empty(chunk, DIM, dtype=tdt, pin_memory=True) for _ in range(2)]
        busy = [None, None]
        i = 0
        with torch.cuda.device(device):
            for src, dst in ((self._cap, out_c), (self._img, out_i)):
                if src is None:
                    continue
                for s0 in range(0, n, chunk):
                    m = min(chunk, n - s0)
                    b = i & 1
                    if busy[b] is not None:
                        busy[b].synchronize()                  # the H2D copy that last read this buffer is done
                    stages[b][:m].copy_(self._tensor(src, slice(rows.start + s0, rows.start + s0 + m)))
                    dst[s0:s0 + m].copy_(stages[b][:m], non_blocking=True)
                    busy[b] = torch.cuda.Event()
                    busy[b].record()
                    i += 1
            torch.cuda.synchronize()
        return out_c, out_i
