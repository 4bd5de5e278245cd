"""Helpers to load the committed golden fixtures (made by oracle/gen_golden.py from the reference)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def bf16_bits_to_f32(bits: np.ndarray) -> np.ndarray:
    return (bits.astype(np.uint32) << 16).view(np.float32)


def load_bank_case(name):
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    meta = json.load(open(os.path.join(GOLDEN, f"{name}.json")))
    if "cap_bf16" in z:
        cap = bf16_bits_to_f32(z["cap_bf16"]); img = bf16_bits_to_f32(z["img_bf16"]); q = bf16_bits_to_f32(z["q_bf16"])
    else:
        cap, img, q = z["cap_f32"], z["img_f32"], z["q_f32"]
    return z, meta, cap, img, q


def make_paths(labels, class_ids, root="/scratch/retrieved/synthetic"):
    paths, cmap = [], {}
    for i, l in enumerate(labels.tolist()):
        cid = str(int(class_ids[int(l)]))
        paths.append(f"{root}/{cid}/{i}.jpg")
        cmap.setdefault(cid, {})[str(i)] = f"synthetic caption {i}"
    return paths, cmap


def assert_walk_equal(got_rows, ref_rows, score_of_row, tol, boundary_tol=None, what=""):
    """Compare two walk-ordered row lists of one class under the north star's parity rule.

    * same length;
    * position-wise the scores agree within ``tol`` (so any reordering is among near-ties);
    * rows present on one side only sit at the k-th boundary: their score is within
      ``boundary_tol`` of the last accepted score.

    ``score_of_row`` maps row ids to the ORACLE's fp32 score.  With ``tol == 0`` this is exact list
    equality.  A non-zero ``tol`` is needed even between two CPU BLAS libraries: on fp32 inputs the
    reference's MKL GEMV returns scores 1 ulp apart for bit-identical rows (row position inside
    MKL's blocking), so its order among exact duplicates is not index order (seen in the bank_f32
    fixture, rows 214/226).
    """
    got = np.asarray(got_rows).tolist(); ref = np.asarray(ref_rows).tolist()
    assert len(got) == len(ref), f"{what}: count {len(got)} != {len(ref)}"
    if got == ref:
        return 0
    assert tol > 0, f"{what}: rows differ and tol == 0: {got[:10]} vs {ref[:10]}"
    boundary_tol = tol if boundary_tol is None else boundary_tol
    sg = np.asarray([score_of_row(r) for r in got], dtype=np.float64)
    sr = np.asarray([score_of_row(r) for r in ref], dtype=np.float64)
    assert np.all(np.abs(sg - sr) <= tol), f"{what}: position-wise score gap {np.abs(sg - sr).max()} > {tol}"
    only_g = set(got) - set(ref); only_r = set(ref) - set(got)
    last = min(sg[-1], sr[-1])
    for r in only_g | only_r:
        assert abs(score_of_row(r) - last) <= boundary_tol, f"{what}: row {r} differs away from the k-th boundary"
    return sum(1 for a, b in zip(got, ref) if a != b)
