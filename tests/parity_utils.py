"""Comparison helpers shared by the parity tests and tools/gpu_diag.py.

Comparator rules (BASELINE.json north_star): keypoint sets are
compared as SETS (the reference appends with atomicInc, cuSIFT_D.cu:513, so order
is unspecified); positions / scales within 1e-3 px, orientations within 1e-3 rad
(the field is in degrees: 0.0573 deg), descriptors within 1e-4 relative L2.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"

POS_TOL = 1e-3            # px (positions, scale)
ORI_TOL_DEG = 1e-3 * 180.0 / np.pi   # 1e-3 rad in degrees
DESC_TOL = 1e-4           # relative L2


def golden_frames():
    z = np.load(GOLDEN / "frames.npz")
    return z["gray1"].astype(np.float32), z["gray2"].astype(np.float32)


def preblur(img: np.ndarray) -> np.ndarray:
    """main.cpp:308-309: cv::GaussianBlur(img, img, Size(3,3), 0.5)."""
    import cv2

    return cv2.GaussianBlur(np.ascontiguousarray(img, np.float32), (3, 3), 0.5)


def kp_key(pts: np.ndarray) -> np.ndarray:
    return np.stack([pts["coords2D"][:, 0], pts["coords2D"][:, 1], pts["scale"]], 1).astype(np.float64)


def match_sets(a: np.ndarray, b: np.ndarray, tol: float = POS_TOL):
    """One-to-one association of two keypoint arrays by nearest (x, y, scale).

    Returns (ia, ib, only_a, only_b): matched index pairs (distance <= tol, mutual
    nearest) and the indices left unmatched on either side."""
    from scipy.spatial import cKDTree

    if len(a) == 0 or len(b) == 0:
        return np.zeros(0, int), np.zeros(0, int), np.arange(len(a)), np.arange(len(b))
    ka, kb = kp_key(a), kp_key(b)
    tb = cKDTree(kb)
    d, j = tb.query(ka)
    ta = cKDTree(ka)
    d2, i2 = ta.query(kb)
    ok = (d <= tol) & (i2[j] == np.arange(len(a)))
    ia = np.nonzero(ok)[0]
    ib = j[ok]
    only_a = np.setdiff1d(np.arange(len(a)), ia)
    only_b = np.setdiff1d(np.arange(len(b)), ib)
    return ia, ib, only_a, only_b


def ang_diff_deg(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    return np.abs(((a - b + 180.0) % 360.0) - 180.0)


def desc_rel_l2(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """||a-b|| / ||b|| per row for [n,128] descriptor arrays."""
    num = np.linalg.norm(a.astype(np.float64) - b.astype(np.float64), axis=1)
    den = np.maximum(np.linalg.norm(b.astype(np.float64), axis=1), 1e-30)
    return num / den


def compare_keypoints(ours: np.ndarray, ref: np.ndarray, tol: float = POS_TOL) -> dict:
    """Summary statistics of ours-vs-ref (both SIFT_DTYPE arrays)."""
    ia, ib, oa, ob = match_sets(ours, ref, tol)
    out = {
        "n_ours": int(len(ours)),
        "n_ref": int(len(ref)),
        "matched": int(len(ia)),
        "only_ours": int(len(oa)),
        "only_ref": int(len(ob)),
    }
    if len(ia):
        A, B = ours[ia], ref[ib]
        out["pos_exact"] = int(np.sum((A["coords2D"] == B["coords2D"]).all(1) & (A["scale"] == B["scale"])))
        out["pos_max"] = float(np.abs(A["coords2D"] - B["coords2D"]).max())
        out["scale_max"] = float(np.abs(A["scale"] - B["scale"]).max())
        out["sharp_max"] = float(np.abs(A["sharpness"] - B["sharpness"]).max())
        edge_rel = np.abs(A["edgeness"] - B["edgeness"]) / np.maximum(np.abs(B["edgeness"]), 1e-12)
        out["edge_rel_max"] = float(edge_rel.max())
        out["subs_equal"] = bool((A["subsampling"] == B["subsampling"]).all())
        do = ang_diff_deg(A["orientation"], B["orientation"])
        out["ori_median_deg"] = float(np.median(do))
        out["ori_within_tol"] = float(np.mean(do <= ORI_TOL_DEG))
        out["ori_max_deg"] = float(do.max())
        dr = desc_rel_l2(A["data"], B["data"])
        out["desc_median"] = float(np.median(dr))
        out["desc_within_tol"] = float(np.mean(dr <= DESC_TOL))
        out["desc_within_1e-3"] = float(np.mean(dr <= 1e-3))
        out["desc_max"] = float(dr.max())
        # descriptors of points whose orientation agrees (isolates the descriptor stage)
        good = do <= ORI_TOL_DEG
        if good.any():
            out["desc_within_tol_given_ori"] = float(np.mean(dr[good] <= DESC_TOL))
            out["desc_max_given_ori"] = float(dr[good].max())
    return out


def canonical_sort(pts: np.ndarray) -> np.ndarray:
    """Keypoints in (subsampling, x, y, scale) order — the list order itself is nondeterministic (atomics)."""
    order = np.lexsort((pts["scale"], pts["coords2D"][:, 1], pts["coords2D"][:, 0], pts["subsampling"]))
    return np.ascontiguousarray(pts[order])


def per_octave_counts(pts: np.ndarray) -> dict:
    sub = pts["subsampling"]
    return {str(float(k)): int((sub == k).sum()) for k in np.unique(sub)}
