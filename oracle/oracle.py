"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes/numpy front end of oracle/oracle.c
(the CPU restatement of the reference hot path) and of oracle/_ref/ref_driver (the
unmodified reference built for sm_100a).  Imported only by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.
The product package (cusift_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "_build" / "liboracle.so"
REF_DRIVER = HERE / "_ref" / "ref_driver"

# numpy view of SiftPoint (cuSIFT.h:10-30), 588 bytes
SIFT_DTYPE = np.dtype(
    [
        ("coords2D", "<f4", (2,)),
        ("scale", "<f4"),
        ("sharpness", "<f4"),
        ("edgeness", "<f4"),
        ("orientation", "<f4"),
        ("score", "<f4"),
        ("ambiguity", "<f4"),
        ("match", "<i4"),
        ("match_xpos", "<f4"),
        ("match_ypos", "<f4"),
        ("match_error", "<f4"),
        ("subsampling", "<f4"),
        ("empty", "<f4", (3,)),
        ("data", "<f4", (128,)),
        ("coords3D", "<f4", (3,)),
    ]
)
assert SIFT_DTYPE.itemsize == 588


def build(force: bool = False) -> Path:
    """Compile oracle.c (gcc) into oracle/_build/liboracle.so."""
    src = HERE / "oracle.c"
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-s", "-C", str(HERE), "oracle"])
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB_PATH))
        fp = C.POINTER(C.c_float)
        ip = C.POINTER(C.c_int)
        vp = C.c_void_p
        L.orc_sizeof_point.restype = C.c_int
        L.orc_scale_down.argtypes = [fp, C.c_int, C.c_int, C.c_int, fp, C.c_int]
        L.orc_scale_down_var.argtypes = [fp, C.c_int, C.c_int, C.c_int, fp, C.c_int, C.c_float]
        L.orc_next_init_blur.argtypes = [C.c_double]
        L.orc_next_init_blur.restype = C.c_double
        L.orc_laplace_weights.argtypes = [C.c_double, fp]
        L.orc_dog.argtypes = [fp, C.c_int, C.c_int, C.c_int, C.c_double, fp]
        L.orc_find_points.argtypes = [fp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, vp, C.c_int, ip]
        L.orc_find_points.restype = C.c_int
        L.orc_orientation.argtypes = [fp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float]
        L.orc_orientation.restype = C.c_float
        L.orc_descriptor.argtypes = [fp, C.c_int, C.c_int, C.c_int, vp, C.c_float]
        L.orc_rootsift.argtypes = [vp, C.c_int]
        L.orc_extract.argtypes = [fp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_float, C.c_float, C.c_float,
                                  C.c_int, vp, C.c_int, ip]
        L.orc_extract.restype = C.c_int
        L.orc_octave_stage.argtypes = [fp, C.c_int, C.c_int, C.c_int, C.c_double, fp, fp, ip, ip]
        L.orc_match.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int]
        L.orc_count_matches.argtypes = [vp, C.c_int, C.c_float, C.c_float]
        L.orc_count_matches.restype = C.c_int
        L.orc_compute_homographies.argtypes = [fp, ip, fp, C.c_int, C.c_int]
        L.orc_test_homographies.argtypes = [fp, fp, ip, C.c_int, C.c_int, C.c_float]
        L.orc_valid_points.argtypes = [vp, C.c_int, C.c_float, C.c_float, ip]
        L.orc_valid_points.restype = C.c_int
        L.orc_find_homography.argtypes = [vp, C.c_int, ip, C.c_int, C.c_float, fp]
        L.orc_find_homography.restype = C.c_int
        L.orc_improve_homography.argtypes = [vp, C.c_int, fp, C.c_int, C.c_float, C.c_float, C.c_float]
        L.orc_improve_homography.restype = C.c_int
        assert L.orc_sizeof_point() == 588
        _lib = L
    return _lib


def _f(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def scale_down(img: np.ndarray, variance: float = 0.5) -> np.ndarray:
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape
    out = np.zeros((h // 2, w // 2), np.float32)
    lib().orc_scale_down_var(_f(img), w, h, w, _f(out), w // 2, variance)
    return out


def laplace_weights(init_blur: float) -> np.ndarray:
    k = np.zeros((8, 9), np.float32)
    lib().orc_laplace_weights(float(init_blur), _f(k))
    return k


def init_blurs(n: int, init_blur: float = 0.0) -> list[float]:
    out = [float(init_blur)]
    for _ in range(n - 1):
        out.append(lib().orc_next_init_blur(out[-1]))
    return out


def dog(base: np.ndarray, init_blur: float) -> np.ndarray:
    base = np.ascontiguousarray(base, np.float32)
    h, w = base.shape
    out = np.zeros((7, h, w), np.float32)
    lib().orc_dog(_f(base), w, h, w, float(init_blur), _f(out))
    return out


def octave_stage(img: np.ndarray, octave: int, init_blur: float = 0.0):
    """(base, dog[7]) of octave `octave` for a full-resolution frame."""
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape
    ww, hh = w, h
    for _ in range(octave):
        ww, hh = ww // 2, hh // 2
    base = np.zeros((hh, ww), np.float32)
    d = np.zeros((7, hh, ww), np.float32)
    cw, ch = C.c_int(0), C.c_int(0)
    lib().orc_octave_stage(_f(img), w, h, octave, float(init_blur), _f(base), _f(d), C.byref(cw), C.byref(ch))
    assert (cw.value, ch.value) == (ww, hh)
    return base, d


def find_points(dogs: np.ndarray, peak_thresh: float, edge_thresh: float, subsampling: float, cap: int = 1 << 18):
    dogs = np.ascontiguousarray(dogs, np.float32)
    _, h, w = dogs.shape
    pts = np.zeros(cap, SIFT_DTYPE)
    mpb = C.c_int(0)
    n = lib().orc_find_points(_f(dogs), w, h, peak_thresh, edge_thresh, subsampling, pts.ctypes.data, cap, C.byref(mpb))
    return pts[: min(n, cap)], n, mpb.value


def extract(img: np.ndarray, num_octaves: int, init_blur: float, peak_thresh: float, edge_thresh: float = 10.0,
            lowest_scale: float = 0.0, rootsift: bool = False, max_pts: int = 1 << 17):
    """SiftData::Extract restated.  Returns (points[min(n,max_pts)], n_found, max candidates per block)."""
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape
    pts = np.zeros(max_pts, SIFT_DTYPE)
    mpb = C.c_int(0)
    n = lib().orc_extract(_f(img), w, h, num_octaves, float(init_blur), peak_thresh, edge_thresh, lowest_scale,
                          int(rootsift), pts.ctypes.data, max_pts, C.byref(mpb))
    return pts[: min(n, max_pts)].copy(), n, mpb.value


def rootsift(pts: np.ndarray) -> np.ndarray:
    pts = pts.copy()
    lib().orc_rootsift(pts.ctypes.data, len(pts))
    return pts


def match(s1: np.ndarray, s2: np.ndarray, distance: str = "l2") -> np.ndarray:
    """Device part of MatchSiftData: returns a copy of s1 with the 5 match fields filled."""
    s1 = np.ascontiguousarray(s1.copy())
    s2 = np.ascontiguousarray(s2)
    lib().orc_match(s1.ctypes.data, len(s1), s2.ctypes.data, len(s2), 1 if distance == "l2" else 0)
    return s1


def count_matches(s1: np.ndarray, score_threshold: float = 999.0, ambiguity_threshold: float = 1.0) -> int:
    s1 = np.ascontiguousarray(s1)
    return lib().orc_count_matches(s1.ctypes.data, len(s1), score_threshold, ambiguity_threshold)


def valid_points(pts: np.ndarray, min_score: float, max_ambiguity: float) -> np.ndarray:
    pts = np.ascontiguousarray(pts)
    idx = np.zeros(len(pts), np.int32)
    n = lib().orc_valid_points(pts.ctypes.data, len(pts), min_score, max_ambiguity, _i(idx))
    return idx[:n]


def find_homography(pts: np.ndarray, rand_pts: np.ndarray, thresh: float = 5.0):
    """rand_pts: int32 [4][numLoops] point indices.  Returns (H[9], best inlier count)."""
    pts = np.ascontiguousarray(pts)
    rand_pts = np.ascontiguousarray(rand_pts, np.int32)
    H = np.zeros(9, np.float32)
    cnt = lib().orc_find_homography(pts.ctypes.data, len(pts), _i(rand_pts), rand_pts.shape[1], thresh, _f(H))
    return H, cnt


def improve_homography(pts: np.ndarray, H: np.ndarray, loops: int, min_score: float, max_amb: float, thresh: float):
    pts = np.ascontiguousarray(pts.copy())
    H = np.ascontiguousarray(H, np.float32).copy()
    nfit = lib().orc_improve_homography(pts.ctypes.data, len(pts), _f(H), loops, min_score, max_amb, thresh)
    return H, nfit, pts


# ----------------------------------------------------------------------------
# Fixture readers (formats of the reference's test/data)
# ----------------------------------------------------------------------------
def read_vlfeat_sift(path) -> np.ndarray:
    """extras/debug.cpp:125-165: u32 n; f32 pts[n][4]; f32 desc[n][128] -> SiftPoint array."""
    raw = Path(path).read_bytes()
    n = int(np.frombuffer(raw, "<u4", 1)[0])
    pts = np.frombuffer(raw, "<f4", 4 * n, 4).reshape(n, 4)
    desc = np.frombuffer(raw, "<f4", 128 * n, 4 + 16 * n).reshape(n, 128)
    out = np.zeros(n, SIFT_DTYPE)
    out["coords2D"] = pts[:, :2]
    out["scale"] = pts[:, 2]
    out["orientation"] = pts[:, 3]
    out["data"] = desc
    return out


def read_match_indices(path):
    """extras/debug.cpp:167-181: u32 n; u32 i[n]; u32 j[n] (1-based)."""
    raw = Path(path).read_bytes()
    n = int(np.frombuffer(raw, "<u4", 1)[0])
    i = np.frombuffer(raw, "<u4", n, 4)
    j = np.frombuffer(raw, "<u4", n, 4 + 4 * n)
    return i.astype(np.int64), j.astype(np.int64)


def read_cusift_golden(path) -> np.ndarray:
    """test/detector.cpp:52-61: u32 n; f32 {x,y,scale,orientation deg}[n]."""
    raw = Path(path).read_bytes()
    n = int(np.frombuffer(raw, "<u4", 1)[0])
    return np.frombuffer(raw, "<f4", 4 * n, 4).reshape(n, 4).copy()


def read_sift_file(path) -> np.ndarray:
    raw = Path(path).read_bytes()
    n = int(np.frombuffer(raw, "<u4", 1)[0])
    return np.frombuffer(raw, SIFT_DTYPE, n, 4).copy()


def write_sift_file(path, pts: np.ndarray) -> None:
    with open(path, "wb") as fp:
        fp.write(np.uint32(len(pts)).tobytes())
        fp.write(np.ascontiguousarray(pts).tobytes())


# ----------------------------------------------------------------------------
# The unmodified reference on the GPU (oracle/_ref/ref_driver)
# ----------------------------------------------------------------------------
def ref_available() -> bool:
    return REF_DRIVER.exists() and os.access(REF_DRIVER, os.X_OK)


def _run_ref(args, timeout=600) -> str:
    res = subprocess.run([str(REF_DRIVER)] + [str(a) for a in args], capture_output=True, text=True, timeout=timeout)
    if res.returncode != 0:
        raise RuntimeError(f"ref_driver {args[0]} failed rc={res.returncode}: {res.stderr[-2000:]}")
    return res.stdout


def ref_extract(img: np.ndarray, workdir, num_octaves, init_blur, peak_thresh, edge_thresh=10.0, lowest_scale=0.0,
                rootsift=False, max_pts=1 << 17, safe=True, tag="ref") -> np.ndarray:
    workdir = Path(workdir)
    h, w = img.shape
    raw = workdir / f"{tag}_img.f32"
    np.ascontiguousarray(img, np.float32).tofile(raw)
    out = workdir / f"{tag}.sift"
    _run_ref(["extract_safe" if safe else "extract", raw, w, h, num_octaves, init_blur, peak_thresh, edge_thresh,
              lowest_scale, max_pts, int(rootsift), out])
    return read_sift_file(out)


def ref_stages(img: np.ndarray, workdir, num_octaves, init_blur=0.0, tag="ref"):
    """[(base, dog[7])] per octave from the reference's ScaleDown / LaplaceMulti."""
    workdir = Path(workdir)
    h, w = img.shape
    raw = workdir / f"{tag}_img.f32"
    np.ascontiguousarray(img, np.float32).tofile(raw)
    out = workdir / f"{tag}.stages"
    _run_ref(["stages", raw, w, h, num_octaves, init_blur, out])
    buf = out.read_bytes()
    n = int(np.frombuffer(buf, "<i4", 1)[0])
    dims = np.frombuffer(buf, "<i4", 2 * n, 4).reshape(n, 2)
    off = 4 + 8 * n
    res = []
    for ow, oh in dims:
        base = np.frombuffer(buf, "<f4", ow * oh, off).reshape(oh, ow)
        off += 4 * ow * oh
        d = np.frombuffer(buf, "<f4", 7 * ow * oh, off).reshape(7, oh, ow)
        off += 4 * 7 * ow * oh
        res.append((base, d))
    return res


def ref_match(s1: np.ndarray, s2: np.ndarray, workdir, distance="l2", score_thr=999.0, amb_thr=1.0, tag="ref"):
    workdir = Path(workdir)
    a, b, o = workdir / f"{tag}_a.sift", workdir / f"{tag}_b.sift", workdir / f"{tag}_m.sift"
    write_sift_file(a, s1)
    write_sift_file(b, s2)
    txt = _run_ref(["match", a, b, 1 if distance == "l2" else 0, score_thr, amb_thr, o])
    nmatch = json.loads(txt.strip().splitlines()[-1])["matches"]
    return read_sift_file(o), nmatch


def ref_homography(s1: np.ndarray, workdir, num_loops, min_score, max_amb, thresh, improve_loops=0,
                   improve_thresh=3.0, tag="ref"):
    workdir = Path(workdir)
    a, o = workdir / f"{tag}_h.sift", workdir / f"{tag}_h.txt"
    write_sift_file(a, s1)
    _run_ref(["homography", a, num_loops, min_score, max_amb, thresh, improve_loops, improve_thresh, o])
    lines = o.read_text().strip().splitlines()
    v = lines[0].split()
    res = {"num_matches": int(v[0]), "H": np.array(v[1:], np.float32)}
    if len(lines) > 1:
        v = lines[1].split()
        res["num_fit"] = int(v[0])
        res["H_improved"] = np.array(v[1:], np.float32)
    return res


# ---------------------------------------------------------------- rigid transform ---
def read_matlab_ransac(path):
    """ReadMATLABRANSAC (extras/debug.cpp:318-372): (coord [n,6] f32, indices [loops,3] 0-based, Rt [12] f32)."""
    b = Path(path).read_bytes()
    n, loops = np.frombuffer(b, np.uint32, 2)
    off = 8
    ci = np.frombuffer(b, np.float32, 3 * n, off).reshape(n, 3); off += 12 * n
    cj = np.frombuffer(b, np.float32, 3 * n, off).reshape(n, 3); off += 12 * n
    idx = np.frombuffer(b, np.int32, 3 * loops, off).reshape(loops, 3) - 1; off += 12 * loops
    Rt = np.frombuffer(b, np.float32, 12, off).copy()
    return np.ascontiguousarray(np.concatenate([ci, cj], 1)), np.ascontiguousarray(idx.astype(np.int32)), Rt


def rigid3d(coord: np.ndarray, idx) -> np.ndarray:
    """estimateRigidTransform3D (extras/rigidTransform.cu:15-205) in numpy: centroids and B in float32 like the
    reference, the smallest singular vector of B from LAPACK (the reference: float Numerical-Recipes dsvd)."""
    f = np.float32
    x = coord[idx, :3].astype(f); y = coord[idx, 3:].astype(f)
    n = f(len(idx))
    xc = np.zeros(3, f); yc = np.zeros(3, f)
    for i in range(len(idx)):                       # sequential float sums, as written
        xc = (xc + x[i]).astype(f); yc = (yc + y[i]).astype(f)
    xc = (xc / n).astype(f); yc = (yc / n).astype(f)
    x = (x - xc).astype(f); y = (y - yc).astype(f)
    B = np.zeros((4, 4), f)
    for i in range(len(idx)):
        d = (y[i] - x[i]).astype(f); s = (y[i] + x[i]).astype(f)
        A = np.array([[0, d[0], d[1], d[2]],
                      [-d[0], 0, -s[2], s[1]],
                      [-d[1], s[2], 0, -s[0]],
                      [-d[2], -s[1], s[0], 0]], f)
        B = (B + (A @ A.T).astype(f)).astype(f)
    _, S, Vt = np.linalg.svd(B.astype(np.float64))
    Q = Vt[int(np.argmin(S))].astype(f)
    q0, q1, q2, q3 = (np.float64(v) for v in Q)
    R = np.array([[1 - 2 * (q2 * q2 + q3 * q3), 2 * (q1 * q2 - q0 * q3), 2 * (q1 * q3 + q0 * q2)],
                  [2 * (q1 * q2 + q0 * q3), 1 - 2 * (q1 * q1 + q3 * q3), 2 * (q2 * q3 - q0 * q1)],
                  [2 * (q1 * q3 - q0 * q2), 2 * (q2 * q3 + q0 * q1), 1 - 2 * (q1 * q1 + q2 * q2)]]).astype(f)
    t = ((R @ (-yc)).astype(f) + xc).astype(f)
    return np.concatenate([R, t[:, None]], 1).astype(f).ravel()


def rigid2d(coord: np.ndarray, a: int, b: int) -> np.ndarray:
    """estimateRigidTransform2D (extras/rigidTransform.cu:214-290)."""
    f = np.float32
    A, Bp = coord[a].astype(f), coord[b].astype(f)
    dxw, dzw = f(A[0] - Bp[0]), f(A[2] - Bp[2])
    lw = f(np.sqrt(f(dxw * dxw + dzw * dzw)))
    dxc, dzc = f(A[3] - Bp[3]), f(A[5] - Bp[5])
    lc = f(np.sqrt(f(dxc * dxc + dzc * dzc)))
    dxwn, dzwn, dxcn, dzcn = f(dxw / lw), f(dzw / lw), f(dxc / lc), f(dzc / lc)
    c = f(dxwn * dxcn + dzwn * dzcn); s = f(dzwn * dxcn - dxwn * dzcn)
    sxw, szw, sxc, szc = f(A[0] + Bp[0]), f(A[2] + Bp[2]), f(A[3] + Bp[3]), f(A[5] + Bp[5])
    return np.array([c, 0, -s, (sxw - c * sxc + s * szc) / 2, 0, 1, 0, 0, s, 0, c, (szw - s * sxc - c * szc) / 2], f)


def rigid_inliers(coord: np.ndarray, Rt: np.ndarray, thresh2: float) -> np.ndarray:
    """testRigidTransform (extras/rigidTransform.cu:292-328): boolean mask."""
    R = Rt.reshape(3, 4).astype(np.float32)
    p = (coord[:, 3:].astype(np.float32) @ R[:, :3].T + R[:, 3]).astype(np.float32)
    err = ((p - coord[:, :3]) ** 2).sum(1).astype(np.float32)
    return err < np.float32(thresh2)


def rigid_transform(coord: np.ndarray, indices: np.ndarray, thresh2: float, type3d: bool = True):
    """EstimateRigidTransformH (extras/rigidTransform.cu:387-520) with given indices: (Rt[12], numInliers, mask)."""
    best, best_cnt, hyp = -1, -1, []
    for l in range(len(indices)):
        Rt = rigid3d(coord, indices[l]) if type3d else rigid2d(coord, int(indices[l][0]), int(indices[l][1]))
        hyp.append(Rt)
        c = int(rigid_inliers(coord, Rt, thresh2).sum())
        if c >= best_cnt:                           # `>=`: the last maximum wins
            best_cnt, best = c, l
    mask = rigid_inliers(coord, hyp[best], thresh2)
    Rt = hyp[best]
    if type3d and best_cnt >= 3:
        Rt = rigid3d(coord, np.nonzero(mask)[0])
    return Rt, best_cnt, mask


def ref_rigid(coord: np.ndarray, indices: np.ndarray, thresh2: float, type3d: bool, workdir, tag="ref"):
    """The unmodified reference's EstimateRigidTransformH through oracle/_ref/ref_driver (needs a GPU)."""
    workdir = Path(workdir)
    c, i, o = workdir / f"{tag}_rt_coord.f32", workdir / f"{tag}_rt_idx.i32", workdir / f"{tag}_rt.txt"
    np.ascontiguousarray(coord, np.float32).tofile(c)
    np.ascontiguousarray(indices, np.int32).tofile(i)
    _run_ref(["rigid", c, len(coord), i, len(indices), repr(float(thresh2)), int(type3d), o])
    v = o.read_text().split()
    return np.array(v[1:13], np.float32), int(v[0]), np.array(v[13:], np.int32).astype(bool)
