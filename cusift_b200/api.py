"""Thin Python harness over the C ABI (tests / bench plumbing, not the product).

The product's host API is the C++ one in include/cusift/ (cuImage, SiftData,
ExtractSift, MatchSiftData, FindHomography); this module only lets pytest and
bench.py drive the same C entry points with numpy buffers.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import CsbParams, lib

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

# csb_compact_point (include/cusift_b200.h): opt-in 288-byte result record, descriptor as fp16
COMPACT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("scale", "<f4"), ("orientation", "<f4"), ("sharpness", "<f4"),
                          ("edgeness", "<f4"), ("subsampling", "<f4"), ("reserved", "<f4"), ("data", "<f2", (128,))])
assert COMPACT_DTYPE.itemsize == 288


class CsbError(RuntimeError):
    pass


def align_up(a: int, b: int) -> int:
    return (a + b - 1) // b * b


def make_params(num_octaves=5, init_blur=0.0, peak_thresh=0.1, edge_thresh=10.0, lowest_scale=0.0, subsampling=1.0,
                rootsift=False) -> CsbParams:
    return CsbParams(int(num_octaves), float(init_blur), float(peak_thresh), float(edge_thresh), float(lowest_scale),
                     float(subsampling), int(bool(rootsift)))


class PinnedArray:
    """Page-locked, device-mapped host array (csb_host_alloc) viewed through numpy."""

    def __init__(self, count: int, dtype=SIFT_DTYPE):
        self.dtype = np.dtype(dtype)
        self.count = int(count)
        self.nbytes = max(1, self.count * self.dtype.itemsize)
        p = C.c_void_p()
        rc = lib().csb_host_alloc(C.byref(p), self.nbytes)
        if rc:
            raise CsbError(f"csb_host_alloc failed: {rc}")
        self.ptr = p.value
        buf = (C.c_char * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=self.count)

    def free(self):
        if self.ptr:
            self.array = None
            lib().csb_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """csb_ctx: one per GPU."""

    def __init__(self, device: int = 0, slots: int = 0):
        self._L = lib()
        h = C.c_void_p()
        rc = self._L.csb_ctx_create(int(device), int(slots), C.byref(h))
        if rc or not h.value:
            raise CsbError(f"csb_ctx_create(device={device}) failed with status {rc} (a CUDA GPU is required)")
        self.h = h
        self._dev_allocs = set()

    # ---- plumbing -------------------------------------------------------
    def _check(self, rc: int, what: str):
        if rc:
            msg = self._L.csb_last_error(self.h)
            raise CsbError(f"{what} failed with status {rc}: {msg.decode() if msg else ''}")

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            for p in list(self._dev_allocs):
                self.free(p)
            self._L.csb_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def num_slots(self) -> int:
        return self._L.csb_ctx_num_slots(self.h)

    def alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._check(self._L.csb_device_alloc(self.h, C.byref(p), max(1, int(nbytes))), "csb_device_alloc")
        self._dev_allocs.add(p.value)
        return p.value

    def free(self, ptr: int):
        if ptr in self._dev_allocs:
            self._dev_allocs.discard(ptr)
            self._check(self._L.csb_device_free(self.h, ptr), "csb_device_free")

    def h2d(self, dptr: int, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        self._check(self._L.csb_memcpy_h2d(self.h, dptr, arr.ctypes.data, arr.nbytes), "csb_memcpy_h2d")

    def d2h(self, arr: np.ndarray, dptr: int):
        assert arr.flags["C_CONTIGUOUS"]
        self._check(self._L.csb_memcpy_d2h(self.h, arr.ctypes.data, dptr, arr.nbytes), "csb_memcpy_d2h")

    def upload_image(self, img: np.ndarray):
        """Dense float32 frame -> pitched device image (cuImage convention). Returns (dptr, pitch_floats)."""
        img = np.ascontiguousarray(img, np.float32)
        h, w = img.shape
        pitch = align_up(w, 128)
        d = self.alloc(4 * pitch * h)
        self._check(self._L.csb_upload_image(self.h, d, pitch, img.ctypes.data, w, h), "csb_upload_image")
        return d, pitch

    def download_image(self, dptr: int, pitch: int, w: int, h: int) -> np.ndarray:
        out = np.empty((h, w), np.float32)
        self._check(self._L.csb_download_image(self.h, out.ctypes.data, dptr, pitch, w, h), "csb_download_image")
        return out

    def upload_sift(self, pts: np.ndarray) -> int:
        pts = np.ascontiguousarray(pts, SIFT_DTYPE)
        d = self.alloc(pts.nbytes)
        if len(pts):
            self.h2d(d, pts)
        return d

    def download_sift(self, dptr: int, n: int) -> np.ndarray:
        out = np.zeros(n, SIFT_DTYPE)
        if n:
            self.d2h(out, dptr)
        return out

    # ---- the hot path ----------------------------------------------------
    def extract(self, img: np.ndarray, params: CsbParams, max_pts: int = 32768, from_host: bool = False,
                keep_device: bool = False):
        """ExtractSift (device-resident frame) or SiftData::Extract (from_host) on one frame.

        Returns the SiftPoint array (numpy, length num_pts); with keep_device also the
        device pointer of the SiftPoint buffer (caller frees via ctx.free)."""
        img = np.ascontiguousarray(img, np.float32)
        h, w = img.shape
        d_sift = self.alloc(588 * max_pts)
        host = np.zeros(max_pts, SIFT_DTYPE)
        n = C.c_int(0)
        try:
            if from_host:
                self._check(self._L.csb_extract_host(self.h, img.ctypes.data, w, h, C.byref(params), d_sift, max_pts,
                                                     host.ctypes.data, C.byref(n)), "csb_extract_host")
            else:
                d_img, pitch = self.upload_image(img)
                try:
                    self._check(self._L.csb_extract(self.h, d_img, w, h, pitch, C.byref(params), d_sift, max_pts,
                                                    host.ctypes.data, C.byref(n)), "csb_extract")
                finally:
                    self.free(d_img)
        except Exception:
            self.free(d_sift)
            raise
        pts = host[: n.value].copy()
        if keep_device:
            return pts, d_sift
        self.free(d_sift)
        return pts

    def extract_batch(self, d_imgs, w: int, h: int, pitch: int, params: CsbParams, d_sifts, h_sifts, max_pts: int,
                      on_host: bool = False) -> np.ndarray:
        """csb_extract_batch over lists of raw pointers (ints). Returns the per-frame counts."""
        n = len(d_imgs)
        imgs_arr = (C.c_void_p * n)(*d_imgs)
        ds_arr = (C.c_void_p * n)(*d_sifts)
        hs_arr = (C.c_void_p * n)(*h_sifts) if h_sifts is not None else None
        counts = np.zeros(n, np.int32)
        self._check(self._L.csb_extract_batch(self.h, n, imgs_arr, int(on_host), w, h, pitch, C.byref(params), ds_arr,
                                              hs_arr, max_pts, counts.ctypes.data_as(C.POINTER(C.c_int))),
                    "csb_extract_batch")
        return counts

    def extract_batch_u8(self, h_imgs, w: int, h: int, stride: int, preblur: bool, params: CsbParams, d_sifts, h_sifts,
                         max_pts: int) -> np.ndarray:
        """csb_extract_batch_u8 over lists of raw host pointers to 8-bit frames. Returns the per-frame counts."""
        n = len(h_imgs)
        imgs_arr = (C.c_void_p * n)(*h_imgs)
        ds_arr = (C.c_void_p * n)(*d_sifts)
        hs_arr = (C.c_void_p * n)(*h_sifts) if h_sifts is not None else None
        counts = np.zeros(n, np.int32)
        self._check(self._L.csb_extract_batch_u8(self.h, n, imgs_arr, w, h, stride, int(preblur), C.byref(params), ds_arr,
                                                 hs_arr, max_pts, counts.ctypes.data_as(C.POINTER(C.c_int))),
                    "csb_extract_batch_u8")
        return counts

    def extract_batch_compact(self, imgs, w: int, h: int, pitch: int, params: CsbParams, d_sifts, h_compact, max_pts: int,
                              source: str = "device") -> np.ndarray:
        """csb_extract_batch_compact: like extract_batch, results delivered as COMPACT_DTYPE records.
        source: "device" (pitched fp32 device images), "host" (dense fp32 host frames) or "host_u8" (dense 8-bit host
        frames; pitch = row stride in bytes)."""
        n = len(imgs)
        mode = {"device": 0, "host": 1, "host_u8": 2}[source]
        imgs_arr = (C.c_void_p * n)(*imgs)
        ds_arr = (C.c_void_p * n)(*d_sifts)
        hs_arr = (C.c_void_p * n)(*h_compact) if h_compact is not None else None
        counts = np.zeros(n, np.int32)
        self._check(self._L.csb_extract_batch_compact(self.h, n, imgs_arr, mode, w, h, pitch, C.byref(params), ds_arr, hs_arr,
                                                      max_pts, counts.ctypes.data_as(C.POINTER(C.c_int))),
                    "csb_extract_batch_compact")
        return counts

    def rigid_transform(self, coord: np.ndarray, indices, num_loops: int, thresh2: float, type3d: bool = True, seed: int = 1):
        """csb_rigid_transform. indices: [num_loops, 3] int32 or None (drawn on the device).
        Returns (Rt[12], num_inliers, mask[num_pts] bool)."""
        coord = np.ascontiguousarray(coord, np.float32)
        n = len(coord)
        idx = None if indices is None else np.ascontiguousarray(indices, np.int32)
        Rt = np.zeros(12, np.float32)
        ninl = C.c_int(0)
        mask = np.zeros(max(n, 1), np.int8)
        self._check(self._L.csb_rigid_transform(
            self.h, coord.ctypes.data_as(C.POINTER(C.c_float)), n, int(type3d),
            None if idx is None else idx.ctypes.data_as(C.POINTER(C.c_int)), num_loops, thresh2, seed,
            Rt.ctypes.data_as(C.POINTER(C.c_float)), C.byref(ninl), mask.ctypes.data_as(C.c_char_p)), "csb_rigid_transform")
        return Rt, ninl.value, mask[:n].astype(bool)

    def ingest_u8(self, img: np.ndarray, preblur: bool) -> np.ndarray:
        """csb_ingest_u8 of a host uint8 frame; returns the fp32 device image downloaded again."""
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        pitch = align_up(w, 128)
        d_dst = self.alloc(4 * pitch * h)
        try:
            self._check(self._L.csb_ingest_u8(self.h, img.ctypes.data, 1, w, h, w, int(preblur), d_dst, pitch), "csb_ingest_u8")
            return self.download_image(d_dst, pitch, w, h)
        finally:
            self.free(d_dst)

    def scale_down(self, img: np.ndarray, variance: float = 0.5) -> np.ndarray:
        img = np.ascontiguousarray(img, np.float32)
        h, w = img.shape
        d_src, sp = self.upload_image(img)
        dp = align_up(w // 2, 128)
        d_dst = self.alloc(4 * dp * (h // 2))
        try:
            self._check(self._L.csb_scale_down_var(self.h, d_src, w, h, sp, d_dst, dp, variance), "csb_scale_down_var")
            return self.download_image(d_dst, dp, w // 2, h // 2)
        finally:
            self.free(d_src)
            self.free(d_dst)

    def rootsift(self, pts: np.ndarray) -> np.ndarray:
        d = self.upload_sift(pts)
        try:
            self._check(self._L.csb_rootsift(self.h, d, len(pts)), "csb_rootsift")
            return self.download_sift(d, len(pts))
        finally:
            self.free(d)

    def debug_octave(self, octave: int, want_base: bool = True):
        """(base or None, dog[7]) of `octave` from the last frame extracted on slot 0."""
        w, h = C.c_int(0), C.c_int(0)
        self._check(self._L.csb_debug_octave(self.h, octave, None, None, C.byref(w), C.byref(h)), "csb_debug_octave")
        dog = np.zeros((7, h.value, w.value), np.float32)
        base = np.zeros((h.value, w.value), np.float32) if (want_base and octave > 0) else None
        self._check(self._L.csb_debug_octave(self.h, octave,
                                             base.ctypes.data_as(C.POINTER(C.c_float)) if base is not None else None,
                                             dog.ctypes.data_as(C.POINTER(C.c_float)), C.byref(w), C.byref(h)),
                    "csb_debug_octave")
        return base, dog

    def match(self, s1: np.ndarray, s2: np.ndarray, distance: str = "l2") -> np.ndarray:
        """Device part of MatchSiftData; returns s1 with the five match fields filled."""
        d1, d2 = self.upload_sift(s1), self.upload_sift(s2)
        try:
            host = np.ascontiguousarray(s1.copy(), SIFT_DTYPE)
            self._check(self._L.csb_match(self.h, d1, len(s1), d2, len(s2), 1 if distance == "l2" else 0,
                                          host.ctypes.data), "csb_match")
            dev = self.download_sift(d1, len(s1))
            for f in ("score", "ambiguity", "match", "match_xpos", "match_ypos"):
                assert np.array_equal(host[f], dev[f], equal_nan=True), f"host/device copies of {f} differ"
            return dev
        finally:
            self.free(d1)
            self.free(d2)

    def find_homography(self, pts: np.ndarray, rand_pts: np.ndarray, thresh: float = 5.0):
        rand_pts = np.ascontiguousarray(rand_pts, np.int32)
        assert rand_pts.ndim == 2 and rand_pts.shape[0] == 4
        d = self.upload_sift(pts)
        try:
            H = np.zeros(9, np.float32)
            n = C.c_int(0)
            self._check(self._L.csb_find_homography(self.h, d, len(pts), rand_pts.ctypes.data_as(C.POINTER(C.c_int)),
                                                    rand_pts.shape[1], thresh, H.ctypes.data_as(C.POINTER(C.c_float)),
                                                    C.byref(n)), "csb_find_homography")
            return H, n.value
        finally:
            self.free(d)

    def improve_homography(self, pts: np.ndarray, H: np.ndarray, loops: int = 5, min_score: float = 0.0,
                           max_ambiguity: float = 0.80, thresh: float = 3.0):
        """csb_improve_homography on a device copy of `pts`: returns (H [9], num_fit, pts with match_error)."""
        d = self.upload_sift(pts)
        try:
            Hc = np.ascontiguousarray(H, np.float32).copy()
            n = C.c_int(0)
            out = np.ascontiguousarray(pts.copy())
            self._check(self._L.csb_improve_homography(self.h, d, len(pts), Hc.ctypes.data_as(C.POINTER(C.c_float)), loops,
                                                       min_score, max_ambiguity, thresh, C.byref(n), out.ctypes.data),
                        "csb_improve_homography")
            return Hc, n.value, out
        finally:
            self.free(d)

    def allpairs(self, d_sifts, counts, pairs, distance="l2", num_loops=1024, min_score=0.0, max_ambiguity=0.80,
                 thresh=5.0, seed=1, pair_ids=None, improve_loops=0, improve_thresh=3.0):
        """csb_allpairs_match_ransac over device SiftPoint arrays (raw pointers).  pairs: [(i, j)].
        Returns (H [n_pairs, 9], inliers [n_pairs], n_valid [n_pairs]); with improve_loops > 0 also
        (H_improved [n_pairs, 9], num_fit [n_pairs]) from the ImproveHomography step appended to every pair."""
        n_sets, n_pairs = len(d_sifts), len(pairs)
        ptrs = (C.c_void_p * n_sets)(*d_sifts)
        cnts = np.ascontiguousarray(counts, np.int32)
        pi = np.ascontiguousarray([p[0] for p in pairs], np.int32)
        pj = np.ascontiguousarray([p[1] for p in pairs], np.int32)
        ids = None if pair_ids is None else np.ascontiguousarray(pair_ids, np.uint32)
        H = np.zeros((max(n_pairs, 1), 9), np.float32)
        inl = np.zeros(max(n_pairs, 1), np.int32)
        nv = np.zeros(max(n_pairs, 1), np.int32)
        ip = C.POINTER(C.c_int)
        if improve_loops > 0:
            H2 = np.zeros((max(n_pairs, 1), 9), np.float32)
            nf = np.zeros(max(n_pairs, 1), np.int32)
            self._check(self._L.csb_allpairs_match_ransac_improve(
                self.h, n_sets, ptrs, cnts.ctypes.data_as(ip), n_pairs, pi.ctypes.data_as(ip), pj.ctypes.data_as(ip),
                None if ids is None else ids.ctypes.data_as(C.POINTER(C.c_uint)), 1 if distance == "l2" else 0, num_loops,
                min_score, max_ambiguity, thresh, seed, improve_loops, improve_thresh, H.ctypes.data_as(C.POINTER(C.c_float)),
                inl.ctypes.data_as(ip), nv.ctypes.data_as(ip), H2.ctypes.data_as(C.POINTER(C.c_float)), nf.ctypes.data_as(ip)),
                "csb_allpairs_match_ransac_improve")
            return H[:n_pairs], inl[:n_pairs], nv[:n_pairs], H2[:n_pairs], nf[:n_pairs]
        self._check(self._L.csb_allpairs_match_ransac(
            self.h, n_sets, ptrs, cnts.ctypes.data_as(ip), n_pairs, pi.ctypes.data_as(ip), pj.ctypes.data_as(ip),
            None if ids is None else ids.ctypes.data_as(C.POINTER(C.c_uint)), 1 if distance == "l2" else 0, num_loops,
            min_score, max_ambiguity, thresh, seed, H.ctypes.data_as(C.POINTER(C.c_float)), inl.ctypes.data_as(ip),
            nv.ctypes.data_as(ip)), "csb_allpairs_match_ransac")
        return H[:n_pairs], inl[:n_pairs], nv[:n_pairs]

    # ---- multi-GPU all-pairs (one process per GPU) -------------------------
    def nccl_comm(self, rank: int, world: int, broadcast_bytes):
        """Creates the library's NCCL communicator.  broadcast_bytes(buf: bytes | None) -> bytes must return rank 0's
        128-byte id on every rank (e.g. via torch.distributed.broadcast_object_list)."""
        idbuf = C.create_string_buffer(128)
        if rank == 0:
            self._check(self._L.csb_nccl_unique_id(idbuf), "csb_nccl_unique_id")
        raw = broadcast_bytes(idbuf.raw if rank == 0 else None)
        idbuf = C.create_string_buffer(raw, 128)
        comm = C.c_void_p()
        self._check(self._L.csb_nccl_comm_create(self.h, rank, world, idbuf, C.byref(comm)), "csb_nccl_comm_create")
        return comm

    def nccl_comm_destroy(self, comm):
        self._L.csb_nccl_comm_destroy(comm)

    def allpairs_distributed(self, comm, rank: int, world: int, d_local_sifts, local_counts, cap: int, distance="l2",
                             num_loops=1024, min_score=0.0, max_ambiguity=0.80, thresh=5.0, seed=1, improve_loops=0,
                             improve_thresh=3.0):
        """csb_allpairs_distributed: returns dict(H, inliers, n_valid[, H_improved, num_fit], timings_ms) for ALL pairs."""
        spr = len(d_local_sifts)
        n_sets = spr * world
        n_pairs = n_sets * (n_sets - 1) // 2
        ptrs = (C.c_void_p * spr)(*d_local_sifts)
        cnts = np.ascontiguousarray(local_counts, np.int32)
        H = np.zeros((max(n_pairs, 1), 9), np.float32)
        inl = np.zeros(max(n_pairs, 1), np.int32)
        nv = np.zeros(max(n_pairs, 1), np.int32)
        H2 = np.zeros((max(n_pairs, 1), 9), np.float32)
        nf = np.zeros(max(n_pairs, 1), np.int32)
        tm = np.zeros(4, np.float64)
        ip, fp = C.POINTER(C.c_int), C.POINTER(C.c_float)
        self._check(self._L.csb_allpairs_distributed(
            self.h, comm, rank, world, spr, ptrs, cnts.ctypes.data_as(ip), cap, 1 if distance == "l2" else 0, num_loops,
            min_score, max_ambiguity, thresh, seed, improve_loops, improve_thresh, H.ctypes.data_as(fp), inl.ctypes.data_as(ip),
            nv.ctypes.data_as(ip), H2.ctypes.data_as(fp) if improve_loops > 0 else None,
            nf.ctypes.data_as(ip) if improve_loops > 0 else None, tm.ctypes.data_as(C.POINTER(C.c_double))),
            "csb_allpairs_distributed")
        out = {"H": H[:n_pairs], "inliers": inl[:n_pairs], "n_valid": nv[:n_pairs], "timings_ms": tm}
        if improve_loops > 0:
            out["H_improved"], out["num_fit"] = H2[:n_pairs], nf[:n_pairs]
        return out

    def sample_points(self, valid: np.ndarray, num_loops: int, seed: int, pair_id: int) -> np.ndarray:
        """Host restatement of the device sample generator (k_ransac_prep): int32 [4][num_loops]."""
        nv = len(valid)
        rp = np.zeros((4, num_loops), np.int32)
        if nv < 8:
            return rp
        for l in range(num_loops):
            picks = []
            for k in range(4):
                attempt = 0
                while True:
                    c = self._L.csb_sample_hash(seed, pair_id, l, k, attempt) % nv
                    attempt += 1
                    if c not in picks:
                        picks.append(c)
                        break
            rp[:, l] = valid[picks]
        return rp

    # ---- measurement ------------------------------------------------------
    def profile(self, on: bool):
        self._check(self._L.csb_profile_enable(self.h, int(on)), "csb_profile_enable")

    def profile_reset(self):
        self._check(self._L.csb_profile_reset(self.h), "csb_profile_reset")

    def profile_table(self) -> dict:
        out = {}
        for i in range(self._L.csb_profile_count(self.h)):
            name, ms, cnt = C.c_char_p(), C.c_double(), C.c_longlong()
            self._L.csb_profile_get(self.h, i, C.byref(name), C.byref(ms), C.byref(cnt))
            out[name.value.decode()] = {"total_ms": ms.value, "launches": cnt.value}
        return out

    def launch_count(self) -> int:
        return int(self._L.csb_launch_count(self.h))
