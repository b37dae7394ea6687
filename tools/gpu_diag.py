#!/usr/bin/env python
"""GPU diagnostics: runs every stage of the product library against the CPU oracle
and the reference build (oracle/_ref) and writes a detailed report to
gpurun_out/diag.json.  Development aid (each stage is isolated so one failure
does not hide the others); the graded checks live in tests/.
"""
from __future__ import annotations

import json
import os
import sys
import time
import traceback
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import cusift_b200 as csb  # noqa: E402
from oracle import oracle as O  # noqa: E402
import parity_utils as PU  # noqa: E402

OUT = ROOT / "gpurun_out"
OUT.mkdir(exist_ok=True)
WORK = OUT / "work"
WORK.mkdir(exist_ok=True)
REPORT = {}
QUICK = "--quick" in sys.argv          # small stages only (used under compute-sanitizer)
SLOW = {"synth_1080p_ours_vs_ref", "synth_1080p_ref_plain", "synth_1080p_vs_oracle", "timing_1080p",
        "timing_ref_1080p", "ref_extract_plain_gray1", "ref_extract_safe_gray1_vs_ours", "ref_twice_gray1",
        "ref_stages_gray1_vs_oracle", "scale_down_vs_oracle"}


def stage(name):
    def deco(fn):
        if QUICK and name in SLOW:
            return fn
        t0 = time.time()
        try:
            REPORT[name] = fn()
        except Exception as e:  # noqa: BLE001
            REPORT[name] = {"error": repr(e), "trace": traceback.format_exc()[-3000:]}
        REPORT[name + "__sec"] = round(time.time() - t0, 2)
        print(f"[{name}] {json.dumps(REPORT[name], default=str)[:1500]}", flush=True)
        (OUT / ("diag_quick.json" if QUICK else "diag.json")).write_text(json.dumps(REPORT, indent=1, default=str))
        return fn
    return deco


def arr_diff(a, b):
    a = np.asarray(a); b = np.asarray(b)
    neq = a != b
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    out = {"shape": list(a.shape), "n_diff": int(neq.sum()), "max_abs": float(d.max()) if d.size else 0.0}
    if neq.any():
        idx = np.argwhere(neq)
        out["first_diff"] = idx[0].tolist()
        out["diff_rows"] = [int(idx[:, -2].min()), int(idx[:, -2].max())]
        out["diff_cols"] = [int(idx[:, -1].min()), int(idx[:, -1].max())]
    return out


g1, g2 = PU.golden_frames()
ctx = csb.Context(0, 4)


@stage("scale_down_vs_oracle")
def _():
    res = {}
    for name, img in (("gray1", g1), ("odd_135x241", csb.synth(241, 135, 7)), ("synth_1080p", csb.synth(1920, 1080, 1000))):
        res[name] = arr_diff(ctx.scale_down(img), O.scale_down(img))
    return res


def dog_check(img, n_oct, init_blur, thresh, label):
    res = {}
    p = csb.make_params(n_oct, init_blur, thresh, 10.0, 0.0)
    pts = ctx.extract(img, p, max_pts=1 << 17)
    res["n_pts"] = len(pts)
    blurs = O.init_blurs(n_oct, init_blur)
    for o in range(n_oct):
        base, dog = ctx.debug_octave(o)
        ob, od = O.octave_stage(img, o, init_blur)
        if base is not None:
            res[f"o{o}_base"] = arr_diff(base, ob)
        res[f"o{o}_dog"] = arr_diff(dog, od)
    return res, pts


@stage("dog_gray1_fused_vs_oracle")
def _():
    r, _pts = dog_check(g1, 6, 0.0, 0.1, "gray1")
    return r


@stage("dog_synth_small_vs_oracle")
def _():
    r, _pts = dog_check(csb.synth(500, 300, 11), 4, 0.0, 1.0, "synth500")
    return r


@stage("dog_gray1_nofuse_vs_oracle")
def _():
    os.environ["CSB_NO_FUSE"] = "1"
    try:
        c2 = csb.Context(0, 1)
    finally:
        os.environ.pop("CSB_NO_FUSE", None)
    global ctx
    old = ctx
    ctx = c2
    try:
        r, _pts = dog_check(g1, 6, 0.0, 0.1, "gray1")
    finally:
        ctx = old
        c2.close()
    return r


@stage("keypoints_gray1_vs_oracle")
def _():
    p = csb.make_params(6, 0.0, 0.1, 10.0, 0.0)
    ours = ctx.extract(g1, p, max_pts=32768)
    orc, n, mpb = O.extract(g1, 6, 0.0, 0.1, 10.0, 0.0, False, 32768)
    np.save(WORK / "ours_gray1.npy", ours)
    r = PU.compare_keypoints(ours, orc)
    r["oct_ours"] = PU.per_octave_counts(ours)
    r["oct_orc"] = PU.per_octave_counts(orc)
    r["max_per_block"] = mpb
    return r


@stage("keypoints_gray1_host_api_vs_device_api")
def _():
    p = csb.make_params(6, 0.0, 0.1, 10.0, 0.0)
    a = ctx.extract(g1, p, max_pts=32768)
    b = ctx.extract(g1, p, max_pts=32768, from_host=True)
    return PU.compare_keypoints(b, a)


@stage("ref_available")
def _():
    return {"exists": O.ref_available()}


@stage("ref_extract_plain_gray1")
def _():
    ref = O.ref_extract(g1, WORK, 6, 0.0, 0.1, 10.0, 0.0, False, 32768, safe=False, tag="plain_g1")
    np.save(WORK / "ref_plain_gray1.npy", ref)
    orc, n, mpb = O.extract(g1, 6, 0.0, 0.1, 10.0, 0.0, False, 32768)
    return {"n": len(ref), "vs_oracle": PU.compare_keypoints(orc, ref)}


@stage("ref_extract_safe_gray1_vs_ours")
def _():
    ref = O.ref_extract(g1, WORK, 6, 0.0, 0.1, 10.0, 0.0, False, 32768, safe=True, tag="safe_g1")
    np.save(WORK / "ref_safe_gray1.npy", ref)
    ours = np.load(WORK / "ours_gray1.npy")
    r = PU.compare_keypoints(ours, ref)
    r["oct_ref"] = PU.per_octave_counts(ref)
    return r


@stage("ref_twice_gray1")
def _():
    a = O.ref_extract(g1, WORK, 6, 0.0, 0.1, 10.0, 0.0, False, 32768, safe=True, tag="safe_g1b")
    b = np.load(WORK / "ref_safe_gray1.npy")
    return PU.compare_keypoints(a, b)


@stage("ref_stages_gray1_vs_oracle")
def _():
    st = O.ref_stages(g1, WORK, 6, 0.0, tag="st_g1")
    res = {}
    for o, (base, dog) in enumerate(st):
        ob, od = O.octave_stage(g1, o, 0.0)
        res[f"o{o}_base"] = arr_diff(base, ob)
        res[f"o{o}_dog"] = arr_diff(dog, od)
    return res


@stage("rootsift_gray1")
def _():
    p = csb.make_params(6, 0.0, 0.1, 10.0, 0.0, rootsift=True)
    ours = ctx.extract(g1, p, max_pts=32768)
    ref = O.ref_extract(g1, WORK, 6, 0.0, 0.1, 10.0, 0.0, True, 32768, safe=True, tag="root_g1")
    r = PU.compare_keypoints(ours, ref)
    plain = np.load(WORK / "ours_gray1.npy")
    # stand-alone conversion of the plain descriptors must equal the fused path
    conv = ctx.rootsift(plain)
    ia, ib, _, _ = PU.match_sets(conv, ours)
    r["standalone_vs_fused_max"] = float(np.abs(conv["data"][ia] - ours["data"][ib]).max()) if len(ia) else None
    r["vs_oracle_rootsift_max"] = float(np.abs(O.rootsift(plain)["data"] - conv["data"]).max())
    return r


SYN = {}


@stage("synth_1080p_ours_vs_ref")
def _():
    img = csb.synth(1920, 1080, 1000)
    SYN["img"] = img
    p = csb.make_params(5, 0.0, 1.0, 10.0, 0.0)
    ours = ctx.extract(img, p, max_pts=32768)
    SYN["ours"] = ours
    ref = O.ref_extract(img, WORK, 5, 0.0, 1.0, 10.0, 0.0, False, 32768, safe=True, tag="syn1080")
    r = PU.compare_keypoints(ours, ref)
    r["oct_ours"] = PU.per_octave_counts(ours)
    r["oct_ref"] = PU.per_octave_counts(ref)
    return r


@stage("synth_1080p_ref_plain")
def _():
    img = SYN["img"]
    ref = O.ref_extract(img, WORK, 5, 0.0, 1.0, 10.0, 0.0, False, 32768, safe=False, tag="syn1080p")
    r = PU.compare_keypoints(SYN["ours"], ref)
    r["oct_ref"] = PU.per_octave_counts(ref)
    return r


@stage("synth_1080p_vs_oracle")
def _():
    orc, n, mpb = O.extract(SYN["img"], 5, 0.0, 1.0, 10.0, 0.0, False, 32768)
    r = PU.compare_keypoints(SYN["ours"], orc)
    r["max_per_block"] = mpb
    return r


@stage("match_fixture_vs_oracle_and_ref")
def _():
    s1 = O.read_vlfeat_sift(PU.GOLDEN / "sift1.bin")
    s2 = O.read_vlfeat_sift(PU.GOLDEN / "sift2.bin")
    res = {}
    for dist in ("l2", "dot"):
        ours = ctx.match(s1, s2, dist)
        orc = O.match(s1, s2, dist)
        ref, nm = O.ref_match(s1, s2, WORK, dist, 1000.0, 0.6, tag="fx_" + dist)
        r = {}
        for nm_, other in (("oracle", orc), ("ref", ref)):
            r[nm_] = {f: int(np.sum(ours[f] != other[f])) for f in ("score", "ambiguity", "match", "match_xpos", "match_ypos")}
        r["ref_matches_0.6"] = nm
        r["ours_matches_0.6"] = int(np.sum((ours["score"] < 1000.0 ** 2) & (ours["ambiguity"] < np.float32(0.6) ** 2)))
        res[dist] = r
    i, j = O.read_match_indices(PU.GOLDEN / "match_indices1_2.bin")
    ours = ctx.match(s1, s2, "l2")
    res["matlab_agree"] = int((ours["match"][i - 1] + 1 == j).sum())
    return res


@stage("pipeline_c1_match_homography")
def _():
    a, b = PU.preblur(g1), PU.preblur(g2)
    p = csb.make_params(6, 0.0, 0.1, 10.0, 0.0)
    k1 = ctx.extract(a, p, max_pts=32768)
    k2 = ctx.extract(b, p, max_pts=32768)
    m = ctx.match(k1, k2, "l2")
    orc = O.match(k1, k2, "l2")
    res = {"n1": len(k1), "n2": len(k2)}
    res["match_vs_oracle"] = {f: int(np.sum(m[f] != orc[f])) for f in ("score", "ambiguity", "match")}
    valid = O.valid_points(m, 0.0, 0.80)
    rng = np.random.default_rng(5)
    loops = 2048
    rp = np.zeros((4, loops), np.int32)
    for l in range(loops):
        rp[:, l] = valid[rng.choice(len(valid), 4, replace=False)]
    H, cnt = ctx.find_homography(m, rp, 5.0)
    Ho, cnto = O.find_homography(m, rp, 5.0)
    res["valid"] = int(len(valid))
    res["inliers_ours"] = cnt
    res["inliers_oracle"] = cnto
    res["H_ours"] = H.tolist()
    res["H_oracle"] = Ho.tolist()
    refh = O.ref_homography(m, WORK, 2048, 0.0, 0.80, 5.0, 5, 3.0, tag="c1")
    res["ref_num_matches"] = refh["num_matches"]
    res["ref_H"] = refh["H"].tolist()
    return res


@stage("batch_api")
def _():
    imgs = [csb.synth(640, 480, 100 + i) for i in range(6)]
    p = csb.make_params(5, 0.0, 1.0, 10.0, 0.0)
    single = [ctx.extract(im, p, max_pts=8192) for im in imgs]
    dptrs, pitch = [], None
    for im in imgs:
        d, pitch = ctx.upload_image(im)
        dptrs.append(d)
    dsifts = [ctx.alloc(588 * 8192) for _ in imgs]
    pins = [csb.PinnedArray(8192) for _ in imgs]
    counts = ctx.extract_batch(dptrs, 640, 480, pitch, p, dsifts, [pa.ptr for pa in pins], 8192)
    res = {"counts": counts.tolist(), "single": [len(s) for s in single]}
    worst = 0
    for i in range(len(imgs)):
        r = PU.compare_keypoints(pins[i].array[: counts[i]].copy(), single[i])
        worst = max(worst, r["only_ours"] + r["only_ref"])
    res["worst_set_diff"] = worst
    # host-frame batch with pageable outputs
    hs = [np.zeros(8192, csb.SIFT_DTYPE) for _ in imgs]
    hostimgs = [np.ascontiguousarray(im) for im in imgs]
    counts2 = ctx.extract_batch([im.ctypes.data for im in hostimgs], 640, 480, 640, p, dsifts,
                                [h.ctypes.data for h in hs], 8192, on_host=True)
    res["counts_host"] = counts2.tolist()
    for d in dptrs + dsifts:
        ctx.free(d)
    return res


@stage("timing_1080p")
def _():
    img = SYN.get("img")
    if img is None:
        img = csb.synth(1920, 1080, 1000)
    p = csb.make_params(5, 0.0, 1.0, 10.0, 0.0)
    d_img, pitch = ctx.upload_image(img)
    n_buf = 8
    dsifts = [ctx.alloc(588 * 16384) for _ in range(n_buf)]
    pins = [csb.PinnedArray(16384) for _ in range(n_buf)]
    res = {}
    # latency: single frame, profile per kernel
    for _ in range(5):
        ctx.extract_batch([d_img], 1920, 1080, pitch, p, dsifts[:1], [pins[0].ptr], 16384)
    t0 = time.perf_counter()
    n = 50
    for _ in range(n):
        ctx.extract_batch([d_img], 1920, 1080, pitch, p, dsifts[:1], [pins[0].ptr], 16384)
    res["latency_ms"] = (time.perf_counter() - t0) / n * 1e3
    ctx.profile(True)
    ctx.profile_reset()
    for _ in range(20):
        ctx.extract_batch([d_img], 1920, 1080, pitch, p, dsifts[:1], [pins[0].ptr], 16384)
    tab = ctx.profile_table()
    ctx.profile(False)
    res["kernels_ms_per_frame"] = {k: v["total_ms"] / 20 for k, v in tab.items()}
    # throughput: 256 frames through the slot pipeline
    frames = 256
    args = ([d_img] * frames, 1920, 1080, pitch, p, [dsifts[i % n_buf] for i in range(frames)],
            [pins[i % n_buf].ptr for i in range(frames)], 16384)
    ctx.extract_batch(*args)
    t0 = time.perf_counter()
    counts = ctx.extract_batch(*args)
    dt = time.perf_counter() - t0
    res["throughput_fps"] = frames / dt
    res["kp_per_frame"] = int(counts[0])
    return res


@stage("timing_ref_1080p")
def _():
    import subprocess
    img = SYN.get("img")
    raw = WORK / "bench1080.f32"
    np.ascontiguousarray(img, np.float32).tofile(raw)
    out = subprocess.run([str(O.REF_DRIVER), "bench", str(raw), "1920", "1080", "5", "0.0", "1.0", "10.0", "0.0",
                          "16384", "3", "20", "1"], capture_output=True, text=True, timeout=300)
    res = {"safe": out.stdout.strip()[-300:], "rc": out.returncode, "err": out.stderr[-300:]}
    out = subprocess.run([str(O.REF_DRIVER), "bench", str(raw), "1920", "1080", "5", "0.0", "1.0", "10.0", "0.0",
                          "16384", "3", "20", "0"], capture_output=True, text=True, timeout=300)
    res["plain"] = out.stdout.strip()[-300:]
    res["plain_rc"] = out.returncode
    res["plain_err"] = out.stderr[-300:]
    return res


import shutil  # noqa: E402
shutil.rmtree(WORK, ignore_errors=True)   # keep gpurun_out small (64 MiB cap)
print("DIAG DONE")
(OUT / ("diag_quick.json" if QUICK else "diag.json")).write_text(json.dumps(REPORT, indent=1, default=str))
