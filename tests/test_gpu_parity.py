"""GPU parity tests: the CUDA path (through the C ABI) against
  (1) oracle/oracle.c, the CPU restatement pinned to the reference's goldens, and
  (2) oracle/_ref/ref_driver, the UNMODIFIED reference built for sm_100a,
on the same inputs.  Tolerances are the ones BASELINE.json's north_star states:
counts / match indices exact, positions & scales 1e-3 px, orientations 1e-3 rad,
descriptors 1e-4 relative L2 — with one documented caveat: the reference does not
meet the descriptor bound against ITSELF run-to-run (shared-memory float atomics +
the texture unit's 8-bit weights make ~0.4 % of descriptors differ by up to 4e-4,
measured by test_reference_self_consistency), so the descriptor assertions are
">= 99 % within 1e-4 and 100 % within 1e-3".
"""
import ctypes
import os

import numpy as np
import pytest

import cusift_b200 as csb
import parity_utils as PU
from oracle import oracle as O

pytestmark = pytest.mark.gpu

needs_ref = pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref/ref_driver not built")


@pytest.fixture(scope="module")
def frames():
    return PU.golden_frames()


@pytest.fixture(scope="module")
def synth1080():
    return csb.synth(1920, 1080, 1000)


# ----------------------------------------------------------------- pyramid ---
@pytest.mark.parametrize("shape,seed", [((480, 640), None), ((135, 241), 7), ((1080, 1920), 1000), ((67, 120), 3)])
def test_scale_down_bit_exact(gpu_ctx, frames, shape, seed):
    img = frames[0] if seed is None else csb.synth(shape[1], shape[0], seed)
    assert np.array_equal(gpu_ctx.scale_down(img), O.scale_down(img))
    for variance in (0.3, 1.0, 2.5):                          # ScaleDown(res, src, variance) takes any variance (cuSIFT.cu:313-338)
        assert np.array_equal(gpu_ctx.scale_down(img, variance), O.scale_down(img, variance)), variance


@pytest.mark.parametrize("case", ["gray1", "synth_500x300", "synth_odd_173x131", "initblur"])
def test_dog_planes_and_octave_bases_bit_exact(gpu_ctx, frames, case):
    """Every DoG plane and every downsampled octave base equals the oracle's bit for bit
    (fused blur+DoG+downsample kernel)."""
    if case == "gray1":
        img, n_oct, ib = frames[0], 6, 0.0
    elif case == "synth_500x300":
        img, n_oct, ib = csb.synth(500, 300, 11), 4, 0.0
    elif case == "synth_odd_173x131":
        img, n_oct, ib = csb.synth(173, 131, 12), 3, 0.0
    else:
        img, n_oct, ib = csb.synth(400, 300, 13), 3, 0.5
    gpu_ctx.extract(img, csb.make_params(n_oct, ib, 1.0), max_pts=65536)
    for o in range(n_oct):
        base, dog = gpu_ctx.debug_octave(o)
        ob, od = O.octave_stage(img, o, ib)
        if o > 0:
            assert np.array_equal(base, ob), f"octave {o} base"
        assert np.array_equal(dog, od), f"octave {o} DoG"


def test_unfused_pipeline_identical(frames):
    """CSB_NO_FUSE=1 (stand-alone ScaleDown + blur/DoG kernels) gives the same planes."""
    os.environ["CSB_NO_FUSE"] = "1"
    try:
        ctx = csb.Context(0, 1)
    finally:
        os.environ.pop("CSB_NO_FUSE", None)
    try:
        img = frames[0]
        ctx.extract(img, csb.make_params(4, 0.0, 0.5), max_pts=65536)
        for o in range(4):
            _, dog = ctx.debug_octave(o)
            assert np.array_equal(dog, O.octave_stage(img, o, 0.0)[1])
    finally:
        ctx.close()


# --------------------------------------------------------------- extraction ---
def check_vs_oracle(ours, orc):
    r = PU.compare_keypoints(ours, orc)
    assert r["n_ours"] == r["n_ref"] == r["matched"], r
    assert r["only_ours"] == 0 and r["only_ref"] == 0
    # measured on the B200: positions 1.5e-5 px, scales 7.6e-6 (MUFU.RCP / ex2 vs exact division), orientations
    # 6.1e-5 deg, descriptors 97.9-99.5 % within 1e-4 relative L2 (median 1.7e-7, max 6.8e-4: a last-ulp angle
    # difference occasionally moves a vote across a bin edge).  Tolerances (north_star): 1e-3 px, 1e-3 rad, 1e-4.
    assert r["pos_max"] < 1e-4 and r["scale_max"] < 5e-5, r
    assert r["sharp_max"] < 1e-5 and r["edge_rel_max"] < 1e-5 and r["subs_equal"]
    assert r["ori_within_tol"] == 1.0 and r["ori_max_deg"] < 1e-3, r       # the oracle carries the fitted texture-filter model
    assert r["desc_within_tol"] >= 0.97 and r["desc_median"] < 1e-6 and r["desc_max"] < 2e-3, r
    return r


def test_extract_gray1_vs_oracle(gpu_ctx, frames):
    """test/detector.cpp parameters on color1.jpg (unsaturated)."""
    ours = gpu_ctx.extract(frames[0], csb.make_params(6, 0.0, 0.1, 10.0, 0.0), max_pts=32768)
    orc, n, mpb = O.extract(frames[0], 6, 0.0, 0.1, 10.0, 0.0, False, 32768)
    assert len(ours) == n == 9508 and mpb < 32
    assert PU.per_octave_counts(ours) == PU.per_octave_counts(orc)
    check_vs_oracle(ours, orc)
    # coarse octaves first, like ExtractSiftLoop (cuSIFT.cu:181-196)
    assert np.all(np.diff(ours["subsampling"]) <= 0)


def test_extract_matches_golden_file(gpu_ctx, frames):
    """All 4096 rows of the reference's own golden (test/data/cusift1_check) are reproduced."""
    from scipy.spatial import cKDTree
    ours = gpu_ctx.extract(frames[0], csb.make_params(6, 0.0, 0.1, 10.0, 0.0), max_pts=32768)
    gold = O.read_cusift_golden(PU.GOLDEN / "cusift1_check.bin")
    d, idx = cKDTree(PU.kp_key(ours)).query(gold[:, :3].astype(np.float64))
    assert d.max() < 1e-4
    do = PU.ang_diff_deg(ours["orientation"][idx], gold[:, 3])
    # the golden file was written by another GPU generation / CUDA version: statistical bound only
    assert np.mean(do <= PU.ORI_TOL_DEG) > 0.92 and np.median(do) < 0.005, (np.mean(do <= PU.ORI_TOL_DEG), do.max())


def test_extract_synth_vs_oracle(gpu_ctx):
    img = csb.synth(640, 480, 21)
    ours = gpu_ctx.extract(img, csb.make_params(5, 0.0, 1.0), max_pts=16384)
    orc, n, mpb = O.extract(img, 5, 0.0, 1.0, 10.0, 0.0, False, 16384)
    assert mpb < 32 and n == len(ours)
    check_vs_oracle(ours, orc)


@needs_ref
@pytest.mark.parametrize("case", ["gray1", "gray2_preblur", "synth_1080p", "synth_640_rootsift", "synth_4k_rootsift",
                                  "initblur_odd_size"])
def test_extract_vs_reference(gpu_ctx, frames, synth1080, workdir, case):
    """Ours vs the unmodified reference on the same GPU: identical keypoint sets with
    bit-identical x, y, scale, sharpness, edgeness; orientation and descriptors within
    the north_star tolerances."""
    root = False
    if case == "gray1":
        img, prm = frames[0], (6, 0.0, 0.1)
    elif case == "gray2_preblur":
        img, prm = PU.preblur(frames[1]), (6, 0.0, 0.1)          # main.cpp:308-309 pre-blur
    elif case == "synth_1080p":
        img, prm = synth1080, (5, 0.0, 1.0)
    elif case == "synth_4k_rootsift":                            # BASELINE config 3 at its own parameters:
        img, prm, root = csb.synth(3840, 2160, 2000), (5, 0.0, 0.1), True   # thresh 0.1 -> ~94.6 k keypoints, maxPts 131072
    elif case == "initblur_odd_size":
        img, prm = csb.synth(517, 389, 91), (4, 0.5, 0.3)
    else:
        img, prm, root = csb.synth(640, 480, 33), (5, 0.0, 0.5), True
    max_pts = 131072 if case == "synth_4k_rootsift" else 65536
    ours = gpu_ctx.extract(img, csb.make_params(*prm, 10.0, 0.0, rootsift=root), max_pts=max_pts)
    ref = O.ref_extract(img, workdir, prm[0], prm[1], prm[2], 10.0, 0.0, root, max_pts, safe=True, tag=case)
    if case == "synth_4k_rootsift":
        assert 90000 < len(ref) < 100000, len(ref)                # survey emulator: 94 585
    r = PU.compare_keypoints(ours, ref)
    assert r["n_ours"] == r["n_ref"] == r["matched"] and r["only_ours"] == 0 and r["only_ref"] == 0, r
    assert r["pos_exact"] == r["matched"], r                  # bit-exact positions and scales
    assert r["sharp_max"] == 0.0 and r["edge_rel_max"] == 0.0 and r["subs_equal"], r
    assert r["ori_max_deg"] < PU.ORI_TOL_DEG, r
    # the reference's descriptors vary run to run (float atomics): >= 99 % within 1e-4; the maximum over ~95 k RootSIFT
    # descriptors (sqrt amplifies differences in near-zero bins) reaches 1e-3 in some runs of the reference itself -
    # test_descriptor_error_is_within_the_reference_own_spread measures that spread
    assert r["desc_within_tol"] >= 0.99 and r["desc_max"] < (2e-3 if root else 1e-3), r
    assert PU.per_octave_counts(ours) == PU.per_octave_counts(ref)


DESC_Q = (0.5, 0.9, 0.99, 0.999, 1.0)


def _desc_err(a, b):
    ia, ib, oa, ob = PU.match_sets(a, b)
    assert len(oa) == 0 and len(ob) == 0
    return PU.desc_rel_l2(a["data"][ia], b["data"][ib]), PU.ang_diff_deg(a["orientation"][ia], b["orientation"][ib])


@needs_ref
@pytest.mark.parametrize("case", ["gray1", "synth_1080p"])
def test_descriptor_error_is_within_the_reference_own_spread(gpu_ctx, frames, synth1080, workdir, case):
    """north_star asks for descriptors within 1e-4 relative L2 of the reference.  The reference does not meet that
    against ITSELF: its descriptor votes are shared-memory float atomicAdds (cuSIFT_D.cu:234-253) and its orientation
    histogram too (:343), so two runs of the unmodified binary on the same frame differ.  This test MEASURES that
    spread (three reference runs, all three pairs pooled) next to ours-vs-reference (our deterministic result against
    each of the three runs, pooled) and asserts, quantile by quantile:
      * maximum: ours-vs-ref <= 1.5 x ref-vs-ref, or below 5e-4 (measured on the B200: 3.77e-4 vs 3.77e-4 in one
        session, 3.73e-4 vs 3.43e-4 in another - the maximum of the reference against itself moves between sessions
        with the order its atomics happen to land in, ours against it does not);
      * median / p90 / p99.9: within a factor 2.5 (+1e-7); p99: within a factor 5.  Measured (gray1, 9508
        keypoints): ref-vs-ref 4.8e-8 / 9.9e-8 / 2.5e-5 / 2.0e-4, ours-vs-ref 8.2e-8 / 1.2e-7 / 5.0e-5 / 2.2e-4.  The
        reference's atomics mostly land in the same order run to run, so it agrees with itself a little more often
        than with any other summation order; ours sums in a fixed tree order.  The p99 sits on the steep part of the
        distribution: the reference's own value was 2.5e-5, 2.5e-5 and 1.5e-5 in three sessions (ours 5.0e-5 to
        5.3e-5 in all of them), hence the wider factor there - the absolute 1e-4 bound below is the one that binds;
      * fraction within 1e-4: at most 0.5 % below the reference's own (measured 99.43 % vs 99.66 %);
      * p99 below north_star's 1e-4 in absolute terms.
    Both rows are printed (pytest -s) and written to gpurun_out/desc_spread_<case>.json."""
    import json
    img, prm = (frames[0], (6, 0.0, 0.1)) if case == "gray1" else (synth1080, (5, 0.0, 1.0))
    refs = [O.ref_extract(img, workdir, prm[0], prm[1], prm[2], 10.0, 0.0, False, 65536, safe=True, tag=f"spread{case}{k}")
            for k in range(3)]
    ours = gpu_ctx.extract(img, csb.make_params(*prm, 10.0, 0.0), max_pts=65536)
    rr = [_desc_err(refs[i], refs[j]) for i, j in ((0, 1), (0, 2), (1, 2))]
    orr = [_desc_err(ours, r) for r in refs]
    d_rr, d_or = np.concatenate([x[0] for x in rr]), np.concatenate([x[0] for x in orr])
    a_rr, a_or = np.concatenate([x[1] for x in rr]), np.concatenate([x[1] for x in orr])
    q_rr, q_or = np.quantile(d_rr, DESC_Q), np.quantile(d_or, DESC_Q)
    rec = {"case": case, "keypoints": int(len(ours)), "quantiles": list(DESC_Q),
           "ref_vs_ref": [float(x) for x in q_rr], "ours_vs_ref": [float(x) for x in q_or],
           "ref_vs_ref_within_1e-4": float(np.mean(d_rr <= 1e-4)), "ours_vs_ref_within_1e-4": float(np.mean(d_or <= 1e-4)),
           "ori_max_deg_ref_vs_ref": float(a_rr.max()), "ori_max_deg_ours_vs_ref": float(a_or.max())}
    print("\ndescriptor relative-L2 error, quantiles", DESC_Q, "\n  ref vs ref :", q_rr, "\n  ours vs ref:", q_or, "\n ", rec)
    try:
        out = PU.ROOT / "gpurun_out"
        out.mkdir(exist_ok=True)
        (out / f"desc_spread_{case}.json").write_text(json.dumps(rec))
    except OSError:
        pass
    for q, a, b in zip(DESC_Q[:-1], q_or[:-1], q_rr[:-1]):
        assert a <= (5.0 if q == 0.99 else 2.5) * b + 1e-7, (q, a, b, rec)
    assert q_or[-1] <= max(1.5 * q_rr[-1], 5e-4), rec
    assert q_or[2] < PU.DESC_TOL, rec
    assert rec["ours_vs_ref_within_1e-4"] >= rec["ref_vs_ref_within_1e-4"] - 0.005, rec
    assert a_or.max() <= max(2.0 * a_rr.max(), 1e-4) and a_or.max() < PU.ORI_TOL_DEG, rec


def test_host_entry_equals_device_entry(gpu_ctx, frames):
    """SiftData::Extract(float*) path (upload inside) == ExtractSift path (frame resident)."""
    p = csb.make_params(5, 0.0, 0.5)
    a = gpu_ctx.extract(frames[0], p, max_pts=32768)
    b = gpu_ctx.extract(frames[0], p, max_pts=32768, from_host=True)
    r = PU.compare_keypoints(a, b)
    assert r["matched"] == len(a) == len(b) and r["pos_exact"] == r["matched"]
    assert r["ori_max_deg"] < PU.ORI_TOL_DEG and r["desc_max"] < 1e-3


def test_saturation_and_parameters(gpu_ctx, frames):
    img = frames[0]
    full = gpu_ctx.extract(img, csb.make_params(6, 0.0, 0.1), max_pts=32768)
    # maxPts saturates like main.cpp's 4096: count clamps, coarse octaves survive
    sat = gpu_ctx.extract(img, csb.make_params(6, 0.0, 0.1), max_pts=4096)
    assert len(sat) == 4096
    assert (sat["subsampling"] > 1).sum() == (full["subsampling"] > 1).sum() == 1555
    # lowestScale skips fine octaves (cuSIFT.cu:194): lowest_scale=2 drops octave 0
    low = gpu_ctx.extract(img, csb.make_params(6, 0.0, 0.1, 10.0, 2.0), max_pts=32768)
    assert len(low) == 1555 and low["subsampling"].min() == 2.0
    # subsampling argument of Extract scales coordinates and scale
    sub2 = gpu_ctx.extract(img, csb.make_params(6, 0.0, 0.1, 10.0, 0.0, 2.0), max_pts=32768)
    r = PU.match_sets(sub2, full, tol=1e9)
    ia, ib = r[0], r[1]
    assert len(sub2) == len(full)
    k2 = np.sort(PU.kp_key(sub2), axis=0)
    k1 = np.sort(PU.kp_key(full) * 2.0, axis=0)
    assert np.allclose(k2, k1, rtol=0, atol=1e-3)
    # higher threshold / edge limit only remove points
    hi = gpu_ctx.extract(img, csb.make_params(6, 0.0, 2.0), max_pts=32768)
    ed = gpu_ctx.extract(img, csb.make_params(6, 0.0, 0.1, 5.0), max_pts=32768)
    assert 0 < len(hi) < len(full) and 0 < len(ed) < len(full)
    assert np.abs(hi["sharpness"]).min() > 1.5


def test_flat_and_tiny_frames(gpu_ctx):
    flat = np.full((64, 96), 77.0, np.float32)
    assert len(gpu_ctx.extract(flat, csb.make_params(3, 0.0, 0.1), max_pts=1024)) == 0
    tiny = csb.synth(40, 24, 5)
    pts = gpu_ctx.extract(tiny, csb.make_params(2, 0.0, 0.5), max_pts=1024)
    orc, n, _ = O.extract(tiny, 2, 0.0, 0.5, 10.0, 0.0, False, 1024)
    assert len(pts) == n


def test_bad_arguments_return_status_not_crash(gpu_ctx):
    L = csb.lib()
    p = csb.make_params(5, 0.0, 1.0)
    n = ctypes.c_int(0)
    assert L.csb_extract(gpu_ctx.h, None, 640, 480, 640, ctypes.byref(p), None, 16, None, ctypes.byref(n)) == 10001
    bad = csb.make_params(9, 0.0, 1.0)
    d = gpu_ctx.alloc(588 * 16)
    dimg, pitch = gpu_ctx.upload_image(np.zeros((480, 640), np.float32))
    assert L.csb_extract(gpu_ctx.h, dimg, 640, 480, pitch, ctypes.byref(bad), d, 16, None, ctypes.byref(n)) == 10003
    assert L.csb_extract(gpu_ctx.h, dimg, 640, 480, 100, ctypes.byref(p), d, 16, None, ctypes.byref(n)) == 10001
    assert b"pitch" in L.csb_last_error(gpu_ctx.h)
    H = np.zeros(9, np.float32)
    rp = np.zeros((4, 16), np.int32)
    assert L.csb_find_homography(gpu_ctx.h, d, 100, rp.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), 15, 5.0,
                                 H.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), ctypes.byref(n)) == 10001
    gpu_ctx.free(d)
    gpu_ctx.free(dimg)


def test_batch_equals_single_and_is_deterministic(gpu_ctx):
    imgs = [csb.synth(640, 480, 100 + i) for i in range(9)]
    p = csb.make_params(5, 0.0, 1.0)
    single = [gpu_ctx.extract(im, p, max_pts=8192) for im in imgs]
    dptrs = []
    for im in imgs:
        d, pitch = gpu_ctx.upload_image(im)
        dptrs.append(d)
    dsifts = [gpu_ctx.alloc(588 * 8192) for _ in imgs]
    pins = [csb.PinnedArray(8192) for _ in imgs]
    try:
        for _ in range(2):
            counts = gpu_ctx.extract_batch(dptrs, 640, 480, pitch, p, dsifts, [pa.ptr for pa in pins], 8192)
            assert counts.tolist() == [len(s) for s in single]
            for i in range(len(imgs)):
                r = PU.compare_keypoints(pins[i].array[: counts[i]].copy(), single[i])
                assert r["matched"] == counts[i] and r["pos_exact"] == r["matched"]
                assert r["desc_max"] < 1e-3
        # dense host frames + pageable result buffers through the same entry point
        hs = [np.zeros(8192, csb.SIFT_DTYPE) for _ in imgs]
        counts2 = gpu_ctx.extract_batch([im.ctypes.data for im in imgs], 640, 480, 640, p, dsifts,
                                        [h.ctypes.data for h in hs], 8192, on_host=True)
        assert counts2.tolist() == counts.tolist()
        r = PU.compare_keypoints(hs[3][: counts2[3]], single[3])
        assert r["matched"] == counts2[3] and r["pos_exact"] == r["matched"]
    finally:
        for d in dptrs + dsifts:
            gpu_ctx.free(d)
        for pa in pins:
            pa.free()


def test_batch_with_one_shared_device_buffer(gpu_ctx):
    """A caller may pass the same d_sift for every frame of a batch (it only wants the host copies):
    frames then wait for the buffer instead of racing on it."""
    imgs = [csb.synth(640, 480, 400 + i) for i in range(6)]
    p = csb.make_params(5, 0.0, 0.5)
    dev = [gpu_ctx.upload_image(im) for im in imgs]
    d_one = gpu_ctx.alloc(588 * 8192)
    pins = [csb.PinnedArray(8192) for _ in imgs]
    try:
        cnt = gpu_ctx.extract_batch([d for d, _ in dev], 640, 480, dev[0][1], p, [d_one] * len(imgs), [q.ptr for q in pins], 8192)
        for k, im in enumerate(imgs):
            want = PU.canonical_sort(gpu_ctx.extract(im, p, max_pts=8192))
            got = PU.canonical_sort(pins[k].array[: cnt[k]].copy())
            assert len(got) == len(want) > 500
            for f in ("coords2D", "scale", "orientation", "data"):
                assert np.array_equal(got[f], want[f]), (k, f)
    finally:
        gpu_ctx.free(d_one)
        for d, _ in dev:
            gpu_ctx.free(d)


def test_full_size_properties_4k_rootsift(gpu_ctx):
    """BASELINE config 3 (3840x2160, ExtractRootSift, ~95 k keypoints): size-independent
    properties — determinism of the keypoint set, unit-L2 RootSIFT descriptors, count in
    the oracle-calibrated range, coarse-first ordering."""
    img = csb.synth(3840, 2160, 2000)
    p = csb.make_params(5, 0.0, 0.1, rootsift=True)
    a = gpu_ctx.extract(img, p, max_pts=131072)
    b = gpu_ctx.extract(img, p, max_pts=131072)
    assert len(a) == len(b) and 90000 < len(a) < 100000, len(a)
    assert np.array_equal(np.sort(PU.kp_key(a), axis=0), np.sort(PU.kp_key(b), axis=0))
    assert np.all(np.diff(a["subsampling"]) <= 0)
    nrm = np.linalg.norm(a["data"].astype(np.float64), axis=1)
    assert np.abs(nrm - 1.0).max() < 1e-3                    # sqrt(L1-normalised) has unit L2 norm
    assert a["data"].min() >= 0.0
    assert 0 <= a["orientation"].min() and a["orientation"].max() < 360.0


# ------------------------------------------------------------------ RootSIFT ---
def test_rootsift_standalone_bit_exact(gpu_ctx, frames):
    pts = gpu_ctx.extract(frames[0], csb.make_params(4, 0.0, 1.0), max_pts=16384)
    conv = gpu_ctx.rootsift(pts)
    assert np.array_equal(conv["data"], O.rootsift(pts)["data"])       # fp32 serial sum + fp64 divide
    fused = gpu_ctx.extract(frames[0], csb.make_params(4, 0.0, 1.0, rootsift=True), max_pts=16384)
    ia, ib, oa, ob = PU.match_sets(fused, conv)
    assert len(oa) == 0 and len(ob) == 0
    assert PU.desc_rel_l2(fused["data"][ia], conv["data"][ib]).max() < 1e-3


# ------------------------------------------------------------------ matching ---
@pytest.mark.parametrize("dist", ["l2", "dot"])
def test_match_fixture_bit_exact(gpu_ctx, workdir, dist):
    s1 = O.read_vlfeat_sift(PU.GOLDEN / "sift1.bin")
    s2 = O.read_vlfeat_sift(PU.GOLDEN / "sift2.bin")
    ours = gpu_ctx.match(s1, s2, dist)
    orc = O.match(s1, s2, dist)
    for f in ("score", "ambiguity", "match", "match_xpos", "match_ypos"):
        assert np.array_equal(ours[f], orc[f]), f
    if dist == "l2":
        i, j = O.read_match_indices(PU.GOLDEN / "match_indices1_2.bin")
        assert int((ours["match"][i - 1] + 1 == j).sum()) == 326         # test/test.cpp:30-40
        assert O.count_matches(ours, 1000.0, 0.6) == 340                 # test/test.cpp:52-55
    if O.ref_available():
        ref, nm = O.ref_match(s1, s2, workdir, dist, 1000.0, 0.6, tag="fx" + dist)
        for f in ("score", "ambiguity", "match", "match_xpos", "match_ypos"):
            assert np.array_equal(ours[f], ref[f]), f
        assert nm == O.count_matches(ours, 1000.0, 0.6)


def test_match_ragged_sizes_ties_and_duplicates(gpu_ctx):
    rng = np.random.default_rng(9)
    for n1, n2 in ((1, 1), (5, 3), (17, 33), (100, 1000), (257, 15)):
        a = np.zeros(n1, csb.SIFT_DTYPE)
        b = np.zeros(n2, csb.SIFT_DTYPE)
        da = np.abs(rng.standard_normal((n1, 128))).astype(np.float32)
        db = np.abs(rng.standard_normal((n2, 128))).astype(np.float32)
        a["data"] = da / np.linalg.norm(da, axis=1, keepdims=True)
        b["data"] = db / np.linalg.norm(db, axis=1, keepdims=True)
        b["coords2D"] = rng.uniform(0, 100, (n2, 2)).astype(np.float32)
        if n2 > 20:                                          # exact duplicates -> tie-break rule
            b["data"][17] = b["data"][3]
            b["data"][24 % n2] = b["data"][3]
            a["data"][0] = b["data"][3]
        for dist in ("l2", "dot"):
            ours, orc = gpu_ctx.match(a, b, dist), O.match(a, b, dist)
            for f in ("score", "ambiguity", "match", "match_xpos", "match_ypos"):
                assert np.array_equal(ours[f], orc[f]), (n1, n2, dist, f)
    # empty sets: nothing to do (matching.cu:282-283)
    assert csb.lib().csb_match(gpu_ctx.h, None, 0, None, 0, 1, None) == 0


def _rand_set(n, seed):
    r = np.random.default_rng(seed)
    s = np.zeros(n, csb.SIFT_DTYPE)
    d = np.abs(r.standard_normal((n, 128))).astype(np.float32)
    s["data"] = d / np.linalg.norm(d, axis=1, keepdims=True)
    s["coords2D"] = r.uniform(0, 1000, (n, 2)).astype(np.float32)
    return s


@pytest.mark.parametrize("n1,n2", [(256, 256), (300, 700), (1000, 3000), (257, 4097), (512, 384), (9000, 2100), (13000, 700), (20000, 600)])
def test_match_tensor_core_path_bit_exact(gpu_ctx, n1, n2):
    """Sets of >= 256 points go through the tcgen05 matcher (fp16 tensor-core scan, fp32 rescoring in
    the reference's k order): still bit-identical to the oracle, incl. exact duplicates / ties.  The sizes cover one
    tile per candidate slice, a slice of pure padding, more query tiles than SMs / 4 (3, 2 and 1 candidate slices)."""
    a, b = _rand_set(n1, n1), _rand_set(n2, n2 + 1)
    b["data"][17] = b["data"][3]
    b["data"][24] = b["data"][3]
    a["data"][0] = b["data"][3]                              # three-way exact tie for query 0
    b["data"][n2 - 1] = a["data"][7]
    b["data"][100] = a["data"][7]                            # duplicate pair far apart (different tiles)
    for dist in ("l2", "dot"):
        ours, orc = gpu_ctx.match(a, b, dist), O.match(a, b, dist)
        for f in ("score", "ambiguity", "match", "match_xpos", "match_ypos"):
            assert np.array_equal(ours[f], orc[f]), (n1, n2, dist, f)


@pytest.mark.parametrize("n1,n2", [(600, 5000), (2048, 8192)])
def test_match_redo_path_bit_exact(n1, n2):
    """More than 16 near-tied candidates in one slice overflow the tensor-core short list; those blocks
    of 16 queries are redone by the exact kernel (candidates cut into slices, merged by k_match_finish).
    Twelve ties spread over the slices fit the lists and are resolved by the rescoring kernel."""
    ctx = csb.Context(0, 1)
    try:
        a, b = _rand_set(n1, 77), _rand_set(n2, 78)
        for q, cols in ((5, range(40, 40 + 21 * 16, 16)), (n1 - 1, range(7, n2, n2 // 11)), (300, range(1000, 1024))):
            for c in cols:
                b["data"][c] = a["data"][q]                   # >= 11 exact duplicates: all tie for best and second
        for dist in ("l2", "dot"):
            before = csb.lib().csb_match_redo_blocks(ctx.h)
            ours, orc = ctx.match(a, b, dist), O.match(a, b, dist)
            assert csb.lib().csb_match_redo_blocks(ctx.h) - before >= 2
            for f in ("score", "ambiguity", "match", "match_xpos", "match_ypos"):
                assert np.array_equal(ours[f], orc[f]), (n1, n2, dist, f)
    finally:
        ctx.close()


def test_match_tensor_core_equals_cuda_core_path(frames):
    """CSB_MATCH_EXACT=1 forces the fp32 CUDA-core kernel; both paths agree bit for bit on real descriptors."""
    os.environ["CSB_MATCH_EXACT"] = "1"
    try:
        exact_ctx = csb.Context(0, 1)
    finally:
        os.environ.pop("CSB_MATCH_EXACT", None)
    tc_ctx = csb.Context(0, 1)
    try:
        p = csb.make_params(5, 0.0, 0.3)
        k1 = tc_ctx.extract(frames[0], p, max_pts=32768)
        k2 = tc_ctx.extract(frames[1], p, max_pts=32768)
        assert len(k1) > 2000 and len(k2) > 2000
        m_tc, m_ex = tc_ctx.match(k1, k2, "l2"), exact_ctx.match(k1, k2, "l2")
        for f in ("score", "ambiguity", "match", "match_xpos", "match_ypos"):
            assert np.array_equal(m_tc[f], m_ex[f]), f
        redo = csb.lib().csb_match_redo_blocks(tc_ctx.h)
        assert redo <= (len(k1) + 15) // 16 // 4              # the exact redo path stays the exception
    finally:
        exact_ctx.close()
        tc_ctx.close()


def test_match_full_size_properties(gpu_ctx):
    """8192 x 8192 (BASELINE config 5 pair size): self-match is the identity with score ~0."""
    rng = np.random.default_rng(4)
    n = 8192
    a = np.zeros(n, csb.SIFT_DTYPE)
    d = np.abs(rng.standard_normal((n, 128))).astype(np.float32)
    a["data"] = d / np.linalg.norm(d, axis=1, keepdims=True)
    m = gpu_ctx.match(a, a, "l2")
    assert np.array_equal(m["match"], np.arange(n))
    assert np.abs(m["score"]).max() < 1e-5 and m["ambiguity"].max() < 1e-3
    sub = O.match(a[:64], a, "l2")
    for f in ("score", "ambiguity", "match"):
        assert np.array_equal(m[f][:64], sub[f])


# ---------------------------------------------------------------- homography ---
def glibc_rand_samples(valid, num_loops):
    """FindHomography's sampling loop (homography.cu:232-244) with libc rand(), seed 1."""
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1)
    nv = len(valid)
    rp = np.zeros((4, num_loops), np.int32)
    for i in range(num_loops):
        p1, p2, p3, p4 = (libc.rand() % nv for _ in range(4))
        while p2 == p1:
            p2 = libc.rand() % nv
        while p3 == p1 or p3 == p2:
            p3 = libc.rand() % nv
        while p4 == p1 or p4 == p2 or p4 == p3:
            p4 = libc.rand() % nv
        rp[:, i] = valid[[p1, p2, p3, p4]]
    return rp


def test_c1_pipeline_match_and_homography(gpu_ctx, frames, workdir):
    """BASELINE config 1 (main.cpp demo): pre-blur, ExtractSift x2, MatchSiftData,
    FindHomography(numLoops, 0.0, 0.80, 5.0) on the reference's own image pair."""
    a, b = PU.preblur(frames[0]), PU.preblur(frames[1])
    p = csb.make_params(6, 0.0, 0.1)
    k1 = gpu_ctx.extract(a, p, max_pts=32768)
    k2 = gpu_ctx.extract(b, p, max_pts=32768)
    assert 10000 < len(k1) < 11500 and 10000 < len(k2) < 11500   # emulator: 10 837 / 10 869
    m = gpu_ctx.match(k1, k2, "l2")
    orc = O.match(k1, k2, "l2")
    for f in ("score", "ambiguity", "match", "match_xpos", "match_ypos"):
        assert np.array_equal(m[f], orc[f]), f
    valid = O.valid_points(m, 0.0, 0.80)
    loops = 2048
    rp = glibc_rand_samples(valid, loops)
    H, cnt = gpu_ctx.find_homography(m, rp, 5.0)
    Ho, cnto = O.find_homography(m, rp, 5.0)
    assert cnt == cnto and cnt > 2500
    assert np.allclose(H, Ho, rtol=1e-4, atol=1e-6)
    assert abs(H[0] - 1) < 0.05 and abs(H[4] - 1) < 0.05          # consecutive RGB-D frames: near identity
    if O.ref_available():
        # the reference draws the same rand() sequence (fresh process, seed 1)
        refh = O.ref_homography(m, workdir, loops, 0.0, 0.80, 5.0, 0, 3.0, tag="c1h")
        n16 = len(m) % 16
        assert abs(refh["num_matches"] - cnt) <= (16 - n16 if n16 else 0)   # its pad slots are uninitialised
        if refh["num_matches"] == cnt:                        # same winning hypothesis
            assert np.allclose(refh["H"], H, rtol=1e-3, atol=1e-5)
        H2, nfit, _ = O.improve_homography(m, H, 5, 0.0, 0.80, 3.0)
        assert nfit > 1500


def test_allpairs_match_ransac_vs_oracle(gpu_ctx):
    """BASELINE config 5 in miniature: every unordered pair of 5 keypoint sets, MatchSiftData +
    FindHomography batched on the device; per pair the result equals the oracle's for the same samples."""
    imgs = [csb.synth(640, 480, 3000 + i) for i in range(5)]
    p = csb.make_params(5, 0.0, 0.5)
    sets = []
    for im in imgs:
        k = gpu_ctx.extract(im, p, max_pts=8192)
        order = np.lexsort((k["scale"], k["coords2D"][:, 1], k["coords2D"][:, 0], k["subsampling"]))   # canonical sort
        sets.append(np.ascontiguousarray(k[order][:1536]))
    assert all(len(s) == 1536 for s in sets)
    dptrs = [gpu_ctx.upload_sift(s) for s in sets]
    pairs = csb.all_pairs(len(sets))
    loops, seed = 256, 7
    try:
        H, inl, nv = gpu_ctx.allpairs(dptrs, [len(s) for s in sets], pairs, "l2", loops, 0.0, 0.80, 5.0, seed)
        for k, (i, j) in enumerate(pairs):
            m = O.match(sets[i], sets[j], "l2")
            valid = O.valid_points(m, 0.0, 0.80)
            assert nv[k] == len(valid), (k, nv[k], len(valid))
            rp = gpu_ctx.sample_points(valid, loops, seed, k)
            Ho, cnto = O.find_homography(m, rp, 5.0)
            assert inl[k] == cnto, (k, inl[k], cnto)
            assert np.allclose(H[k], Ho, rtol=1e-3, atol=1e-5), (k, H[k], Ho)
        # the last pair processed with query set i leaves its match fields in set i (as in the reference)
        last = {}
        for (i, j) in pairs:
            last[i] = j
        for i, j in last.items():
            dev = gpu_ctx.download_sift(dptrs[i], len(sets[i]))
            ref = O.match(sets[i], sets[j], "l2")
            for f in ("score", "ambiguity", "match"):
                assert np.array_equal(dev[f], ref[f]), (i, j, f)
    finally:
        for d in dptrs:
            gpu_ctx.free(d)


def test_extrema_dense_fallback_matches_list_path(frames, tmp_path):
    """A tile whose flagged pixels exceed the per-CTA list is re-examined pixel by pixel; with the list
    shrunk to 4 entries (CSB_XT_CAP) practically every tile takes that path and the keypoints stay the same."""
    import subprocess
    import sys
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import cusift_b200 as csb, parity_utils as PU\n"
        "g1, _ = PU.golden_frames(); ctx = csb.Context(0, 1)\n"
        "k = ctx.extract(PU.preblur(g1), csb.make_params(5, 0.0, 0.5), max_pts=32768)\n"
        "np.save(sys.argv[1], PU.canonical_sort(k))\n" % (str(PU.ROOT), str(PU.ROOT / "tests")))
    outs = []
    for cap in ("", "4"):
        env = dict(os.environ)
        env.pop("CSB_XT_CAP", None)
        if cap:
            env["CSB_XT_CAP"] = cap
        out = tmp_path / f"k{cap or 'dflt'}.npy"
        subprocess.run([sys.executable, "-c", code, str(out)], check=True, env=env, timeout=300)
        outs.append(np.load(out))
    a, b = outs
    assert len(a) == len(b) and len(a) > 1000
    for f in ("coords2D", "scale", "sharpness", "edgeness", "orientation", "subsampling"):
        assert np.array_equal(a[f], b[f]), f


# ------------------------------------------------------------------- ingest ---
@pytest.mark.parametrize("shape", [(480, 640), (67, 121), (1, 5), (3, 1)])
def test_ingest_u8_matches_opencv_bit_for_bit(gpu_ctx, shape):
    """main.cpp:301-309: imread(.,0).convertTo(CV_32FC1) then GaussianBlur(Size(3,3), 0.5), done on the device."""
    import cv2
    u8 = np.random.default_rng(5).integers(0, 256, shape, dtype=np.uint8)
    assert np.array_equal(gpu_ctx.ingest_u8(u8, False), u8.astype(np.float32))
    want = cv2.GaussianBlur(u8.astype(np.float32), (3, 3), 0.5)
    got = gpu_ctx.ingest_u8(u8, True)
    if shape[1] % 16 == 0:
        assert np.array_equal(got, want.reshape(shape))        # OpenCV's SIMD body: fma(c,k0,(l+r)k1) / fma(u+d,k1,c k0)
    else:                                                       # its scalar row tails round differently (not fused)
        np.testing.assert_allclose(got, want.reshape(shape), rtol=2.5e-7, atol=0)


def test_extract_batch_u8_equals_float_path(gpu_ctx, frames):
    """8-bit upload + device pre-blur gives the same keypoints as the reference flow (host convertTo +
    GaussianBlur + float upload), bit for bit."""
    g1, g2 = PU.golden_frames()
    p = csb.make_params(5, 0.0, 0.5)
    imgs = [np.ascontiguousarray(g, np.uint8) for g in (g1, g2, g1)]
    pins = [csb.PinnedArray(8192) for _ in imgs]
    ds = [gpu_ctx.alloc(588 * 8192) for _ in imgs]
    try:
        cnt = gpu_ctx.extract_batch_u8([im.ctypes.data for im in imgs], 640, 480, 640, True, p, ds, [q.ptr for q in pins], 8192)
        for k, im in enumerate(imgs):
            want = PU.canonical_sort(gpu_ctx.extract(PU.preblur(im), p, max_pts=8192))
            got = PU.canonical_sort(pins[k].array[: cnt[k]].copy())
            assert len(got) == len(want) > 1000
            for f in ("coords2D", "scale", "sharpness", "edgeness", "orientation", "data"):
                assert np.array_equal(got[f], want[f]), (k, f)
    finally:
        for d in ds:
            gpu_ctx.free(d)


# ---------------------------------------------------------- rigid transform ---
def _rigid_scene(n, n_out, seed):
    r = np.random.default_rng(seed)
    ang = r.uniform(-0.6, 0.6, 3)
    cx, sx, cy, sy, cz, sz = np.cos(ang[0]), np.sin(ang[0]), np.cos(ang[1]), np.sin(ang[1]), np.cos(ang[2]), np.sin(ang[2])
    R = (np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]) @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]) @
         np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]))
    t = r.uniform(-0.5, 0.5, 3)
    mov = r.uniform(-1, 1, (n, 3)) + np.array([0, 0, 2.0])
    ref = mov @ R.T + t + r.normal(0, 0.002, (n, 3))
    ref[:n_out] = r.uniform(-2, 2, (n_out, 3))
    return np.ascontiguousarray(np.concatenate([ref, mov], 1), np.float32), R, t


def test_rigid_transform_golden_and_reference(gpu_ctx, workdir):
    """EstimateRigidTransformH on the reference's golden file with its stored indices: equals the numpy oracle,
    the MATLAB Rt stored in the file and the unmodified reference run on this GPU."""
    coord, idx, Rt_matlab = O.read_matlab_ransac(PU.GOLDEN / "rigid_ransac.bin")
    Rt, n_inl, mask = gpu_ctx.rigid_transform(coord, idx, len(idx), 0.05 * 0.05, True)
    Rt_o, n_o, mask_o = O.rigid_transform(coord, idx, 0.05 * 0.05, True)
    assert n_inl == n_o == 114 and np.array_equal(mask, mask_o)
    assert np.abs(Rt - Rt_o).max() < 1e-5 and np.abs(Rt - Rt_matlab).max() < 1e-5
    Rt_r, n_r, mask_r = O.ref_rigid(coord, idx, 0.05 * 0.05, True, workdir)
    assert n_r == n_inl and np.array_equal(mask_r, mask)
    assert np.abs(Rt - Rt_r).max() < 1e-4                    # the reference's float SVD vs eigen-decomposition in double


@pytest.mark.parametrize("type3d", [True, False])
def test_rigid_transform_synthetic_vs_oracle(gpu_ctx, type3d):
    coord, R, t = _rigid_scene(2000, 600, 11)
    if not type3d:                                           # planar motion about y for the two-point fit
        a = 0.4
        R = np.array([[np.cos(a), 0, -np.sin(a)], [0, 1, 0], [np.sin(a), 0, np.cos(a)]])
        mov = coord[:, 3:].astype(np.float64)
        ref = mov @ R.T + np.array([0.3, 0.0, -0.2])
        ref[:600] = np.random.default_rng(3).uniform(-2, 2, (600, 3))
        coord = np.ascontiguousarray(np.concatenate([ref, mov], 1), np.float32)
    r = np.random.default_rng(12)
    idx = np.stack([r.choice(2000, 3, replace=False) for _ in range(256)]).astype(np.int32)
    Rt, n_inl, mask = gpu_ctx.rigid_transform(coord, idx, 256, 0.01 * 0.01, type3d)
    Rt_o, n_o, mask_o = O.rigid_transform(coord, idx, 0.01 * 0.01, type3d)
    assert abs(n_inl - n_o) <= 2 and (mask != mask_o).sum() <= 2   # borderline points may flip with FMA contraction
    assert n_inl > 1200
    assert np.abs(Rt - Rt_o).max() < 2e-4
    assert np.abs(Rt.reshape(3, 4)[:, :3] - R).max() < 5e-3


def test_rigid_transform_device_sampling_and_edges(gpu_ctx):
    coord, R, t = _rigid_scene(500, 100, 21)
    Rt, n_inl, mask = gpu_ctx.rigid_transform(coord, None, 512, 0.01 * 0.01, True, seed=5)
    assert n_inl >= 380 and mask.sum() == n_inl
    assert np.abs(Rt.reshape(3, 4)[:, :3] - R).max() < 5e-3 and np.abs(Rt.reshape(3, 4)[:, 3] - t).max() < 5e-3
    again = gpu_ctx.rigid_transform(coord, None, 512, 0.01 * 0.01, True, seed=5)
    assert np.array_equal(again[0], Rt) and again[1] == n_inl          # deterministic for a given seed
    # the generator restated on the host draws the same indices
    L = csb.lib()
    idx = np.zeros((512, 3), np.int32)
    for l in range(512):
        picks = []
        for k in range(3):
            a = 0
            while True:
                c = L.csb_rigid_sample_hash(5, l, k, a) % 500
                a += 1
                if c not in picks:
                    picks.append(c)
                    break
        idx[l] = picks
    same = gpu_ctx.rigid_transform(coord, idx, 512, 0.01 * 0.01, True)
    assert np.array_equal(same[0], Rt) and same[1] == n_inl
    # fewer than 3 points: identity, no inliers (nothing to fit)
    Rt0, n0, _ = gpu_ctx.rigid_transform(coord[:2], None, 16, 1.0, True)
    assert n0 == 0 and np.array_equal(Rt0.reshape(3, 4), np.eye(3, 4, dtype=np.float32))


# ------------------------------------------------------- verdict r1: parity gaps ---
def _c1_matched(gpu_ctx, frames):
    a, b = PU.preblur(frames[0]), PU.preblur(frames[1])
    p = csb.make_params(6, 0.0, 0.1)
    k1 = gpu_ctx.extract(a, p, max_pts=32768)
    k2 = gpu_ctx.extract(b, p, max_pts=32768)
    return gpu_ctx.match(k1, k2, "l2")


def test_improve_homography_host_shim_device_oracle_reference(gpu_ctx, frames, workdir):
    """ImproveHomography (extras/homography.cu:271-337, main.cpp:335): the drop-in C++ function (host, OpenCV-free), the
    device version behind csb_improve_homography, the CPU oracle and - where built - the UNMODIFIED reference, all
    started from the same H on the same matched points: numFit equal, H within 1e-6, match_error within 1e-5."""
    import json
    import subprocess
    m = _c1_matched(gpu_ctx, frames)
    valid = O.valid_points(m, 0.0, 0.80)
    if O.ref_available():
        refh = O.ref_homography(m, workdir, 2048, 0.0, 0.80, 5.0, 5, 3.0, tag="imp")
        H0 = refh["H"].astype(np.float32)
    else:
        rp = glibc_rand_samples(valid, 2048)
        H0, _ = gpu_ctx.find_homography(m, rp, 5.0)
    H_o, nfit_o, pts_o = O.improve_homography(m, H0, 5, 0.0, 0.80, 3.0)
    assert nfit_o > 1500
    # device
    H_d, nfit_d, pts_d = gpu_ctx.improve_homography(m, H0, 5, 0.0, 0.80, 3.0)
    assert nfit_d == nfit_o, (nfit_d, nfit_o)
    assert np.allclose(H_d, H_o, rtol=1e-6, atol=1e-6), (H_d, H_o)
    assert np.allclose(pts_d["match_error"], pts_o["match_error"], rtol=1e-5, atol=1e-5)
    # C++ shim (host path)
    exe = PU.ROOT / "build" / "csb_improve"
    assert exe.exists(), "build/csb_improve missing: run make demo"
    fin, fout = workdir / "imp_in.sift", workdir / "imp_out.sift"
    O.write_sift_file(fin, m)
    res = subprocess.run([str(exe), str(fin), str(fout), "5", "0.0", "0.80", "3.0"] + ["%.9g" % v for v in H0],
                         capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stderr
    js = json.loads(res.stdout.strip().splitlines()[-1])
    assert js["numFit"] == nfit_o
    assert np.allclose(np.array(js["H"], np.float32), H_o, rtol=1e-6, atol=1e-6)
    pts_s = O.read_sift_file(fout)
    assert np.allclose(pts_s["match_error"], pts_o["match_error"], rtol=1e-5, atol=1e-5)
    if O.ref_available():
        assert refh["num_fit"] == nfit_o, (refh["num_fit"], nfit_o)
        assert np.allclose(refh["H_improved"], H_o, rtol=1e-6, atol=1e-6), (refh["H_improved"], H_o)
    # no loops: only the inlier count / match_error pass
    H_z, nfit_z, _ = gpu_ctx.improve_homography(m, H0, 0, 0.0, 0.80, 3.0)
    H_zo, nfit_zo, _ = O.improve_homography(m, H0, 0, 0.0, 0.80, 3.0)
    assert nfit_z == nfit_zo and np.allclose(H_z, H_zo, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("n", [7, 300, 8192, 12345])
def test_improve_homography_cluster_kernel_sizes(gpu_ctx, n):
    """The device ImproveHomography runs as one 8-CTA cluster per job; a thread keeps its first four points in
    registers and reads further ones from memory.  Sizes around those limits (fewer points than threads, exactly
    8192 = 4 per thread, more), with outliers and filtered-out points, against the oracle."""
    r = np.random.default_rng(100 + n)
    pts = np.zeros(n, csb.SIFT_DTYPE)
    xy = r.uniform(0, 1000, (n, 2)).astype(np.float32)
    Ht = np.array([1.02, 0.01, 3.0, -0.015, 0.98, -2.0, 1e-5, -2e-5, 1.0])
    den = Ht[6] * xy[:, 0] + Ht[7] * xy[:, 1] + 1.0
    mx = (Ht[0] * xy[:, 0] + Ht[1] * xy[:, 1] + Ht[2]) / den + r.normal(0, 0.4, n)
    my = (Ht[3] * xy[:, 0] + Ht[4] * xy[:, 1] + Ht[5]) / den + r.normal(0, 0.4, n)
    out = r.random(n) < 0.2
    mx[out] = r.uniform(0, 1000, out.sum())
    my[out] = r.uniform(0, 1000, out.sum())
    pts["coords2D"] = xy
    pts["match_xpos"], pts["match_ypos"] = mx.astype(np.float32), my.astype(np.float32)
    pts["score"] = r.uniform(0.0, 1.0, n).astype(np.float32)
    pts["ambiguity"] = r.uniform(0.3, 1.0, n).astype(np.float32)      # about 30 % fail the 0.80 filter
    H0 = (Ht + np.array([0.01, -0.004, 1.0, 0.003, 0.01, -1.0, 0, 0, 0])).astype(np.float32)
    H_o, nfit_o, pts_o = O.improve_homography(pts, H0, 5, 0.1, 0.80, 3.0)
    H_d, nfit_d, pts_d = gpu_ctx.improve_homography(pts, H0, 5, 0.1, 0.80, 3.0)
    assert abs(nfit_d - nfit_o) <= 1, (n, nfit_d, nfit_o)            # a point exactly on the limit may round either way
    assert np.allclose(H_d, H_o, rtol=1e-5, atol=1e-6), (n, H_d, H_o)
    assert np.allclose(pts_d["match_error"], pts_o["match_error"], rtol=1e-4, atol=1e-4)


def test_allpairs_with_improve_vs_oracle(gpu_ctx):
    """The batched pipeline with ImproveHomography appended to every pair (device IRLS, one 8-CTA cluster per pair)."""
    imgs = [csb.synth(640, 480, 3100 + i) for i in range(3)]
    p = csb.make_params(5, 0.0, 0.5)
    sets = []
    for im in imgs:
        k = gpu_ctx.extract(im, p, max_pts=8192)
        sets.append(np.ascontiguousarray(PU.canonical_sort(k)[:1280]))
    # make set 1 a warped copy of set 0 so that one pair has a real homography to refine
    sets[1] = sets[0].copy()
    x, y = sets[0]["coords2D"][:, 0].astype(np.float64), sets[0]["coords2D"][:, 1].astype(np.float64)
    den = 1e-5 * x - 2e-5 * y + 1.0
    sets[1]["coords2D"][:, 0] = ((1.02 * x + 0.01 * y + 3.0) / den).astype(np.float32)
    sets[1]["coords2D"][:, 1] = ((-0.015 * x + 0.98 * y - 2.0) / den).astype(np.float32)
    dptrs = [gpu_ctx.upload_sift(s) for s in sets]
    pairs = csb.all_pairs(len(sets))
    loops, seed = 256, 11
    try:
        H, inl, nv, H2, nf = gpu_ctx.allpairs(dptrs, [len(s) for s in sets], pairs, "l2", loops, 0.0, 0.80, 5.0, seed,
                                              improve_loops=5, improve_thresh=3.0)
        for k, (i, j) in enumerate(pairs):
            m = O.match(sets[i], sets[j], "l2")
            valid = O.valid_points(m, 0.0, 0.80)
            rp = gpu_ctx.sample_points(valid, loops, seed, k)
            Ho, cnto = O.find_homography(m, rp, 5.0)
            assert inl[k] == cnto and nv[k] == len(valid)
            H2o, nfo, _ = O.improve_homography(m, H[k], 5, 0.0, 0.80, 3.0)
            assert nf[k] == nfo, (k, nf[k], nfo)
            assert np.allclose(H2[k], H2o, rtol=1e-5, atol=1e-5), (k, H2[k], H2o)
        k01 = pairs.index((0, 1))
        assert nf[k01] > 1000 and abs(H2[k01][0] - 1.02) < 1e-3 and abs(H2[k01][2] - 3.0) < 5e-2, (nf[k01], H2[k01])
    finally:
        for d in dptrs:
            gpu_ctx.free(d)


@pytest.mark.parametrize("kind", ["scaled_x10", "signed", "signed_padded", "range_0_255", "huge", "norm_1p01"])
def test_match_descriptors_outside_the_fp16_domain_use_the_exact_kernel(gpu_ctx, kind):
    """MatchSiftData accepts ANY SiftPoint.data (extras/matching.cu:232-362 is plain fp32).  The tensor-core prefilter's
    error bound needs descriptors that are finite in fp16 with squared norm <= 1.002 - and non-negative when the set
    needs padding rows (their score 0 must not exceed a real score); k_pack_f16 checks that and the call falls back to
    the exact fp32 kernel otherwise.  Results must equal the oracle bit for bit either way."""
    n1, n2 = (700, 1536) if kind == "signed" else (700, 1500)
    a, b = _rand_set(n1, 501), _rand_set(n2, 502)
    r = np.random.default_rng(9)
    expect_fallback = True
    if kind == "scaled_x10":
        a["data"] *= 10.0
        b["data"] *= 10.0
    elif kind == "signed":                                    # unit norm, mixed signs, no padding rows: inside the domain
        b["data"] *= r.choice([-1.0, 1.0], b["data"].shape).astype(np.float32)
        expect_fallback = False
    elif kind == "signed_padded":                             # the same with 36 padding rows in the candidate set
        b["data"] *= r.choice([-1.0, 1.0], b["data"].shape).astype(np.float32)
    elif kind == "range_0_255":
        a["data"] = np.round(a["data"] * 512).clip(0, 255)
        b["data"] = np.round(b["data"] * 512).clip(0, 255)
    elif kind == "huge":
        b["data"][77, 5] = 1.0e6                              # > 65504: inf in fp16
    else:
        a["data"] *= np.float32(1.01)                         # |q|^2 = 1.02 > 1.002
    L = csb.lib()
    for dist in ("l2", "dot"):
        before = L.csb_match_domain_fallbacks(gpu_ctx.h)
        ours, orc = gpu_ctx.match(a, b, dist), O.match(a, b, dist)
        took = L.csb_match_domain_fallbacks(gpu_ctx.h) - before
        assert took == (1 if expect_fallback else 0), (kind, dist, took)
        for f in ("score", "ambiguity", "match", "match_xpos", "match_ypos"):
            assert np.array_equal(ours[f], orc[f]), (kind, dist, f)


def test_allpairs_out_of_domain_set_falls_back(gpu_ctx):
    sets = [_rand_set(512, 601), _rand_set(512, 602), _rand_set(512, 603)]
    sets[1]["data"] *= 3.0
    dptrs = [gpu_ctx.upload_sift(s) for s in sets]
    pairs = csb.all_pairs(3)
    try:
        before = csb.lib().csb_match_domain_fallbacks(gpu_ctx.h)
        gpu_ctx.allpairs(dptrs, [512] * 3, pairs, "l2", 64, 0.0, 0.80, 5.0, 3)
        assert csb.lib().csb_match_domain_fallbacks(gpu_ctx.h) - before == 2        # pairs (0,1) and (1,2)
        last = {}
        for (i, j) in pairs:
            last[i] = j
        for i, j in last.items():
            dev = gpu_ctx.download_sift(dptrs[i], 512)
            ref = O.match(sets[i], sets[j], "l2")
            for f in ("score", "ambiguity", "match"):
                assert np.array_equal(dev[f], ref[f]), (i, j, f)
    finally:
        for d in dptrs:
            gpu_ctx.free(d)


@pytest.mark.parametrize("pitch", [640, 648, 768, 1024])
def test_extract_with_noncanonical_source_pitch(gpu_ctx, frames, pitch):
    """cuImage::Allocate takes any pitch (cuImage.cu:16-45); only the SOURCE image uses it - the DoG stack and its TMA
    descriptors keep the slot's own pitch.  A pitch must still satisfy cudaCreateTextureObject (32-byte row alignment =
    multiple of 8 floats; the reference binds the same texture, cuSIFT.cu:218-236): anything else is rejected with the
    CUDA error as status (test below)."""
    img = frames[0]
    h, w = img.shape
    want = PU.canonical_sort(gpu_ctx.extract(img, csb.make_params(5, 0.0, 0.5), max_pts=32768))
    padded = np.full((h, pitch), -1.0e6, np.float32)            # poison between width and pitch
    padded[:, :w] = img
    d_img = gpu_ctx.alloc(padded.nbytes)
    d_sift = gpu_ctx.alloc(588 * 32768)
    pin = csb.PinnedArray(32768)
    try:
        gpu_ctx.h2d(d_img, padded)
        cnt = gpu_ctx.extract_batch([d_img], w, h, pitch, csb.make_params(5, 0.0, 0.5), [d_sift], [pin.ptr], 32768)
        got = PU.canonical_sort(pin.array[: cnt[0]].copy())
        assert len(got) == len(want) > 1000
        for f in ("coords2D", "scale", "sharpness", "edgeness", "orientation", "data", "subsampling"):
            assert np.array_equal(got[f], want[f]), (pitch, f)
    finally:
        gpu_ctx.free(d_img)
        gpu_ctx.free(d_sift)
        pin.free()


def test_unbindable_pitch_is_an_error_status(gpu_ctx, frames):
    """A source pitch the texture unit cannot bind (644 floats = 2576 bytes, not a multiple of 32) comes back as a CUDA
    status from csb_extract (the reference would print the CUDA error and exit, cutils.h:24-48); the context stays usable."""
    img = frames[0]
    h, w = img.shape
    d_img = gpu_ctx.alloc(644 * h * 4)
    d_sift = gpu_ctx.alloc(588 * 1024)
    try:
        n = ctypes.c_int(0)
        p = csb.make_params(5, 0.0, 0.5)
        rc = csb.lib().csb_extract(gpu_ctx.h, d_img, w, h, 644, ctypes.byref(p), d_sift, 1024, None, ctypes.byref(n))
        assert rc != 0 and b"cudaCreateTextureObject" in csb.lib().csb_last_error(gpu_ctx.h)
        assert len(gpu_ctx.extract(img, p, max_pts=32768)) > 1000
    finally:
        gpu_ctx.free(d_img)
        gpu_ctx.free(d_sift)


def test_allpairs_distributed_single_rank_equals_batched_call(gpu_ctx):
    """csb_allpairs_distributed with world = 1 (no NCCL needed): same pairs, same flattened order, same results as
    csb_allpairs_match_ransac_improve.  The multi-rank exchange itself is exercised by tools/allpairs_bench.py --check
    under torchrun (2+ GPUs): every rank ends with the results of a single-rank run."""
    sets = [_rand_set(600 + 10 * k, 700 + k) for k in range(4)]
    for s in sets:
        s["match_xpos"] = 0
    dptrs = [gpu_ctx.upload_sift(s) for s in sets]
    cnts = [len(s) for s in sets]
    pairs = csb.all_pairs(4)
    try:
        ref = gpu_ctx.allpairs(dptrs, cnts, pairs, "l2", 128, 0.0, 0.80, 5.0, 5, None, improve_loops=2, improve_thresh=3.0)
        for d, s in zip(dptrs, sets):
            gpu_ctx.h2d(d, s)                                         # the call above wrote match fields: restore
        out = gpu_ctx.allpairs_distributed(None, 0, 1, dptrs, cnts, 640, "l2", 128, 0.0, 0.80, 5.0, 5, 2, 3.0)
        assert np.array_equal(out["inliers"], ref[1]) and np.array_equal(out["n_valid"], ref[2])
        assert np.array_equal(out["H"], ref[0]) and np.array_equal(out["num_fit"], ref[4])
        assert np.allclose(out["H_improved"], ref[3], rtol=1e-6, atol=1e-6)
    finally:
        for d in dptrs:
            gpu_ctx.free(d)


@pytest.mark.parametrize("source", ["device", "host", "host_u8"])
def test_compact_results_equal_the_siftpoint_records(gpu_ctx, frames, source):
    """Opt-in compact result mode (288-byte records, fp16 descriptor): header fields identical to the SiftPoint records of
    the same frame, descriptor == the fp32 descriptor rounded to fp16, for all three frame sources."""
    imgs = [np.ascontiguousarray(g) for g in frames] + [csb.synth(640, 480, 77)]
    if source == "host_u8":
        imgs = [np.clip(np.rint(im), 0, 255).astype(np.uint8) for im in imgs]
    p = csb.make_params(5, 0.0, 0.5)
    ds = [gpu_ctx.alloc(588 * 16384) for _ in imgs]
    pins = [csb.PinnedArray(16384, csb.COMPACT_DTYPE) for _ in imgs]
    dev = []
    try:
        if source == "device":
            dev = [gpu_ctx.upload_image(im) for im in imgs]
            cnt = gpu_ctx.extract_batch_compact([d for d, _ in dev], 640, 480, dev[0][1], p, ds, [q.ptr for q in pins], 16384)
        elif source == "host":
            cnt = gpu_ctx.extract_batch_compact([im.ctypes.data for im in imgs], 640, 480, 640, p, ds, [q.ptr for q in pins],
                                                16384, source="host")
        else:
            cnt = gpu_ctx.extract_batch_compact([im.ctypes.data for im in imgs], 640, 480, 640, p, ds, [q.ptr for q in pins],
                                                16384, source="host_u8")
        for k, im in enumerate(imgs):
            full = gpu_ctx.download_sift(ds[k], int(cnt[k]))                 # the device array still holds SiftPoint records
            want = gpu_ctx.extract(im.astype(np.float32), p, max_pts=16384)
            assert len(full) == len(want) == cnt[k] > 1000
            c = pins[k].array[: cnt[k]]
            assert np.array_equal(c["x"], full["coords2D"][:, 0]) and np.array_equal(c["y"], full["coords2D"][:, 1])
            for f in ("scale", "orientation", "sharpness", "edgeness", "subsampling"):
                assert np.array_equal(c[f], full[f]), f
            assert np.array_equal(c["data"], full["data"].astype(np.float16))
            assert np.array_equal(np.sort(PU.kp_key(full), axis=0), np.sort(PU.kp_key(want), axis=0))
    finally:
        for d in ds:
            gpu_ctx.free(d)
        for d, _ in dev:
            gpu_ctx.free(d)
        for q in pins:
            q.free()
