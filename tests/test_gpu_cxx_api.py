"""GPU: the reference-style C++ API (include/cusift/*.h) exercised by a C++ consumer written the
way the reference's main.cpp demo and test/detector.cpp use it (tests/cpp/csb_demo.cpp)."""
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

import parity_utils as PU
from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
DEMO = ROOT / "build" / "csb_demo"


@pytest.mark.skipif(not DEMO.exists(), reason="build/csb_demo not built (make demo)")
def test_main_cpp_style_pipeline(workdir):
    g1, g2 = PU.golden_frames()
    a, b = PU.preblur(g1), PU.preblur(g2)                   # main.cpp:308-309
    pa, pb = workdir / "l.f32", workdir / "r.f32"
    a.tofile(pa)
    b.tofile(pb)
    out = subprocess.run([str(DEMO), str(pa), str(pb), "640", "480", "0"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    r = json.loads(out.stdout.strip().splitlines()[-1])
    # maxPts = 4096 saturates exactly like the reference's demo (main.cpp:324)
    assert r["numPts1"] == 4096 and r["numPts2"] == 4096
    assert r["matches_all"] == 4096 and 0 < r["matches_ratio"] < 4096 and r["ptrs_ok"] == 1
    assert r["numMatches"] > 500 and r["numFit"] > 300
    H = np.array(r["H"])
    assert abs(H[0] - 1) < 0.05 and abs(H[4] - 1) < 0.05 and H[8] == 1.0
    # HEAD-generation SiftData::Extract on the pre-blurred frame: emulator prediction 10 837
    orc_n = O.extract(a, 6, 0.0, 0.1, 10.0, 0.0, False, 32768)[1]
    assert r["head_pts"] == orc_n
    assert abs(r["desc_norm2"] - 1.0) < 1e-3 and abs(r["rootsift_norm2"] - 1.0) < 1e-3
    assert abs(r["half00"] - float(O.scale_down(a)[0, 0])) < 1e-6


REFTESTS = ROOT / "build" / "csb_ref_tests"


@pytest.mark.skipif(not REFTESTS.exists(), reason="build/csb_ref_tests not built (make demo)")
def test_reference_test_suite_restated_in_cpp(workdir):
    """test/test.cpp (Matching.*) and test/detector.cpp (Detector.DetectorCUSIFTTest) of the reference,
    restated in C++ on the drop-in headers + extras/debug.h readers, against the reference's goldens."""
    g1, _ = PU.golden_frames()
    raw = workdir / "gray1.f32"
    g1.astype(np.float32).tofile(raw)
    out = subprocess.run([str(REFTESTS), str(PU.GOLDEN), str(raw)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, (out.stdout[-500:], out.stderr[-3000:])
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert r["nn_checked"] == 326 and r["nn_equal"] == 326      # test.cpp:37-40
    assert r["ratio_matches"] == 340                            # test.cpp:55
    assert r["det_pts"] == 4096 and r["det_rows"] == 4096       # detector.cpp:68
    assert r["det_found"] == 4096 and r["det_found_abs"] == 4096
    assert r["rt_inliers"] == 114 and r["rt_maxdiff"] < 1e-5     # test.cpp:58-110 vs the MATLAB Rt in the file
    assert r["rt_inliers_random"] >= 114 and r["rt_maxdiff_random"] < 2e-2
    assert r["failed"] == 0


REF_MAIN = ROOT / "build" / "ref_main_demo"
REF_SUITE = ROOT / "build" / "ref_test_suite"


def _write_pgm(path, img_u8):
    with open(path, "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (img_u8.shape[1], img_u8.shape[0]))
        f.write(np.ascontiguousarray(img_u8, np.uint8).tobytes())


@pytest.mark.skipif(not REF_MAIN.exists(), reason="build/ref_main_demo not built (make refharness, needs /root/reference)")
def test_reference_main_cpp_unchanged(gpu_ctx, workdir):
    """The reference's OWN main.cpp (demo(): imread -> GaussianBlur -> cuImage -> InitSiftData -> ExtractSift x2 ->
    MatchSiftData -> FindHomography -> ImproveHomography -> PrintMatchData -> imwrite, main.cpp:285-350), compiled
    unchanged against include/cusift/ + tests/compat/ and linked with libcusift_b200.so, run on its own image pair
    (as PGM: the compat imread has no JPEG decoder).  Its printed counts must equal the same pipeline through the C ABI."""
    import re
    import cusift_b200 as csb
    g1, g2 = PU.golden_frames()
    l, r, o = workdir / "left.pgm", workdir / "right.pgm", workdir / "out.pgm"
    _write_pgm(l, g1.astype(np.uint8))
    _write_pgm(r, g2.astype(np.uint8))
    out = subprocess.run([str(REF_MAIN), str(l), str(r), str(o), "0"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    txt = out.stdout
    assert "Image size = (640,480)" in txt
    m1 = re.search(r"Number of original features: (\d+) (\d+)", txt)
    m2 = re.search(r"Number of matching features: (\d+) (\d+)", txt)
    assert m1 and m2, txt[-1000:]
    assert (int(m1.group(1)), int(m1.group(2))) == (4096, 4096)          # main.cpp:324: maxPts 4096 saturates
    num_fit, num_matches = int(m2.group(1)), int(m2.group(2))
    assert num_matches > 500 and num_fit > 300
    assert o.exists() and o.stat().st_size > 640 * 480                   # imwrite of the annotated left image
    # the 3x3 pre-blur of the compat GaussianBlur equals OpenCV's (and therefore the frames the parity tests use)
    a = PU.preblur(g1)
    got = gpu_ctx.extract(a, csb.make_params(6, 0.0, 0.1), max_pts=4096)
    assert len(got) == 4096


@pytest.mark.skipif(not REF_SUITE.exists(), reason="build/ref_test_suite not built (make refharness, needs /root/reference)")
def test_reference_test_cpp_unchanged(workdir):
    """The reference's OWN test/test.cpp (Matching.MatchingTest: 326 MATLAB nearest neighbours, Matching.MatchingRatioTest:
    340 matches, RigidTransform.*), compiled unchanged with the gtest / opencv2 / vl compatibility headers and run from a
    build/ directory next to a test/data/ tree holding the reference's fixtures (as the reference's ctest does)."""
    import shutil
    root = workdir / "reftree"
    (root / "build").mkdir(parents=True, exist_ok=True)
    data = root / "test" / "data"
    for sub in ("sift", "match_indices", "match"):
        (data / sub).mkdir(parents=True, exist_ok=True)
    shutil.copyfile(PU.GOLDEN / "sift1.bin", data / "sift" / "sift1")
    shutil.copyfile(PU.GOLDEN / "sift2.bin", data / "sift" / "sift2")
    shutil.copyfile(PU.GOLDEN / "match_indices1_2.bin", data / "match_indices" / "match_indices1_2")
    shutil.copyfile(PU.GOLDEN / "match1_2.bin", data / "match" / "match1_2")
    shutil.copyfile(PU.GOLDEN / "rigid_ransac.bin", data / "RigidTransform_RANSAC.bin")
    out = subprocess.run([str(REF_SUITE)], cwd=root / "build", capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-1500:])
    assert "[  PASSED  ] 6 tests." in out.stdout, out.stdout[-1500:]
    assert "FAILED" not in out.stdout
    assert "now have 340" in out.stderr                                  # test.cpp:54 (and EXPECT_EQ(340, ...) passed)
    assert "Inliers / total: 114 / 120" in out.stderr                    # test.cpp:94 with the golden indices
