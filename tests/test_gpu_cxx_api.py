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
