"""CPU: the C-ABI library loads and exports every symbol include/cusift_b200.h declares,
the layouts match the reference's, and nothing computes without a GPU (no fallback)."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

import cusift_b200 as csb
from cusift_b200._lib import HEADER_PATH, LIB_PATH, SIGNATURES, CsbParams

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = HEADER_PATH.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(csb_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built_in_tree():
    assert LIB_PATH.exists(), "run `make lib` (or __graft_entry__.build())"
    assert LIB_PATH.parent == ROOT / "cusift_b200"


def test_every_declared_symbol_is_exported_and_bound():
    syms = declared_symbols()
    assert len(syms) >= 25
    L = csb.lib()
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/cusift_b200.h but not exported"
        assert s in SIGNATURES, f"{s} has no ctypes signature"
    assert set(SIGNATURES) == set(syms)


def test_layouts_match_reference():
    L = csb.lib()
    assert L.csb_sizeof_sift_point() == 588                 # cuSIFT.h:10-30
    assert csb.SIFT_DTYPE.itemsize == 588
    assert csb.SIFT_DTYPE.fields["data"][1] == 64 and csb.SIFT_DTYPE.fields["coords3D"][1] == 576
    assert csb.SIFT_DTYPE.fields["match"][1] == 32 and csb.SIFT_DTYPE.fields["subsampling"][1] == 48
    assert C.sizeof(CsbParams) == 40
    assert L.csb_version() == 100


def test_cuda_code_targets_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", str(LIB_PATH)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_reference_cxx_api_is_exported():
    """The union of both API generations must be linkable."""
    out = subprocess.run(["nm", "-DC", "--defined-only", str(LIB_PATH)], capture_output=True, text=True).stdout
    for sym in ("InitSiftData(SiftData&, int, bool, bool)", "FreeSiftData(SiftData&)",
                "ExtractSift(SiftData&, cuImage&, int, double, float, float, float)",
                "ExtractRootSift(SiftData&, cuImage&, int, double, float, float, float)",
                "SiftData::Extract(float*, int, int, float)", "SiftData::SiftData(int, bool, bool)",
                "SiftData::Synchronize()", "SiftData::ConvertSiftToRootSift()",
                "cuImage::Allocate(int, int, int, bool, float*, float*)", "cuImage::HostToDevice()",
                "ScaleDown(cuImage&, cuImage&, float)",
                "MatchSiftData(SiftData&, SiftData&, MatchSiftDistance, float, float, MatchType)",
                "FindHomography(SiftData&, float*, int*, int, float, float, float)",
                "ImproveHomography(SiftData&, float*, int, float, float, float)",
                # next rows: extras/rigidTransform.h and the OpenCV-free part of extras/debug.h
                "EstimateRigidTransformH(float const*, float*, int*, int, int, float, RigidTransformType, int*, char*)",
                "ReadVLFeatSiftData(SiftData&, char const*)", "ReadMATLABMatchIndices(char const*, unsigned int*, unsigned int*)",
                "ReadMATLABRANSAC(char const*, std::vector<int, std::allocator<int> >&, float*)",
                "AddSiftData(SiftData&, SiftPoint*, int)", "PrintSiftData(SiftData&)"):
        assert sym in out, sym


def test_no_cpu_fallback():
    """Without a GPU a context cannot be created; with one this test is vacuous."""
    try:
        ctx = csb.Context(0, 1)
    except csb.CsbError as e:
        assert "GPU" in str(e)
        return
    ctx.close()


def test_product_never_touches_the_oracle():
    for p in list((ROOT / "cusift_b200").rglob("*.py")) + list((ROOT / "cusift_b200" / "csrc").glob("*")) + \
            list((ROOT / "include").rglob("*.h")):
        txt = p.read_text(errors="ignore")
        assert "oracle" not in txt.lower(), f"{p} mentions the oracle"


def test_synth_is_deterministic_and_in_range():
    a = csb.synth(320, 200, 5)
    b = csb.synth(320, 200, 5)
    assert a.dtype == np.float32 and a.shape == (200, 320)
    assert np.array_equal(a, b) and a.min() >= 0 and a.max() <= 255
    assert not np.array_equal(a, csb.synth(320, 200, 6))
    s = csb.synth(1920, 1080, 1000)                         # SURVEY.md Appendix B statistics
    assert s.min() == 0.0 and s.max() == 255.0 and abs(float(s.mean()) - 127.905) < 0.01


def test_param_struct_roundtrip():
    p = csb.make_params(5, 0.25, 1.0, 10.0, 0.0, 2.0, True)
    assert (p.num_octaves, p.init_blur, p.peak_thresh, p.edge_thresh, p.subsampling, p.rootsift) == (5, 0.25, 1.0, 10.0, 2.0, 1)
