"""CPU: pins oracle/oracle.c against every golden vector the reference's own tests
hold for the hot path.  No GPU, no product code."""
import numpy as np
import pytest

import parity_utils as PU
from oracle import oracle as O


@pytest.fixture(scope="module")
def fixtures():
    s1 = O.read_vlfeat_sift(PU.GOLDEN / "sift1.bin")
    s2 = O.read_vlfeat_sift(PU.GOLDEN / "sift2.bin")
    return s1, s2


def test_fixture_shapes(fixtures):
    s1, s2 = fixtures
    assert len(s1) == 884 and len(s2) == 856
    n = np.linalg.norm(s1["data"], axis=1)
    assert n.min() > 0.9999 and n.max() < 1.0001          # VLFeat descriptors are unit length


def test_matching_nn_indices_match_matlab(fixtures):
    """test/test.cpp:30-40 — every MATLAB pair (i, j): match(i-1)+1 == j."""
    s1, s2 = fixtures
    m = O.match(s1, s2, "l2")
    i, j = O.read_match_indices(PU.GOLDEN / "match_indices1_2.bin")
    assert len(i) == 326
    assert int((m["match"][i - 1] + 1 == j).sum()) == 326


def test_matching_ratio_count(fixtures):
    """test/test.cpp:52-55 — MatchSiftData(.., L2, 1000, 0.6) yields exactly 340 matches."""
    s1, s2 = fixtures
    m = O.match(s1, s2, "l2")
    assert O.count_matches(m, 1000.0, 0.6) == 340
    assert O.count_matches(m) == 884                       # defaults keep every query


def test_matching_scores_are_consistent(fixtures):
    s1, s2 = fixtures
    m = O.match(s1, s2, "l2")
    dot = s1["data"].astype(np.float64) @ s2["data"].astype(np.float64).T
    l2 = 2 - 2 * dot
    assert np.array_equal(m["match"], l2.argmin(1))
    assert np.allclose(m["score"], l2.min(1), atol=2e-6)
    second = np.sort(l2, 1)[:, 1]
    big = l2.min(1) > 1e-3                                  # near-zero fp32 scores are rounding noise
    assert np.allclose(m["ambiguity"][big], (l2.min(1) / (second + 1e-6))[big], atol=1e-4)
    assert np.array_equal(m["match_xpos"], s2["coords2D"][m["match"], 0])
    md = O.match(s1, s2, "dot")
    assert np.array_equal(md["match"], dot.argmax(1))


def test_match_tie_break_rule():
    """FindMinCorr on exact ties (matching.cu:229-257): winner = argmin of
    (score, bitrev4(col % 16), col // 16); a duplicate lands in `second` -> ambiguity ~ 1."""
    rng = np.random.default_rng(3)
    base = np.abs(rng.standard_normal((40, 128))).astype(np.float32)
    base /= np.linalg.norm(base, axis=1, keepdims=True)
    s2 = np.zeros(40, O.SIFT_DTYPE)
    s2["data"] = base
    s2["data"][17] = s2["data"][3]          # exact duplicates of candidate 3 ...
    s2["data"][24] = s2["data"][3]          # ... in lanes 1 and 8
    s1 = np.zeros(1, O.SIFT_DTYPE)
    s1["data"][0] = s2["data"][3]
    m = O.match(s1, s2, "l2")
    # lanes: 3 -> 3, 17 -> 1, 24 -> 8 ; bit-reversed lane priority 0,8,4,12,2,10,6,14,1,9,... => lane 8 wins
    assert m["match"][0] == 24
    assert abs(m["ambiguity"][0] - m["score"][0] / (m["score"][0] + 1e-6)) < 1e-6


def test_detector_golden_cusift1_check():
    """test/detector.cpp:41-84 on test/data/color1.jpg (== gray1): every one of the 4096 rows of
    the (saturated) golden file must be one of the oracle's unsaturated keypoints."""
    g1, _ = PU.golden_frames()
    pts, n, mpb = O.extract(g1, 6, 0.0, 0.1, 10.0, 0.0, False, 32768)
    assert n == 9508                                        # SURVEY.md 8c prediction
    assert PU.per_octave_counts(pts) == {"1.0": 7953, "2.0": 1180, "4.0": 261, "8.0": 87, "16.0": 20, "32.0": 7}
    assert mpb < 32                                         # reference's 32-entry block list never wraps
    gold = O.read_cusift_golden(PU.GOLDEN / "cusift1_check.bin")
    assert gold.shape == (4096, 4)
    from scipy.spatial import cKDTree
    d, idx = cKDTree(PU.kp_key(pts)).query(gold[:, :3].astype(np.float64))
    assert d.max() < 1e-4, f"golden keypoint not reproduced: max dist {d.max()}"
    assert (d == 0).sum() > 2900                            # most are bit-identical
    # all coarse-octave keypoints are in the golden file (they are extracted first)
    coarse = pts[pts["subsampling"] > 1]
    assert len(coarse) == 1555
    assert set(np.nonzero(pts["subsampling"] > 1)[0]) <= set(idx.tolist())
    # orientation: with the texture filter restated from hardware measurements (oracle.c tex2d) every
    # golden row agrees within the reference's own run-to-run spread (float atomics, ~6e-5 deg)
    do = PU.ang_diff_deg(pts["orientation"][idx], gold[:, 3])
    assert np.median(do) < 1e-4
    assert do.max() < PU.ORI_TOL_DEG, do.max()


def test_scale_down_constants_and_shape():
    """cuSIFT.cu:330-338 weights and the fork's asymmetric column filter (cuSIFT_D.cu:123-177)."""
    k = np.array([np.exp(-4.0), np.exp(-1.0), 1.0], np.float64)
    k = k / (2 * k[0] + 2 * k[1] + k[2])
    img = np.zeros((32, 48), np.float32)
    img[10, 20] = 1.0
    out = O.scale_down(img)
    assert out.shape == (16, 24)
    # impulse at row 10 = 2j -> j=5 with k2; = 2j+2 -> j=4 with k0; = 2j+3 never (odd); = 2j+1 / 2j-1 never (even row)
    col = out[:, 10]
    assert abs(col[5] - k[2] * k[2]) < 1e-7 and abs(col[4] - k[0] * k[2]) < 1e-7 and col[6] == 0.0
    img[:] = 0
    img[11, 20] = 1.0                                      # odd row: 2j+1 (j=5), 2j-1 (j=6), 2j+3 (j=4)
    col = O.scale_down(img)[:, 10]
    assert abs(col[5] - k[1] * k[2]) < 1e-7 and abs(col[6] - k[1] * k[2]) < 1e-7 and abs(col[4] - k[0] * k[2]) < 1e-7
    flat = np.full((37, 51), 7.0, np.float32)
    assert np.allclose(O.scale_down(flat), 7.0, atol=1e-5)


def test_blur_schedule_and_weights():
    assert np.allclose(O.init_blurs(6), [0, 0.25, 0.279509, 0.286411, 0.288111, 0.288534], atol=1e-6)   # cuSIFT.cu:188
    k = O.laplace_weights(0.0)
    assert k.shape == (8, 9)
    assert np.allclose(k.sum(1), 1.0, atol=1e-6) and np.allclose(k, k[:, ::-1])
    sig = 2.0 ** ((np.arange(8) - 1) / 5.0)                # sigma_i = 2^((i-1)/5)
    ref = np.exp(-np.arange(-4, 5)[None, :] ** 2 / (2 * sig[:, None] ** 2))
    ref /= ref.sum(1, keepdims=True)
    assert np.allclose(k, ref, atol=1e-6)


def test_dog_of_flat_and_border_rule():
    flat = np.full((40, 60), 100.0, np.float32)
    d = O.dog(flat, 0.0)
    assert np.abs(d).max() < 1e-4
    pts, n, _ = O.find_points(d, 0.01, 10.0, 1.0)
    assert n == 0
    rng = np.random.default_rng(0)
    img = (rng.random((64, 96)) * 255).astype(np.float32)
    pts, n, _ = O.find_points(O.dog(img, 0.0), 0.1, 10.0, 1.0)
    x, y = np.round(pts["coords2D"][:, 0]), np.round(pts["coords2D"][:, 1])
    assert n > 0 and x.min() >= 0.5 and y.min() >= 0.5     # image-border pixels never win (clamped neighbours)


def test_rootsift_definition():
    rng = np.random.default_rng(1)
    pts = np.zeros(5, O.SIFT_DTYPE)
    d = np.abs(rng.standard_normal((5, 128))).astype(np.float32)
    pts["data"] = d / np.linalg.norm(d, axis=1, keepdims=True)
    r = O.rootsift(pts)["data"]
    assert np.allclose(r, np.sqrt(pts["data"] / pts["data"].sum(1, keepdims=True)), atol=1e-6)
    assert np.allclose(np.linalg.norm(r, axis=1), 1.0, atol=1e-5)


def test_homography_recovers_known_transform():
    rng = np.random.default_rng(2)
    n = 200
    H = np.array([1.02, 0.01, 3.0, -0.02, 0.98, -2.0, 1e-5, -2e-5, 1.0])
    pts = np.zeros(n, O.SIFT_DTYPE)
    xy = rng.uniform(0, 600, (n, 2)).astype(np.float32)
    den = H[6] * xy[:, 0] + H[7] * xy[:, 1] + 1
    pts["coords2D"] = xy
    pts["match_xpos"] = (H[0] * xy[:, 0] + H[1] * xy[:, 1] + H[2]) / den
    pts["match_ypos"] = (H[3] * xy[:, 0] + H[4] * xy[:, 1] + H[5]) / den
    out = rng.choice(n, 60, replace=False)                 # 30 % outliers
    pts["match_xpos"][out] += rng.uniform(20, 80, 60).astype(np.float32)
    pts["score"] = 0.5
    pts["ambiguity"] = 0.5
    valid = O.valid_points(pts, 0.0, 0.8)
    assert len(valid) == n
    loops = 256
    rp = np.stack([valid[rng.choice(n, 4, replace=False)] for _ in range(loops)], 1).astype(np.int32)
    Hf, cnt = O.find_homography(pts, rp, 2.0)
    assert 138 <= cnt <= 140 + 8                            # 140 inliers (+ at most the 8 zero pad slots, n_up = 208)
    H2, nfit, _ = O.improve_homography(pts, Hf, 5, 0.0, 0.8, 2.0)
    assert nfit >= 138
    assert np.allclose(H2[:6], H[:6], atol=5e-2) and np.allclose(H2[6:8], H[6:8], atol=1e-4)


def test_rigid_transform_oracle_reproduces_matlab_rt():
    """test/test.cpp:58-110 (RigidTransform.RANSACWithIndices) prints its result next to the MATLAB Rt stored in
    test/data/RigidTransform_RANSAC.bin; the numpy restatement of EstimateRigidTransformH reproduces that Rt."""
    coord, idx, Rt = O.read_matlab_ransac(PU.GOLDEN / "rigid_ransac.bin")
    assert coord.shape == (120, 6) and idx.shape == (10, 3) and idx.min() >= 0 and idx.max() < 120
    got, n_inl, mask = O.rigid_transform(coord, idx, 0.05 * 0.05, True)
    assert n_inl == 114 and mask.sum() == 114
    assert np.abs(got - Rt).max() < 1e-6
    R = got.reshape(3, 4)[:, :3]
    assert np.abs(R @ R.T - np.eye(3)).max() < 1e-6 and abs(np.linalg.det(R) - 1) < 1e-6


def test_rigid_transform_oracle_planar_and_3d_recover_known_motion():
    """estimateRigidTransform2D / 3D restatements on synthetic correspondences with 25 % outliers."""
    r = np.random.default_rng(4)
    mov = r.uniform(-1, 1, (400, 3)) + np.array([0, 0, 2.0])
    a = 0.35
    R = np.array([[np.cos(a), 0, -np.sin(a)], [0, 1, 0], [np.sin(a), 0, np.cos(a)]])      # rotation about y (planar x-z motion)
    t = np.array([0.2, 0.0, -0.1])
    ref = mov @ R.T + t
    ref[:100] = r.uniform(-2, 2, (100, 3))
    coord = np.ascontiguousarray(np.concatenate([ref, mov], 1), np.float32)
    idx = np.stack([r.choice(400, 3, replace=False) for _ in range(64)]).astype(np.int32)
    for type3d in (False, True):
        Rt, n_inl, mask = O.rigid_transform(coord, idx, 1e-3 ** 2, type3d)
        assert n_inl >= 295 and mask[100:].mean() > 0.98 and mask[:100].sum() <= 3
        assert np.abs(Rt.reshape(3, 4)[:, :3] - R).max() < 1e-4 and np.abs(Rt.reshape(3, 4)[:, 3] - t).max() < 1e-4
