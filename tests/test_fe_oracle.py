"""CPU tests (no GPU): the oracle is pinned against cv2 4.13.0 and against the committed fixtures.

oracle/fe_cv_restate.c restates OpenCV's pyrDown / calcOpticalFlowPyrLK / goodFeaturesToTrack (the
arithmetic behind reference feature_tracker.cpp:118-153,198, which lives in un-vendored OpenCV); the
CUDA kernels follow that restatement, so it must be bit-equal to cv2 itself.
"""
import os

import numpy as np
import pytest

from conftest import make_texture_image, warp_image

cv2 = pytest.importorskip("cv2")
from oracle import cv_restate as R  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_cv2_version_is_the_pinned_one():
    assert cv2.__version__.startswith("4.13"), "oracle pinned to opencv 4.13.x"


@pytest.mark.parametrize("shape", [(640, 480), (321, 243), (1280, 720)])
def test_restated_pyr_down(shape):
    img = make_texture_image(1, *shape)
    for _ in range(3):
        assert np.array_equal(R.pyr_down(img), cv2.pyrDown(img))
        img = cv2.pyrDown(img)


@pytest.mark.parametrize("seed,contrast,sigma", [(0, 1.0, 2.0), (1, 2.5, 1.5), (2, 0.3, 3.0)])
def test_restated_min_eig_bit_exact(seed, contrast, sigma):
    img = make_texture_image(seed, contrast=contrast, sigma=sigma)
    assert np.array_equal(bits(R.min_eig(img)), bits(cv2.cornerMinEigenVal(img, 3, ksize=3)))


def test_naive_box_sum_is_not_what_cv2_does():
    """Documents why the column sum has to be replayed as a running double sum."""
    tot = 0
    for seed in range(3):
        img = make_texture_image(seed, contrast=1.0 + seed)
        tot += int((bits(R.min_eig(img, 1)) != bits(cv2.cornerMinEigenVal(img, 3, ksize=3))).sum())
    assert tot > 0


@pytest.mark.parametrize("max_level", [3, 1, 0])
def test_restated_lk_bit_exact(max_level):
    for seed, contrast in ((0, 1.0), (1, 2.5)):
        a = make_texture_image(seed, contrast=contrast)
        b = warp_image(a, 3.3 + seed, -2.1, 0.7)
        rng = np.random.default_rng(100 + seed)
        pts = np.stack([rng.uniform(-5, 645, 400), rng.uniform(-5, 485, 400)], 1).astype(np.float32)
        q_cv, st_cv, _ = cv2.calcOpticalFlowPyrLK(a, b, pts.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=max_level)
        q, st = R.lk(a, b, pts, max_level)
        assert np.array_equal(st, st_cv.ravel())
        assert np.array_equal(bits(q), bits(q_cv.reshape(-1, 2)))


def test_lk_accumulation_order_matters():
    """Sequential float accumulation is NOT cv2's order; the SSE lane order is."""
    a = make_texture_image(1, contrast=2.5)
    b = warp_image(a, 4.3, -2.1, 0.7)
    rng = np.random.default_rng(5)
    pts = np.stack([rng.uniform(0, 640, 400), rng.uniform(0, 480, 400)], 1).astype(np.float32)
    q_cv, _, _ = cv2.calcOpticalFlowPyrLK(a, b, pts.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3)
    q_seq, _ = R.lk(a, b, pts, 3, order=R.ORDER_SEQUENTIAL)
    assert (bits(q_seq) != bits(q_cv.reshape(-1, 2))).any()


def test_restated_lk_initial_flow():
    a = make_texture_image(5, contrast=1.5)
    b = warp_image(a, -2.6, 1.4, -0.4)
    rng = np.random.default_rng(7)
    pts = np.stack([rng.uniform(0, 640, 300), rng.uniform(0, 480, 300)], 1).astype(np.float32)
    fwd, _ = R.lk(a, b, pts, 3)
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01)
    r_cv, st_cv, _ = cv2.calcOpticalFlowPyrLK(b, a, fwd.reshape(-1, 1, 2), pts.reshape(-1, 1, 2).copy(), winSize=(21, 21),
                                              maxLevel=1, criteria=crit, flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
    r, st = R.lk(b, a, fwd, 1, init=pts)
    assert np.array_equal(st, st_cv.ravel()) and np.array_equal(bits(r), bits(r_cv.reshape(-1, 2)))


@pytest.mark.parametrize("max_corners,min_dist,ndisks", [(150, 30, 0), (40, 30, 60), (500, 15, 120), (1000, 7, 0)])
def test_restated_gftt(max_corners, min_dist, ndisks):
    for seed in range(2):
        img = make_texture_image(seed, contrast=1.0 + 0.5 * seed)
        rng = np.random.default_rng(seed)
        mask = np.full(img.shape, 255, np.uint8)
        for _ in range(ndisks):
            cv2.circle(mask, (int(rng.integers(0, 640)), int(rng.integers(0, 480))), min_dist, 0, -1)
        want = cv2.goodFeaturesToTrack(img, max_corners, 0.01, min_dist, mask=mask)
        want = np.zeros((0, 2), np.float32) if want is None else want.reshape(-1, 2)
        got = R.gftt(img, max_corners, 0.01, min_dist, mask)
        assert got.shape == want.shape and np.array_equal(got, want)


def test_circle_is_integer_disk():
    """cv::circle(mask, p, r, 0, -1) == {d^2 <= r^2}: the GPU never rasterises with OpenCV."""
    rng = np.random.default_rng(0)
    for r in (1, 7, 15, 30, 40):
        for _ in range(10):
            cx, cy = int(rng.integers(-5, 70)), int(rng.integers(-5, 70))
            m = np.full((64, 64), 255, np.uint8)
            cv2.circle(m, (cx, cy), r, 0, -1)
            ys, xs = np.mgrid[0:64, 0:64]
            assert np.array_equal(m == 0, (xs - cx) ** 2 + (ys - cy) ** 2 <= r * r)


@pytest.mark.parametrize("name", ["fe_c2_seed0", "fe_c3_seed1"])
def test_oracle_reproduces_golden(name):
    import hashlib
    from ground_fusion_b200.synth import SyntheticStream
    from oracle.fe_oracle import IDC_CAM, FeatureTrackerOracle, PinholeCamera
    g = np.load(os.path.join(GOLD, name + ".npz"))
    seed, w, h, max_cnt, min_dist, frames = [int(v) for v in g["meta"]]
    frames = min(frames, 12)   # keep the CPU suite short; the GPU test replays the whole fixture
    cam = PinholeCamera(**IDC_CAM)
    st = SyntheticStream(seed=seed, width=w, height=h)
    ft = FeatureTrackerOracle(cam, max_cnt, min_dist, 1, 1)
    for k in range(frames):
        t, gray, depth = st.frame(k)
        ff = ft.trackImage(t, gray, depth)
        assert np.array_equal(np.array(ft.ids, np.int32), g["ids_%d" % k])
        assert np.array_equal(ft.last_status, g["status_%d" % k])
        assert np.array_equal(ft.last_n_pts, g["npts_%d" % k])
        assert np.array_equal(np.array([ff[i] for i in ft.ids]).reshape(-1, 8), g["obs_%d" % k])


def test_oracle_first_frame_properties():
    from ground_fusion_b200.synth import SyntheticStream
    from oracle.fe_oracle import IDC_CAM, FeatureTrackerOracle, PinholeCamera
    ft = FeatureTrackerOracle(PinholeCamera(**IDC_CAM), 150, 30, 1, 1)
    t, gray, depth = SyntheticStream(seed=3).frame(0)
    ff = ft.trackImage(t, gray, depth)
    assert sorted(ff) == list(range(150)) and ft.n_id == 150
    p = np.array([ff[i][3:5] for i in range(150)])
    d2 = ((p[:, None] - p[None]) ** 2).sum(-1) + np.eye(150) * 1e9
    assert d2.min() >= 30 * 30              # min-distance property of the detector
    assert all(v[5] == 0 and v[6] == 0 for v in ff.values())   # no velocity on the first frame


def test_fast_oracle_equals_loop_oracle():
    """bench.py's CPU arm times FeatureTrackerOracleFast (glue vectorised / in C); it must be the same function as the
    line-by-line loop restatement: ids, inlier masks, new corners, observation vectors, incl. setPrediction /
    removeOutliers feedback and frames without a depth image."""
    from ground_fusion_b200.synth import SyntheticStream
    from oracle.fe_oracle import IDC_CAM, FeatureTrackerOracle, FeatureTrackerOracleFast, PinholeCamera
    cam = PinholeCamera(**IDC_CAM)
    st = SyntheticStream(seed=3)
    a = FeatureTrackerOracle(cam, 150, 30, 1, 1)
    b = FeatureTrackerOracleFast(cam, 150, 30, 1, 1)
    rng = np.random.default_rng(0)
    for k in range(14):
        t, gray, depth = st.frame(k)
        if k == 9:
            depth = None
        fa, fb = a.trackImage(t, gray, depth), b.trackImage(t, gray, depth)
        assert list(a.ids) == list(b.ids) and list(a.track_cnt) == list(b.track_cnt), "frame %d" % k
        assert np.array_equal(a.last_status, b.last_status), "frame %d" % k
        assert np.array_equal(a.last_n_pts, b.last_n_pts), "frame %d" % k
        assert sorted(fa) == sorted(fb)
        for fid in fa:
            assert np.array_equal(fa[fid], fb[fid]), "frame %d id %d" % (k, fid)
        if k in (4, 7) and fa:
            pred = {}
            for fid in list(fa)[::2]:
                d = max(fa[fid][7], 0.5)
                pred[fid] = (fa[fid][0] * d + rng.normal(0, 0.002), fa[fid][1] * d + rng.normal(0, 0.002), d)
            rm = set(list(fa)[1::9])                      # estimator.cpp:1132-1136: removeOutliers, then setPrediction
            a.removeOutliers(rm); b.removeOutliers(rm)
            a.setPrediction(pred); b.setPrediction(pred)
    assert b.t_cv > 0
