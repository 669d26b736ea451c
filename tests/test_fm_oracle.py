"""CPU tests of the FeatureManager restatement (oracle/fm_oracle.py; reference vins_estimator/src/estimator/feature_manager.cpp)
and of the replay loop (ground_fusion_b200/replay.py) driven by the CPU oracles.  The reference has no tests for this class:
the restatement is pinned by geometric ground truth (a landmark seen from known poses must come back with its true depth)."""
import numpy as np
import pytest

from oracle.fm_oracle import INIT_DEPTH, FeatureManagerOracle


def scene(seed=0, n_frames=11, n_lm=40, noise=0.0, depth_noise=0.0):
    """Camera poses along a gentle arc (body = camera: ric = I, tic = 0) and landmarks seen in runs of consecutive frames."""
    rng = np.random.default_rng(seed)
    Ps = np.array([[0.12 * k, 0.03 * np.sin(0.5 * k), 0.02 * k] for k in range(n_frames)])
    Rs = []
    for k in range(n_frames):
        a = 0.02 * k
        Rs.append(np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]))
    Rs = np.array(Rs)
    lms = np.stack([rng.uniform(-1.5, 2.5, n_lm), rng.uniform(-1, 1, n_lm), rng.uniform(2.0, 7.0, n_lm)], 1)
    frames = {}
    for l in range(n_lm):
        s0 = int(rng.integers(0, n_frames - 4)); m = int(rng.integers(2, n_frames - s0 + 1))
        for f in range(s0, s0 + m):
            pc = Rs[f].T @ (lms[l] - Ps[f])
            v8 = np.array([pc[0] / pc[2] + rng.normal(0, noise), pc[1] / pc[2] + rng.normal(0, noise), 1.0, 0, 0, 0, 0, pc[2] + rng.normal(0, depth_noise)])
            frames.setdefault(f, {})[l] = v8
    return Ps, Rs, lms, frames


def fill(fm, frames, n_frames=11):
    for f in range(n_frames):
        fm.addFeatureCheckParallax(f, frames.get(f, {}), 0.0)


def test_triangulate_recovers_the_true_depth_without_depth_measurements():
    Ps, Rs, lms, frames = scene(seed=1)
    fm = FeatureManagerOracle(depth_threshold=0.05)          # every RGB-D depth is beyond the threshold: DLT only
    fill(fm, frames)
    fm.triangulateWithDepth(10, Ps, Rs, np.zeros(3), np.eye(3))
    assert all(it.estimated_depth < 0 for it in fm.feature)
    fm.triangulate(10, Ps, Rs, np.zeros(3), np.eye(3))
    n = 0
    for it in fm.feature:
        if len(it.feature_per_frame) < 4:
            assert it.estimated_depth < 0
            continue
        true = (Rs[it.start_frame].T @ (lms[it.feature_id] - Ps[it.start_frame]))[2]
        assert it.estimate_flag == 2 and abs(it.estimated_depth - true) < 1e-9 * max(1, true)
        n += 1
    assert n > 10


def test_triangulate_with_depth_averages_the_verified_depths():
    Ps, Rs, lms, frames = scene(seed=2, depth_noise=0.01)
    fm = FeatureManagerOracle(depth_threshold=5.0)
    fill(fm, frames)
    fm.triangulateAll(10, Ps, Rs, np.zeros(3), np.eye(3))
    got1 = got2 = 0
    for it in fm.feature:
        if len(it.feature_per_frame) < 4:
            continue
        true = (Rs[it.start_frame].T @ (lms[it.feature_id] - Ps[it.start_frame]))[2]
        assert it.estimated_depth > 0
        if it.estimate_flag == 1:
            assert abs(it.estimated_depth - true) < 0.05; got1 += 1        # mean of noisy depths transferred to the start frame
        else:
            assert it.estimate_flag == 2 and abs(it.estimated_depth - true) < 1e-6; got2 += 1   # every depth beyond 5 m: DLT
    assert got1 > 5 and got2 > 0


def test_depth_vector_round_trip_and_failures():
    Ps, Rs, lms, frames = scene(seed=3)
    fm = FeatureManagerOracle(depth_threshold=10.0)
    fill(fm, frames)
    fm.triangulateAll(10, Ps, Rs, np.zeros(3), np.eye(3))
    x = fm.getDepthVector()
    assert len(x) == fm.getFeatureCount() == sum(1 for _ in fm.iter_ba_features())
    x2 = x.copy(); x2[0] = -0.5
    fm.setDepth(x2)
    n0 = len(fm.feature)
    fm.removeFailures()
    assert len(fm.feature) == n0 - 1


def test_window_shifts_keep_the_observation_lists_consistent():
    Ps, Rs, lms, frames = scene(seed=4)
    fm = FeatureManagerOracle(depth_threshold=10.0)
    fill(fm, frames)
    fm.triangulateAll(10, Ps, Rs, np.zeros(3), np.eye(3))
    before = {it.feature_id: (it.start_frame, len(it.feature_per_frame), it.estimated_depth) for it in fm.feature}
    fm.removeBackShiftDepth(Rs[0], Ps[0], Rs[1], Ps[1])
    for it in fm.feature:
        s0, m, d = before[it.feature_id]
        if s0 == 0:
            assert it.start_frame == 0 and len(it.feature_per_frame) == m - 1 and m - 1 >= 2
            if d > 0 and m >= 4:
                true = (Rs[1].T @ (lms[it.feature_id] - Ps[1]))[2]
                assert abs(it.estimated_depth - true) < 1e-6
        else:
            assert it.start_frame == s0 - 1 and len(it.feature_per_frame) == m
    assert all(not (s0 == 0 and m - 1 < 2) or fid not in {it.feature_id for it in fm.feature} for fid, (s0, m, _) in before.items())
    # removeFront: the second newest frame disappears, the newest takes its index
    fm2 = FeatureManagerOracle(); fill(fm2, frames)
    b2 = {it.feature_id: (it.start_frame, len(it.feature_per_frame)) for it in fm2.feature}
    fm2.removeFront(10)
    for it in fm2.feature:
        s0, m = b2[it.feature_id]
        if s0 == 10:
            assert it.start_frame == 9
        elif s0 + m - 1 >= 9:
            assert len(it.feature_per_frame) == m - 1
        else:
            assert len(it.feature_per_frame) == m


def test_keyframe_decision_follows_the_parallax():
    """addFeatureCheckParallax (feature_manager.cpp:57-116): few tracked / long-tracked features -> keyframe; otherwise the mean
    displacement between the second and third newest frames against MIN_PARALLAX = 10 px / 600."""
    fm = FeatureManagerOracle()
    pts = {k: np.array([0.01 * (k % 10), 0.01 * (k // 10), 1, 0, 0, 0, 0, 2.0]) for k in range(60)}
    for f in range(5):
        assert fm.addFeatureCheckParallax(f, pts, 0.0) is (f < 3)               # long_track_num < 40 until a fourth observation exists; then static: no parallax
    assert fm.addFeatureCheckParallax(5, pts, 0.0) is False
    moved = {k: v + np.array([0.03, 0, 0, 0, 0, 0, 0, 0]) for k, v in pts.items()}
    assert fm.addFeatureCheckParallax(6, moved, 0.0) is False                    # parallax is measured one frame late (frames count-2, count-1)
    assert fm.addFeatureCheckParallax(7, moved, 0.0) is True                     # 0.03 > 10 / 600
    assert abs(fm.last_average_parallax - 0.03 * 600) < 1e-9


def test_replay_loop_with_the_cpu_oracles_tracks_the_trajectory():
    """FeatureTracker -> FeatureManager -> optimization (+ MARGIN_OLD / MARGIN_SECOND_NEW, window slides) over a synthetic RGB-D +
    IMU stream, all components the CPU oracles: the loop stays within millimetres of the ground truth."""
    from ground_fusion_b200.replay import replay
    from ground_fusion_b200.synth import IDC_CAM, SyntheticStream
    from oracle.replay_adapters import oracle_components
    cam = dict(IDC_CAM, k1=0.0, k2=0.0, p1=0.0, p2=0.0)        # the synthetic renderer is an ideal pinhole
    tr, fm, ba = oracle_components(cam, depth_threshold=4.0)
    r = replay(SyntheticStream(seed=0), tr, fm, ba, 36)
    assert len(r["P_est"]) == 26 and r["n_margin_old"] >= 2 and r["n_margin_second_new"] >= 5
    assert r["ate_m"] < 5e-3


def _prepared(seed=3, noise=1e-4):
    Ps, Rs, lms, frames = scene(seed=seed, n_lm=150, noise=noise)
    fm = FeatureManagerOracle(depth_threshold=50.0)
    fill(fm, frames)
    tic, ric = np.array([0.02, -0.01, 0.03]), np.eye(3)
    fm.triangulateAll(10, Ps, Rs, np.zeros(3), np.eye(3))      # the scene is generated with body = camera
    return fm, Ps, Rs, lms, tic, ric


def test_outlier_rejection_and_consistency_check_flag_the_corrupted_landmarks():
    """Estimator::outliersRejection / movingConsistencyCheckW (estimator.cpp:3909-4011) on a consistent scene: nothing is flagged;
    a landmark whose depth is wrong by a factor re-projects off its later observations and is flagged by both."""
    fm, Ps, Rs, lms, _, _ = _prepared()
    z, I = np.zeros(3), np.eye(3)
    assert fm.outliersRejection(Ps, Rs, z, I) == set() and fm.movingConsistencyCheckW(Ps, Rs, z, I) == set()
    long_ = [it for it in fm.feature if len(it.feature_per_frame) >= 6 and it.start_frame < 8 and it.estimated_depth > 0]
    bad = {long_[0].feature_id, long_[3].feature_id}
    for it in fm.feature:
        if it.feature_id in bad:
            it.estimated_depth *= 0.35
    assert fm.outliersRejection(Ps, Rs, z, I) == bad
    assert bad <= fm.movingConsistencyCheckW(Ps, Rs, z, I)


def test_prediction_lands_on_the_true_landmark_under_constant_velocity():
    """predictPtsInNextFrame (estimator.cpp:3853-3886): with poses that really move at constant velocity the predicted camera-frame
    point is the landmark seen from the next pose."""
    rng = np.random.default_rng(0)
    a = 0.03
    dR = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    dP = np.array([0.1, 0.01, 0.02])
    Rs, Ps = [np.eye(3)], [np.zeros(3)]
    for _ in range(11):
        Ps.append(Ps[-1] + Rs[-1] @ dP); Rs.append(Rs[-1] @ dR)
    Rs, Ps = np.array(Rs), np.array(Ps)
    tic, ric = np.array([0.05, 0.0, 0.02]), np.array([[0, 0, 1.0], [-1, 0, 0], [0, -1, 0]])
    fm = FeatureManagerOracle()
    lms = np.stack([rng.uniform(3, 8, 30), rng.uniform(-2, 2, 30), rng.uniform(-1, 1, 30)], 1)
    for f in range(11):
        img = {}
        for l in range(30):
            pc = ric.T @ (Rs[f].T @ (lms[l] - Ps[f]) - tic)
            img[l] = np.array([pc[0] / pc[2], pc[1] / pc[2], 1.0, 0, 0, 0, 0, pc[2]])
        fm.addFeatureCheckParallax(f, img, 0.0)
    for it in fm.feature:
        pc = ric.T @ (Rs[0].T @ (lms[it.feature_id] - Ps[0]) - tic)
        it.estimated_depth = pc[2]
    pred = fm.predictPtsInNextFrame(10, Ps, Rs, tic, ric)
    assert len(pred) == 30
    for l, p in pred.items():
        want = ric.T @ (Rs[11].T @ (lms[l] - Ps[11]) - tic)
        assert np.allclose(p, want, atol=1e-9)
    assert fm.predictPtsInNextFrame(1, Ps, Rs, tic, ric) == {}


@pytest.mark.parametrize("use_mcc", [False, True])
def test_replay_with_the_estimator_feeding_the_tracker_back(use_mcc):
    """The loop with MULTIPLE_THREAD 0 semantics (estimator.cpp:1104-1136): after every optimisation the tracker gets
    removeOutliers + setPrediction; driven by the oracles the trajectory stays within millimetres of the ground truth."""
    from ground_fusion_b200.replay import replay
    from ground_fusion_b200.synth import IDC_CAM, SyntheticStream
    from oracle.replay_adapters import oracle_components
    cam = dict(IDC_CAM, k1=0.0, k2=0.0, p1=0.0, p2=0.0)
    tr, fm, ba = oracle_components(cam, depth_threshold=4.0)
    r = replay(SyntheticStream(seed=0), tr, fm, ba, 30, feedback=True, use_mcc=use_mcc)
    assert len(r["iterations"]) == 20 and r["n_predicted"] > 20 * 50
    assert r["ate_m"] < 5e-3, r["ate_m"]
