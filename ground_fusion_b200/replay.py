"""Bag-free replay loop: FeatureTracker -> FeatureManager -> Estimator::optimization() (+ marginalisation, window slide) over a
synthetic RGB-D + IMU stream, modelled on the reference's offline driver (vins_estimator/src/KITTIOdomTest.cpp:82-121: read a
frame, inputImage, write the pose) and on Estimator::{processIMU, processImage, optimization, slideWindow}
(vins_estimator/src/estimator/estimator.cpp:743-783, 843-1163, 2890-3636, 3638-3806).

The three components are passed in (duck-typed): the product wires the CUDA FeatureTracker / FeatureManager / BundleAdjuster;
tests and bench.py wire the CPU oracles through the same loop to get the reference pipeline's trajectory and the ATE
(BASELINE.json metric: "ATE mm").  What is NOT replayed: initialisation -- the first WINDOW_SIZE + 1 frames
take their states from the ground truth, as the text of DESIGN.md says -- wheel / GNSS inputs, failure detection.
"""
import math

import numpy as np

from .ba_problem import Problem
from .synth_ba import Preintegration, R_to_q, q_to_R

WINDOW_SIZE = 10
G = np.array([0.0, 0.0, 9.805])      # parameters.cpp:74, groundchallenge.yaml g_norm


class ImuFromStream:
    """IMU samples consistent with SyntheticStream.pose: body frame = camera frame (ric = I, tic = 0), specific force
    a_m = R_wb^T (a_w + G), body rate from the rotation; derivatives by central differences of the analytic pose."""

    def __init__(self, stream, h=1e-3):
        self.st, self.h = stream, h

    def pose_t(self, t):
        R, C = self.st.pose(t * self.st.fps)          # pose(k) evaluates the trajectory at k / fps
        return R.T, C                                 # R_wb, P

    def sample(self, t):
        h = self.h
        R0, P0 = self.pose_t(t); Rp, Pp = self.pose_t(t + h); Rm, Pm = self.pose_t(t - h)
        a_w = (Pp - 2 * P0 + Pm) / (h * h)
        W = R0.T @ (Rp - Rm) / (2 * h)
        gyr = np.array([W[2, 1] - W[1, 2], W[0, 2] - W[2, 0], W[1, 0] - W[0, 1]]) * 0.5
        return R0.T @ (a_w + G), gyr

    def velocity(self, t):
        h = self.h
        return (self.pose_t(t + h)[1] - self.pose_t(t - h)[1]) / (2 * h)


class WindowEstimator:
    """Estimator's sliding window (estimator.h:229-341) with the NON_LINEAR branch of processImage."""

    def __init__(self, fm, ba, imu_propagate=True, max_iter=8, tracker=None, use_mcc=False):
        """tracker: the front end, when the estimator feeds it back as the reference does with MULTIPLE_THREAD 0 (removeOutliers +
        setPrediction after every optimisation, estimator.cpp:1132-1136); None: no feedback (MULTIPLE_THREAD 1).  use_mcc: USE_MCC."""
        W = WINDOW_SIZE
        self.fm, self.ba = fm, ba
        self.tracker, self.use_mcc = tracker, use_mcc
        self.n_removed = self.n_predicted = 0
        self.Ps = np.zeros((W + 1, 3)); self.Rs = np.tile(np.eye(3), (W + 1, 1, 1)); self.Vs = np.zeros((W + 1, 3))
        self.Bas = np.zeros((W + 1, 3)); self.Bgs = np.zeros((W + 1, 3))
        self.pre = [None] * (W + 1); self.bufs = [[] for _ in range(W + 1)]
        self.tic, self.ric, self.td = np.zeros(3), np.eye(3), 0.0
        self.frame_count = 0; self.first_imu = False; self.acc_0 = self.gyr_0 = None
        self.prior = None
        self.imu_propagate, self.max_iter = imu_propagate, max_iter
        self.nonlinear = False
        self.last_summary = None; self.n_old = self.n_second_new = 0

    # estimator.cpp:743-783 (imu_propagate restores the state propagation this fork leaves to the wheel dead reckoning)
    def processIMU(self, dt, acc, gyr):
        fc = self.frame_count
        if not self.first_imu:
            self.first_imu = True; self.acc_0, self.gyr_0 = acc.copy(), gyr.copy()
        if self.pre[fc] is None:
            self.pre[fc] = Preintegration(self.acc_0, self.gyr_0, self.Bas[fc], self.Bgs[fc])
        if fc != 0:
            self.pre[fc].push_back(dt, acc, gyr); self.bufs[fc].append((dt, acc.copy(), gyr.copy()))
            if self.imu_propagate:
                j = fc
                un_acc_0 = self.Rs[j] @ (self.acc_0 - self.Bas[j]) - G
                un_gyr = 0.5 * (self.gyr_0 + gyr) - self.Bgs[j]
                th = un_gyr * dt
                dq = np.array([th[0] / 2, th[1] / 2, th[2] / 2, 1.0]); dq /= np.linalg.norm(dq)       # Utility::deltaQ
                self.Rs[j] = self.Rs[j] @ q_to_R(dq)
                un_acc_1 = self.Rs[j] @ (acc - self.Bas[j]) - G
                un_acc = 0.5 * (un_acc_0 + un_acc_1)
                self.Ps[j] = self.Ps[j] + dt * self.Vs[j] + 0.5 * dt * dt * un_acc
                self.Vs[j] = self.Vs[j] + dt * un_acc
        self.acc_0, self.gyr_0 = acc.copy(), gyr.copy()

    def set_state(self, fc, R, P, V):
        self.Rs[fc], self.Ps[fc], self.Vs[fc] = R, P, V

    # estimator.cpp:843-1163
    def processImage(self, image):
        fc = self.frame_count
        self.margin_old = bool(self.fm.addFeatureCheckParallax(fc, image, self.td))
        if not self.nonlinear:
            if fc < WINDOW_SIZE:
                self.frame_count += 1
                nf = self.frame_count                    # slideWindow is not run while the window fills: the new frame starts from the last one
                self.Ps[nf], self.Rs[nf], self.Vs[nf], self.Bas[nf], self.Bgs[nf] = self.Ps[fc].copy(), self.Rs[fc].copy(), self.Vs[fc].copy(), self.Bas[fc].copy(), self.Bgs[fc].copy()
                return None
            self.nonlinear = True
        self.fm.triangulateAll(fc, self.Ps, self.Rs, self.tic, self.ric)
        removeIndex = set()
        if self.use_mcc:                                  # estimator.cpp:1104-1109
            removeIndex = self.fm.movingConsistencyCheckW(self.Ps, self.Rs, self.tic, self.ric)
            self.fm.removeOutlier(removeIndex)
        self.optimization()
        if not self.use_mcc:                              # :1123-1129 -- this set shadows the outer one: the tracker is not told about it
            self.fm.removeOutlier(self.fm.movingConsistencyCheckW(self.Ps, self.Rs, self.tic, self.ric))
        if self.tracker is not None:                      # :1131-1136 (MULTIPLE_THREAD 0)
            self.tracker.removeOutliers(removeIndex)
            pred = self.fm.predictPtsInNextFrame(self.frame_count, self.Ps, self.Rs, self.tic, self.ric)
            self.tracker.setPrediction(pred)
            self.n_removed += len(removeIndex); self.n_predicted += len(pred)
        self.slideWindow()
        self.fm.removeFailures()
        return self.Ps[WINDOW_SIZE].copy(), self.Rs[WINDOW_SIZE].copy()

    # estimator.cpp:2890-3636
    def optimization(self):
        W = WINDOW_SIZE
        feats = list(self.fm.iter_ba_features())
        pb = Problem(W + 1, len(feats))
        for i in range(W + 1):                           # vector2double (estimator.cpp:2276-2353)
            pb.para_pose[i, :3] = self.Ps[i]; pb.para_pose[i, 3:] = R_to_q(self.Rs[i])
            pb.para_speed_bias[i] = np.concatenate([self.Vs[i], self.Bas[i], self.Bgs[i]])
        pb.para_ex_pose[:3] = self.tic; pb.para_ex_pose[3:] = R_to_q(self.ric)
        pb.para_td[0] = self.td
        pb.ex_pose_const = 1; pb.td_const = 1            # ESTIMATE_EXTRINSIC 0, ESTIMATE_TD 0
        pb.max_num_iterations = self.max_iter
        rows = []
        for k, (start, obs, est_depth, flag) in enumerate(feats):
            pb.para_feature[k] = 1.0 / est_depth
            pb.feature_const[k] = 1 if flag == 1 else 0   # depth from the depth image: SetParameterBlockConstant (estimator.cpp:3291-3292)
            p0, v0, t0 = obs[0]
            for a in range(1, len(obs)):
                pj, vj, tj = obs[a]
                rows.append((start, start + a, k, p0, pj, v0, vj, t0, tj))
        pb.set_visual(rows)
        pb.set_imu([self.pre[j].as_dict(j - 1, j) for j in range(1, W + 1) if self.pre[j] is not None and self.pre[j].sum_dt <= 10.0])
        pb.prior = self.prior
        R0_before, P0_before = self.Rs[0].copy(), self.Ps[0].copy()
        self.last_summary = self.ba.optimization(pb)
        # double2vector (estimator.cpp:2440-2569)
        from .estimator import double2vector
        Rs, Ps, Vs = double2vector(pb, R0_before, P0_before, True)
        self.Rs, self.Ps, self.Vs = Rs, Ps, Vs
        self.Bas = pb.para_speed_bias[:, 3:6].copy(); self.Bgs = pb.para_speed_bias[:, 6:9].copy()
        self.fm.setDepth(pb.para_feature[:len(feats)])
        # the prior of the next window is linearised at the solved blocks as they stand (vector2double at estimator.cpp:3337 re-reads
        # the rotated state; the rotation is the identity up to the yaw fix of a window that has a prior)
        for i in range(W + 1):
            pb.para_pose[i, :3] = self.Ps[i]; pb.para_pose[i, 3:] = R_to_q(self.Rs[i])
            pb.para_speed_bias[i, :3] = self.Vs[i]
        if self.margin_old:
            self.prior = self.ba.marginalize_old(pb); self.n_old += 1
        else:
            new = self.ba.marginalize_second_new(pb)
            if new is not None:
                self.prior = new
            self.n_second_new += 1

    # estimator.cpp:3638-3806
    def slideWindow(self):
        W = WINDOW_SIZE
        if self.margin_old:
            back_R0, back_P0 = self.Rs[0].copy(), self.Ps[0].copy()
            for i in range(W):
                self.Rs[i], self.Ps[i], self.Vs[i], self.Bas[i], self.Bgs[i] = self.Rs[i + 1].copy(), self.Ps[i + 1].copy(), self.Vs[i + 1].copy(), self.Bas[i + 1].copy(), self.Bgs[i + 1].copy()
                self.pre[i], self.bufs[i] = self.pre[i + 1], self.bufs[i + 1]
            self.pre[W] = Preintegration(self.acc_0, self.gyr_0, self.Bas[W], self.Bgs[W]); self.bufs[W] = []
            R0, R1 = back_R0 @ self.ric, self.Rs[0] @ self.ric
            P0, P1 = back_P0 + back_R0 @ self.tic, self.Ps[0] + self.Rs[0] @ self.tic
            self.fm.removeBackShiftDepth(R0, P0, R1, P1)          # slideWindowOld (estimator.cpp:3808-3837)
        else:
            self.Ps[W - 1], self.Rs[W - 1] = self.Ps[W].copy(), self.Rs[W].copy()
            for dt, acc, gyr in self.bufs[W]:
                self.pre[W - 1].push_back(dt, acc, gyr); self.bufs[W - 1].append((dt, acc, gyr))
            self.Vs[W - 1], self.Bas[W - 1], self.Bgs[W - 1] = self.Vs[W].copy(), self.Bas[W].copy(), self.Bgs[W].copy()
            self.pre[W] = Preintegration(self.acc_0, self.gyr_0, self.Bas[W], self.Bgs[W]); self.bufs[W] = []
            self.fm.removeFront(self.frame_count)                 # slideWindowNew (estimator.cpp:3839-3851)


def replay(stream, tracker, fm, ba, n_frames, imu_per_frame=7, imu_propagate=True, max_iter=8, feedback=False, use_mcc=False):
    """Runs n_frames of `stream` through tracker -> fm -> ba.  Returns dict(t, P_est, P_gt, R_est, R_gt, keyframes, ...).
    tracker.trackImage(t, gray, depth) -> {id: v[8]};  fm / ba: see WindowEstimator.  feedback: the estimator calls the tracker's
    removeOutliers / setPrediction after every optimisation (the reference with MULTIPLE_THREAD 0)."""
    imu = ImuFromStream(stream)
    est = WindowEstimator(fm, ba, imu_propagate=imu_propagate, max_iter=max_iter, tracker=tracker if feedback else None, use_mcc=use_mcc)
    out_t, out_P, out_R, gt_P, gt_R, iters = [], [], [], [], [], []
    t_prev = None
    for k in range(n_frames):
        t, gray, depth = stream.frame(k)
        if t_prev is not None:                          # IMU samples of (t_prev, t]
            dt = (t - t_prev) / imu_per_frame
            for s in range(1, imu_per_frame + 1):
                acc, gyr = imu.sample(t_prev + s * dt)
                est.processIMU(dt, acc, gyr)
        else:
            acc, gyr = imu.sample(t); est.processIMU(0.0, acc, gyr)
        image = tracker.trackImage(t, gray, depth)
        if not est.nonlinear:                           # initialisation is out of scope: ground-truth states while the window fills
            R_wb, P = imu.pose_t(t)
            est.set_state(est.frame_count, R_wb, P, imu.velocity(t))
        res = est.processImage(image)
        t_prev = t
        if res is not None:
            R_wb, P = imu.pose_t(t)
            out_t.append(t); out_P.append(res[0]); out_R.append(res[1]); gt_P.append(P); gt_R.append(R_wb)
            iters.append(est.last_summary["iterations"] if isinstance(est.last_summary, dict) else est.last_summary.get("iterations"))
    P_est, P_gt = np.array(out_P), np.array(gt_P)
    ate = float(np.sqrt(np.mean(np.sum((P_est - P_gt) ** 2, axis=1)))) if len(P_est) else float("nan")
    return {"t": np.array(out_t), "P_est": P_est, "P_gt": P_gt, "R_est": np.array(out_R), "R_gt": np.array(gt_R), "ate_m": ate,
            "n_margin_old": est.n_old, "n_margin_second_new": est.n_second_new, "iterations": iters,
            "n_removed": est.n_removed, "n_predicted": est.n_predicted}
