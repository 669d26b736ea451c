"""Host-side container of one Estimator::optimization() problem (include/gf_b200.h: gf_ba_problem).

Holds the para_* arrays of the reference's Estimator (estimator.h:229-237) as numpy arrays plus the factor
tables, and exposes them as the ctypes struct both the CUDA library and the CPU oracle consume.
"""
import ctypes

import numpy as np

from ._lib import (BaImuFactor, BaPrior, BaProblem, BaVisualFactor, BaWheelFactor, _dp)


def _ptr(a):
    return a.ctypes.data_as(_dp)


class Prior:
    """MarginalizationInfo as consumed by MarginalizationFactor (marginalization_factor.cpp:332-392)."""

    def __init__(self, kinds, indices, idx, x0, J, r):
        self.kinds, self.indices, self.idx = list(kinds), list(indices), list(idx)
        self.x0 = np.ascontiguousarray(x0, np.float64)
        self.J = np.ascontiguousarray(J, np.float64)
        self.r = np.ascontiguousarray(r, np.float64)
        self.n = len(self.r)

    def struct(self):
        s = BaPrior()
        s.n, s.n_blocks = self.n, len(self.kinds)
        for b, (k, i, c) in enumerate(zip(self.kinds, self.indices, self.idx)):
            s.block_kind[b], s.block_index[b], s.block_idx[b] = k, i, c
        s.x0, s.linearized_jacobians, s.linearized_residuals = _ptr(self.x0), _ptr(self.J), _ptr(self.r)
        return s


class Problem:
    def __init__(self, n_frames, n_features):
        self.n_frames, self.n_features = n_frames, n_features
        self.para_pose = np.zeros((n_frames, 7)); self.para_pose[:, 6] = 1
        self.para_speed_bias = np.zeros((n_frames, 9))
        self.para_ex_pose = np.array([0, 0, 0, 0, 0, 0, 1.0])
        self.para_feature = np.ones(max(n_features, 1))
        self.para_td = np.zeros(1)
        self.para_ex_wheel = np.array([0, 0, 0, 0, 0, 0, 1.0]); self.para_ix_wheel = np.ones(3); self.para_td_wheel = np.zeros(1)
        self.feature_const = np.zeros(max(n_features, 1), np.uint8)
        self.frames_const = self.pose0_const = 0
        self.ex_pose_const = self.td_const = self.ex_wheel_const = self.ix_wheel_const = self.td_wheel_const = 1
        self.visual = (BaVisualFactor * 1)(); self.n_visual = 0
        self.imu = (BaImuFactor * 1)(); self.n_imu = 0
        self.wheel = (BaWheelFactor * 1)(); self.n_wheel = 0
        self.n_plane = 0; self.plane_frames = np.zeros(1, np.int32)        # PlaneFactor (USE_PLANE)
        self.para_plane_R = np.array([0, 0, 0, 1.0]); self.para_plane_Z = np.zeros(1)
        self.plane_const = 1; self.plane_r_subset_mask = 0b100; self.plane_sqrt_info = np.array([100.0, 100.0, 100.0])
        self.ex_wheel_subset_mask = 0                  # PoseSubsetParameterization of the wheel extrinsic (bit k: component k frozen in Plus)
        self.prior = None
        self.gravity = np.array([0.0, 0.0, 9.805])
        self.visual_sqrt_info = 600.0 / 1.5          # FOCAL_LENGTH / 1.5 (estimator.cpp:193)
        self.max_num_iterations = 8
        self._keep = None

    def set_visual(self, rows):
        """rows: iterable of (imu_i, imu_j, feature, pts_i[3], pts_j[3], vel_i[2], vel_j[2], td_i, td_j)"""
        rows = list(rows)
        self.n_visual = len(rows)
        self.visual = (BaVisualFactor * max(len(rows), 1))()
        for f, (a, b, k, pi, pj, vi, vj, ti, tj) in zip(self.visual, rows):
            f.imu_i, f.imu_j, f.feature = int(a), int(b), int(k)
            f.pts_i[:] = list(pi); f.pts_j[:] = list(pj); f.vel_i[:] = list(vi); f.vel_j[:] = list(vj); f.td_i, f.td_j = ti, tj

    def set_imu(self, rows):
        """rows: dicts with i, j, sum_dt, delta_p, delta_q(xyzw), delta_v, linearized_ba, linearized_bg, jacobian(15x15), covariance(15x15)"""
        rows = list(rows)
        self.n_imu = len(rows)
        self.imu = (BaImuFactor * max(len(rows), 1))()
        for f, d in zip(self.imu, rows):
            f.i, f.j, f.sum_dt = int(d["i"]), int(d["j"]), float(d["sum_dt"])
            f.delta_p[:] = list(d["delta_p"]); f.delta_q[:] = list(d["delta_q"]); f.delta_v[:] = list(d["delta_v"])
            f.linearized_ba[:] = list(d["linearized_ba"]); f.linearized_bg[:] = list(d["linearized_bg"])
            f.jacobian[:] = list(np.asarray(d["jacobian"], float).ravel()); f.covariance[:] = list(np.asarray(d["covariance"], float).ravel())

    def set_wheel(self, rows):
        """rows: dicts with i, j, sum_dt, delta_p, delta_q(xyzw), jacobian(6x3), covariance(6x6), linearized_sx/sy/sw/td,
        linearized_vel, linearized_gyr, vel_1, gyr_1 (the WheelIntegrationBase members WheelFactor::Evaluate reads)."""
        rows = list(rows)
        self.n_wheel = len(rows)
        self.wheel = (BaWheelFactor * max(len(rows), 1))()
        for f, d in zip(self.wheel, rows):
            f.i, f.j, f.sum_dt = int(d["i"]), int(d["j"]), float(d["sum_dt"])
            f.delta_p[:] = list(d["delta_p"]); f.delta_q[:] = list(d["delta_q"])
            f.jacobian[:] = list(np.asarray(d["jacobian"], float).ravel()); f.covariance[:] = list(np.asarray(d["covariance"], float).ravel())
            f.linearized_sx, f.linearized_sy, f.linearized_sw, f.linearized_td = (float(d[k]) for k in ("linearized_sx", "linearized_sy", "linearized_sw", "linearized_td"))
            f.linearized_vel[:] = list(d["linearized_vel"]); f.linearized_gyr[:] = list(d["linearized_gyr"])
            f.vel_1[:] = list(d["vel_1"]); f.gyr_1[:] = list(d["gyr_1"])

    def set_plane(self, frames):
        """One PlaneFactor per listed frame (estimator.cpp:3152-3166)."""
        frames = list(frames)                      # materialise once: `frames` may be a generator
        self.plane_frames = np.ascontiguousarray(frames, np.int32) if frames else np.zeros(1, np.int32)
        self.n_plane = len(frames)

    def struct(self):
        p = BaProblem()
        p.n_frames, p.n_features, p.n_visual, p.n_imu, p.n_wheel = self.n_frames, self.n_features, self.n_visual, self.n_imu, self.n_wheel
        p.max_num_iterations = self.max_num_iterations
        p.para_pose, p.para_speed_bias, p.para_ex_pose = _ptr(self.para_pose), _ptr(self.para_speed_bias), _ptr(self.para_ex_pose)
        p.para_feature, p.para_td = _ptr(self.para_feature), _ptr(self.para_td)
        p.para_ex_wheel, p.para_ix_wheel, p.para_td_wheel = _ptr(self.para_ex_wheel), _ptr(self.para_ix_wheel), _ptr(self.para_td_wheel)
        p.feature_const = self.feature_const.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))
        p.frames_const, p.pose0_const, p.ex_pose_const, p.td_const = self.frames_const, self.pose0_const, self.ex_pose_const, self.td_const
        p.ex_wheel_const, p.ix_wheel_const, p.td_wheel_const = self.ex_wheel_const, self.ix_wheel_const, self.td_wheel_const
        p.visual = ctypes.cast(self.visual, ctypes.POINTER(BaVisualFactor))
        p.imu = ctypes.cast(self.imu, ctypes.POINTER(BaImuFactor))
        p.wheel = ctypes.cast(self.wheel, ctypes.POINTER(BaWheelFactor))
        self._keep = self.prior.struct() if self.prior is not None else None
        p.prior = ctypes.pointer(self._keep) if self._keep is not None else None
        p.gravity[:] = list(self.gravity)
        p.visual_sqrt_info = self.visual_sqrt_info
        p.ex_wheel_subset_mask = int(self.ex_wheel_subset_mask)
        p.n_plane = int(self.n_plane)
        p.plane_frames = self.plane_frames.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
        p.para_plane_R, p.para_plane_Z = _ptr(self.para_plane_R), _ptr(self.para_plane_Z)
        p.plane_const, p.plane_r_subset_mask = int(self.plane_const), int(self.plane_r_subset_mask)
        p.plane_sqrt_info[:] = list(self.plane_sqrt_info)
        return p

    def clone(self):
        import copy
        q = Problem(self.n_frames, self.n_features)
        for k in ("para_pose", "para_speed_bias", "para_ex_pose", "para_feature", "para_td", "para_ex_wheel", "para_ix_wheel",
                  "para_td_wheel", "feature_const", "gravity", "para_plane_R", "para_plane_Z", "plane_frames", "plane_sqrt_info"):
            setattr(q, k, getattr(self, k).copy())
        for k in ("frames_const", "pose0_const", "ex_pose_const", "td_const", "ex_wheel_const", "ix_wheel_const", "td_wheel_const",
                  "n_visual", "n_imu", "n_wheel", "visual_sqrt_info", "max_num_iterations", "prior", "ex_wheel_subset_mask", "n_plane", "plane_const",
                  "plane_r_subset_mask"):
            setattr(q, k, getattr(self, k))
        q.visual = (BaVisualFactor * max(self.n_visual, 1))(); ctypes.memmove(q.visual, self.visual, ctypes.sizeof(BaVisualFactor) * self.n_visual)
        q.imu = (BaImuFactor * max(self.n_imu, 1))(); ctypes.memmove(q.imu, self.imu, ctypes.sizeof(BaImuFactor) * self.n_imu)
        q.wheel = (BaWheelFactor * max(self.n_wheel, 1))(); ctypes.memmove(q.wheel, self.wheel, ctypes.sizeof(BaWheelFactor) * self.n_wheel)
        return q
