"""Seeded synthetic sliding-window problems for the back end.

Data generation only -- not on the hot path.  A planar ground-vehicle trajectory is sampled into IMU
measurements (200 Hz, noise of config/realsense/groundchallenge.yaml:114-117), pre-integrated exactly as the
reference's host glue does (IntegrationBase::midPointIntegration, factor/integration_base.h:63-137; this is
sequential per-sample work that stays on the host in the reference too), and landmarks are projected into
the camera with 0.5 px noise.  RGB-D features (70 %) carry a measured depth and are held constant in the
solve, the rest are free and get Schur-eliminated (estimator.cpp:3291-3292).
"""
import math

import numpy as np

from .ba_problem import Prior, Problem

ACC_N, GYR_N, ACC_W, GYR_W = 1.2374091609523514e-02, 3.0032654435730201e-03, 1.9218003442176448e-04, 5.4692100664858005e-05
G_NORM = 9.805
BODY_T_CAM0 = np.array([[0.99957087, 0.00215313, 0.02921355, 0.03668114],
                        [-0.00192891, 0.99996848, -0.00770122, -0.00477653],
                        [-0.02922921, 0.00764156, 0.99954353, 0.0316039],
                        [0, 0, 0, 1.0]])          # config/realsense/groundchallenge.yaml:46-52


# ---- quaternion helpers, storage x y z w (para_Pose order) -----------------------------------------
def q_mul(a, b):
    ax, ay, az, aw = a; bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz])


def q_to_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def R_to_q(R):
    t = np.trace(R)
    if t > 0:
        s = math.sqrt(t + 1.0) * 2; q = [(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s]
    else:
        i = int(np.argmax(np.diag(R))); j, k = (i + 1) % 3, (i + 2) % 3
        s = math.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = [0, 0, 0, 0]; q[i] = 0.25 * s; q[j] = (R[j, i] + R[i, j]) / s; q[k] = (R[k, i] + R[i, k]) / s; q[3] = (R[k, j] - R[j, k]) / s
    q = np.array(q); return q / np.linalg.norm(q) * (1 if q[3] >= 0 else -1)


def skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0.0]])


def rot_exp(w):
    th = np.linalg.norm(w)
    if th < 1e-12:
        return np.eye(3) + skew(w)
    K = skew(w / th)
    return np.eye(3) + math.sin(th) * K + (1 - math.cos(th)) * K @ K


# ---- IntegrationBase (factor/integration_base.h:20-137) --------------------------------------------
class Preintegration:
    def __init__(self, acc_0, gyr_0, ba, bg):
        self.acc_0, self.gyr_0 = np.array(acc_0, float), np.array(gyr_0, float)
        self.linearized_ba, self.linearized_bg = np.array(ba, float), np.array(bg, float)
        self.jacobian = np.eye(15); self.covariance = np.zeros((15, 15)); self.sum_dt = 0.0
        self.delta_p = np.zeros(3); self.delta_q = np.array([0, 0, 0, 1.0]); self.delta_v = np.zeros(3)
        n = np.zeros((18, 18))
        for o, v in ((0, ACC_N), (3, GYR_N), (6, ACC_N), (9, GYR_N), (12, ACC_W), (15, GYR_W)):
            n[o:o + 3, o:o + 3] = v * v * np.eye(3)
        self.noise = n

    def push_back(self, dt, acc_1, gyr_1):
        acc_0, gyr_0, ba, bg = self.acc_0, self.gyr_0, self.linearized_ba, self.linearized_bg
        dq, dp, dv = self.delta_q, self.delta_p, self.delta_v
        Rq = q_to_R(dq)
        un_acc_0 = Rq @ (acc_0 - ba)
        un_gyr = 0.5 * (gyr_0 + gyr_1) - bg
        rq = q_mul(dq, np.array([un_gyr[0] * dt / 2, un_gyr[1] * dt / 2, un_gyr[2] * dt / 2, 1.0]))
        Rr = q_to_R(rq)
        un_acc_1 = Rr @ (acc_1 - ba)
        un_acc = 0.5 * (un_acc_0 + un_acc_1)
        rp = dp + dv * dt + 0.5 * un_acc * dt * dt
        rv = dv + un_acc * dt
        R_w_x, R_a_0_x, R_a_1_x = skew(un_gyr), skew(acc_0 - ba), skew(acc_1 - ba)
        I3 = np.eye(3)
        F = np.zeros((15, 15))
        F[0:3, 0:3] = I3
        F[0:3, 3:6] = -0.25 * Rq @ R_a_0_x * dt * dt + -0.25 * Rr @ R_a_1_x @ (I3 - R_w_x * dt) * dt * dt
        F[0:3, 6:9] = I3 * dt
        F[0:3, 9:12] = -0.25 * (Rq + Rr) * dt * dt
        F[0:3, 12:15] = -0.25 * Rr @ R_a_1_x * dt * dt * -dt
        F[3:6, 3:6] = I3 - R_w_x * dt
        F[3:6, 12:15] = -1.0 * I3 * dt
        F[6:9, 3:6] = -0.5 * Rq @ R_a_0_x * dt + -0.5 * Rr @ R_a_1_x @ (I3 - R_w_x * dt) * dt
        F[6:9, 6:9] = I3
        F[6:9, 9:12] = -0.5 * (Rq + Rr) * dt
        F[6:9, 12:15] = -0.5 * Rr @ R_a_1_x * dt * -dt
        F[9:12, 9:12] = I3; F[12:15, 12:15] = I3
        V = np.zeros((15, 18))
        V[0:3, 0:3] = 0.25 * Rq * dt * dt
        V[0:3, 3:6] = 0.25 * -Rr @ R_a_1_x * dt * dt * 0.5 * dt
        V[0:3, 6:9] = 0.25 * Rr * dt * dt
        V[0:3, 9:12] = V[0:3, 3:6]
        V[3:6, 3:6] = 0.5 * I3 * dt; V[3:6, 9:12] = 0.5 * I3 * dt
        V[6:9, 0:3] = 0.5 * Rq * dt
        V[6:9, 3:6] = 0.5 * -Rr @ R_a_1_x * dt * 0.5 * dt
        V[6:9, 6:9] = 0.5 * Rr * dt
        V[6:9, 9:12] = V[6:9, 3:6]
        V[9:12, 12:15] = I3 * dt; V[12:15, 15:18] = I3 * dt
        self.jacobian = F @ self.jacobian
        self.covariance = F @ self.covariance @ F.T + V @ self.noise @ V.T
        self.delta_p, self.delta_v = rp, rv
        self.delta_q = rq / np.linalg.norm(rq)
        self.sum_dt += dt
        self.acc_0, self.gyr_0 = np.array(acc_1, float), np.array(gyr_1, float)

    def as_dict(self, i, j):
        return dict(i=i, j=j, sum_dt=self.sum_dt, delta_p=self.delta_p, delta_q=self.delta_q, delta_v=self.delta_v,
                    linearized_ba=self.linearized_ba, linearized_bg=self.linearized_bg, jacobian=self.jacobian, covariance=self.covariance)


# ---- trajectory ------------------------------------------------------------------------------------
R0 = np.array([[0, 0, 1.0], [-1, 0, 0], [0, -1, 0]])   # body (x right, y down, z forward) -> heading frame


class Trajectory:
    def __init__(self, seed=0, speed=1.0):
        rng = np.random.default_rng(seed)
        self.a = 4.0 + rng.uniform(-0.5, 0.5); self.b = 2.5 + rng.uniform(-0.3, 0.3)
        self.w = 0.35 * speed; self.ph = rng.uniform(0, 2 * math.pi)
        self.ba = rng.normal(0, 0.02, 3); self.bg = rng.normal(0, 0.002, 3)

    def p(self, t):
        return np.array([self.a * math.sin(self.w * t + self.ph), self.b * math.sin(2 * self.w * t + 2 * self.ph), 0.03 * math.sin(1.3 * t)])

    def v(self, t):
        return np.array([self.a * self.w * math.cos(self.w * t + self.ph), 2 * self.b * self.w * math.cos(2 * self.w * t + 2 * self.ph), 0.039 * math.cos(1.3 * t)])

    def acc(self, t):
        return np.array([-self.a * self.w ** 2 * math.sin(self.w * t + self.ph), -4 * self.b * self.w ** 2 * math.sin(2 * self.w * t + 2 * self.ph), -0.0507 * math.sin(1.3 * t)])

    def R(self, t):
        v = self.v(t)
        yaw = math.atan2(v[1], v[0])
        roll, pitch = 0.02 * math.sin(0.9 * t), 0.015 * math.sin(1.1 * t + 0.4)
        cy, sy = math.cos(yaw), math.sin(yaw)
        Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1.0]])
        return Rz @ rot_exp(np.array([roll, pitch, 0.0])) @ R0

    def omega_b(self, t, h=1e-5):
        Rm, Rp = self.R(t - h), self.R(t + h)
        dR = (Rp - Rm) / (2 * h)
        W = self.R(t).T @ dR
        return np.array([W[2, 1] - W[1, 2], W[0, 2] - W[2, 0], W[1, 0] - W[0, 1]]) * 0.5

    def imu(self, t, rng=None):
        g = np.array([0, 0, G_NORM])
        a = self.R(t).T @ (self.acc(t) + g) + self.ba
        w = self.omega_b(t) + self.bg
        if rng is not None:
            a = a + rng.normal(0, ACC_N * math.sqrt(200.0), 3) * 0.1
            w = w + rng.normal(0, GYR_N * math.sqrt(200.0), 3) * 0.1
        return a, w


def rot_log(R):
    c = max(-1.0, min(1.0, (np.trace(R) - 1.0) / 2.0))
    th = math.acos(c)
    w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    return 0.5 * w if th < 1e-9 else th / (2.0 * math.sin(th)) * w


# body (IMU) <- wheel odometer frame: a small mounting rotation and a lever arm (synthetic; the shipped configs read it from
# body_T_wheel in the yaml)
BODY_T_WHEEL = np.eye(4)
BODY_T_WHEEL[:3, :3] = rot_exp(np.array([0.01, -0.02, 0.03]))
BODY_T_WHEEL[:3, 3] = [0.08, -0.02, -0.15]


def wheel_factors(tr, times, rng, frame_dt, td_true=0.0):
    """Synthetic WheelIntegrationBase results between consecutive frames: the relative pose of the odometer frame with
    noise, plausible sensitivities to the scale intrinsics (sx, sy scale the planar velocity, sw the yaw rate), a 6x6
    covariance, and the boundary wheel speeds / rates the time-offset compensation reads."""
    Rio, tio = BODY_T_WHEEL[:3, :3], BODY_T_WHEEL[:3, 3]
    rows = []
    for f in range(len(times) - 1):
        ta, tb = times[f], times[f + 1]
        Ra, Rb = tr.R(ta) @ Rio, tr.R(tb) @ Rio
        pa, pb_ = tr.p(ta) + tr.R(ta) @ tio, tr.p(tb) + tr.R(tb) @ tio
        dp = Ra.T @ (pb_ - pa)
        dR = Ra.T @ Rb
        dth = rot_log(dR)
        dp_n = dp + rng.normal(0, 0.004, 3)
        dq_n = R_to_q(dR @ rot_exp(rng.normal(0, 0.002, 3)))
        J = np.zeros((6, 3))
        J[:3, 0] = [dp[0], 0.1 * dp[1], 0.0]
        J[:3, 1] = [0.1 * dp[0], dp[1], 0.0]
        J[:3, 2] = 0.5 * np.cross(dth, dp)
        J[3:, 2] = dth
        A = rng.normal(0, 1, (6, 6))
        cov = np.diag([2e-5, 2e-5, 4e-5, 4e-6, 4e-6, 2e-6]) + 1e-7 * (A @ A.T)
        def wheel_meas(t):
            Ro = tr.R(t) @ Rio
            v_o = Ro.T @ (tr.v(t) + tr.R(t) @ np.cross(tr.omega_b(t), tio))
            w_o = Rio.T @ tr.omega_b(t)
            return v_o + rng.normal(0, 0.01, 3), w_o + rng.normal(0, 0.002, 3)
        v0, w0 = wheel_meas(ta); v1, w1 = wheel_meas(tb)
        rows.append(dict(i=f, j=f + 1, sum_dt=frame_dt, delta_p=dp_n, delta_q=dq_n, jacobian=J, covariance=cov,
                         linearized_sx=1.0, linearized_sy=1.0, linearized_sw=1.0, linearized_td=0.0,
                         linearized_vel=v0, linearized_gyr=w0, vel_1=v1, gyr_1=w1))
    return rows


def make_window(seed=0, n_frames=11, n_landmarks=220, frame_dt=0.1, imu_rate=200, t0=2.0, free_fraction=0.3,
                pix_noise=0.5 / 460.0, pose_noise=(0.03, 0.01), with_prior=False, with_wheel=False, wheel_free=(True, True, True),
                with_plane=False, plane_free=True):
    """One optimisation problem around a true trajectory.  Returns (Problem, truth dict)."""
    rng = np.random.default_rng(seed)
    tr = Trajectory(seed)
    ric, tic = BODY_T_CAM0[:3, :3], BODY_T_CAM0[:3, 3]
    times = t0 + frame_dt * np.arange(n_frames)
    P = np.array([tr.p(t) for t in times]); Rm = [tr.R(t) for t in times]; Vv = np.array([tr.v(t) for t in times])
    # landmarks in front of the path
    lms = []
    while len(lms) < n_landmarks:
        k = rng.integers(0, n_frames)
        d = rng.uniform(1.0, 8.0)
        pc = np.array([rng.uniform(-0.45, 0.45) * d, rng.uniform(-0.33, 0.33) * d, d])
        lms.append(Rm[k] @ (ric @ pc + tic) + P[k])
    lms = np.array(lms)
    obs = {}       # landmark -> list of (frame, pts3)
    for f in range(n_frames):
        pc = (ric.T @ ((Rm[f].T @ (lms - P[f]).T) - tic[:, None])).T
        for l in range(n_landmarks):
            z = pc[l, 2]
            if z > 0.4 and abs(pc[l, 0] / z) < 0.5 and abs(pc[l, 1] / z) < 0.37:
                obs.setdefault(l, []).append((f, np.array([pc[l, 0] / z + rng.normal(0, pix_noise), pc[l, 1] / z + rng.normal(0, pix_noise), 1.0]), z))
    feats = []
    for l, lst in obs.items():
        # keep the longest run of consecutive frames (KLT tracks are contiguous)
        runs, cur = [], [lst[0]]
        for a, b in zip(lst, lst[1:]):
            if b[0] == a[0] + 1:
                cur.append(b)
            else:
                runs.append(cur); cur = [b]
        runs.append(cur)
        run = max(runs, key=len)
        if len(run) >= 4:
            feats.append((l, run))
    pb = Problem(n_frames, len(feats))
    vis = []
    for k, (l, run) in enumerate(feats):
        f0, p0, z0 = run[0]
        is_const = rng.random() >= free_fraction
        pb.feature_const[k] = 1 if is_const else 0
        depth = z0 * (1 + rng.normal(0, 0.01 if is_const else 0.10))
        pb.para_feature[k] = 1.0 / depth
        for idx, (f, pj, _) in enumerate(run[1:]):
            prevp = run[idx][1]
            velj = (pj[:2] - prevp[:2]) / frame_dt
            vis.append((f0, f, k, p0, pj, np.zeros(2), velj, 0.0, 0.0))
    pb.set_visual(vis)
    # IMU pre-integration between consecutive frames, biases linearised at the (perturbed) estimate
    ba_est = tr.ba + rng.normal(0, 0.005, 3); bg_est = tr.bg + rng.normal(0, 0.0005, 3)
    imus = []
    n_sub = int(round(frame_dt * imu_rate))
    for f in range(n_frames - 1):
        ts = np.linspace(times[f], times[f + 1], n_sub + 1)
        a0, w0 = tr.imu(ts[0], rng)
        pre = Preintegration(a0, w0, ba_est, bg_est)
        for t in ts[1:]:
            a1, w1 = tr.imu(t, rng)
            pre.push_back(frame_dt / n_sub, a1, w1)
        imus.append(pre.as_dict(f, f + 1))
    pb.set_imu(imus)
    # initial estimates
    for f in range(n_frames):
        dth = rng.normal(0, pose_noise[1], 3)
        pb.para_pose[f, :3] = P[f] + rng.normal(0, pose_noise[0], 3)
        pb.para_pose[f, 3:] = R_to_q(Rm[f] @ rot_exp(dth))
        pb.para_speed_bias[f, :3] = Vv[f] + rng.normal(0, 0.05, 3)
        pb.para_speed_bias[f, 3:6] = ba_est; pb.para_speed_bias[f, 6:9] = bg_est
    pb.para_ex_pose[:3] = tic; pb.para_ex_pose[3:] = R_to_q(ric)
    if with_wheel:
        pb.set_wheel(wheel_factors(tr, times, rng, frame_dt))
        pb.para_ex_wheel[:3] = BODY_T_WHEEL[:3, 3] + rng.normal(0, 0.01, 3)
        pb.para_ex_wheel[3:] = R_to_q(BODY_T_WHEEL[:3, :3] @ rot_exp(rng.normal(0, 0.005, 3)))
        pb.para_ix_wheel[:] = [1.01, 0.99, 1.005]
        pb.para_td_wheel[0] = 0.003
        pb.ex_wheel_const, pb.ix_wheel_const, pb.td_wheel_const = (0 if wheel_free[0] else 1), (0 if wheel_free[1] else 1), (0 if wheel_free[2] else 1)
    if with_plane:
        # ground plane through the odometer origins: z_world of the odometer ~ const; plane frame = world rotated slightly
        pb.set_plane(range(n_frames))
        pb.para_ex_wheel[:3] = BODY_T_WHEEL[:3, 3] + rng.normal(0, 0.01, 3)
        pb.para_ex_wheel[3:] = R_to_q(BODY_T_WHEEL[:3, :3] @ rot_exp(rng.normal(0, 0.005, 3)))
        nrm = np.mean([Rm[f] @ BODY_T_WHEEL[:3, :3] @ np.array([0, 0, 1.0]) for f in range(n_frames)], axis=0)
        nrm /= np.linalg.norm(nrm)
        ax = np.cross(nrm, [0, 0, 1.0]); sn = np.linalg.norm(ax)
        Rpw = rot_exp(ax / sn * math.asin(min(1.0, sn))) if sn > 1e-12 else np.eye(3)     # takes the mean odometer z axis to e3
        Rpw = Rpw @ rot_exp(rng.normal(0, 0.003, 3))
        pb.para_plane_R[:] = R_to_q(Rpw)
        pb.para_plane_Z[0] = -float(np.mean([(Rpw @ (P[f] + Rm[f] @ BODY_T_WHEEL[:3, 3]))[2] for f in range(n_frames)])) + rng.normal(0, 0.01)
        pb.plane_sqrt_info[:] = [50.0, 50.0, 20.0]
        pb.plane_const = 0 if plane_free else 1
    truth = dict(P=P, R=Rm, V=Vv, ba=tr.ba, bg=tr.bg, times=times, landmarks=lms)
    return pb, truth
