"""numpy restatement of the host glue around the solve -- TEST INFRASTRUCTURE ONLY.

double2vector follows reference vins_estimator/src/estimator/estimator.cpp:2440-2494 with Utility::R2ypr / ypr2R from
utility/utility.h:78-120 (degrees, yaw-pitch-roll about z-y-x)."""
import math

import numpy as np


def q_to_R(q):
    x, y, z, w = np.asarray(q, float) / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def R2ypr(R):
    n, o, a = R[:, 0], R[:, 1], R[:, 2]
    y = math.atan2(n[1], n[0])
    p = math.atan2(-n[2], n[0] * math.cos(y) + n[1] * math.sin(y))
    r = math.atan2(a[0] * math.sin(y) - a[1] * math.cos(y), -o[0] * math.sin(y) + o[1] * math.cos(y))
    return np.array([y, p, r]) / math.pi * 180.0


def ypr2R(ypr):
    y, p, r = np.asarray(ypr, float) / 180.0 * math.pi
    Rz = np.array([[math.cos(y), -math.sin(y), 0], [math.sin(y), math.cos(y), 0], [0, 0, 1]])
    Ry = np.array([[math.cos(p), 0, math.sin(p)], [0, 1, 0], [-math.sin(p), 0, math.cos(p)]])
    Rx = np.array([[1, 0, 0], [0, math.cos(r), -math.sin(r)], [0, math.sin(r), math.cos(r)]])
    return Rz @ Ry @ Rx


def double2vector(para_pose, para_speed_bias, R0_before, P0_before, use_imu=True):
    F = len(para_pose)
    if not use_imu:
        return np.array([q_to_R(para_pose[i, 3:]) for i in range(F)]), para_pose[:, :3].copy(), np.zeros((F, 3))
    origin_R0 = R2ypr(np.asarray(R0_before))
    R00 = q_to_R(para_pose[0, 3:])
    origin_R00 = R2ypr(R00)
    rot_diff = ypr2R([origin_R0[0] - origin_R00[0], 0, 0])
    if abs(abs(origin_R0[1]) - 90) < 1.0 or abs(abs(origin_R00[1]) - 90) < 1.0:
        rot_diff = np.asarray(R0_before) @ R00.T
    Rs = np.array([rot_diff @ q_to_R(para_pose[i, 3:]) for i in range(F)])
    Ps = np.array([rot_diff @ (para_pose[i, :3] - para_pose[0, :3]) + P0_before for i in range(F)])
    Vs = np.array([rot_diff @ para_speed_bias[i, :3] for i in range(F)])
    return Rs, Ps, Vs
