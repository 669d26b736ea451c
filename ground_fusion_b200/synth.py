"""Seeded synthetic RGB-D streams (SURVEY.md section 8(d) "synthetic inputs").

Data generation only -- not on the hot path.  A band-limited noise texture (plus a few bright boxes so
that saturated pixels exist) is placed on a plane and rendered through a moving pinhole camera by a
homography; depth is the exact ray/plane intersection in millimetres (u16, 0 beyond 10 m).
cv2 is used for the warp when it is importable, otherwise a numpy bilinear sampler.
"""
import math

import numpy as np

# config/realsense/idc_cam.yaml of the reference (PINHOLE: fx fy cx cy, radtan k1 k2 p1 p2): the camera of configs C1-C3
IDC_CAM = dict(fx=6.2097277909374247e+02, fy=6.2212293397677581e+02, cx=3.1175896455154810e+02,
               cy=2.4718077836114819e+02, k1=1.4865749308203452e-01, k2=-4.6815685578576460e-01,
               p1=1.6205585303208318e-03, p2=-8.9101576735577930e-03)


def idc_params8():
    """gf_tracker_cfg.pinhole order: fx fy cx cy k1 k2 p1 p2."""
    c = IDC_CAM
    return [c["fx"], c["fy"], c["cx"], c["cy"], c["k1"], c["k2"], c["p1"], c["p2"]]

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


def _gauss_blur_np(a, sigma):
    r = int(3 * sigma + 0.5)
    k = np.exp(-0.5 * (np.arange(-r, r + 1) / sigma) ** 2)
    k /= k.sum()
    a = np.apply_along_axis(lambda v: np.convolve(np.pad(v, r, mode="wrap"), k, mode="valid"), 1, a)
    a = np.apply_along_axis(lambda v: np.convolve(np.pad(v, r, mode="wrap"), k, mode="valid"), 0, a)
    return a


def make_texture(seed, size=(1536, 2048), sigma=2.0, contrast=60.0):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal(size).astype(np.float32)
    if cv2 is not None:
        a = cv2.GaussianBlur(a, (0, 0), sigma)
    else:
        a = _gauss_blur_np(a, sigma).astype(np.float32)
    a = (a - a.mean()) / a.std()
    tex = np.clip(128.0 + contrast * a, 0, 255)
    # a few saturated boxes (exercise the grey>250 rule) and dark boxes (flat regions)
    for _ in range(24):
        y, x = int(rng.integers(0, size[0] - 60)), int(rng.integers(0, size[1] - 60))
        hh, ww = int(rng.integers(12, 48)), int(rng.integers(12, 48))
        tex[y:y + hh, x:x + ww] = 255.0 if rng.random() < 0.6 else 20.0
    return tex.astype(np.uint8)


def _rot(rx, ry, rz):
    cx, sx, cy, sy, cz, sz = math.cos(rx), math.sin(rx), math.cos(ry), math.sin(ry), math.cos(rz), math.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


class SyntheticStream:
    """frame(k) -> (t, gray u8 HxW, depth u16 HxW).  Deterministic in (seed, k)."""

    def __init__(self, seed=0, width=640, height=480, fps=30.0, K=None, noise_sigma=1.0,
                 plane_dist=4.0, speed=1.0):
        self.seed, self.w, self.h, self.fps = seed, width, height, fps
        sc = width / 640.0
        self.K = np.array(K if K is not None else
                          [[620.97277909374247 * sc, 0, 311.75896455154810 * sc],
                           [0, 622.12293397677581 * sc, 247.18077836114819 * sc], [0, 0, 1]], float)
        self.tex = make_texture(seed)
        self.noise_sigma = noise_sigma
        self.d0 = plane_dist
        self.speed = speed
        # texture pixel pitch on the plane (metres/pixel) chosen so the texture is ~1 px per image px
        self.pitch = plane_dist / self.K[0, 0]

    def pose(self, k):
        """Camera-from-world rotation R and camera centre C (world), smooth figure-eight-like motion."""
        t = k / self.fps * self.speed
        C = np.array([0.9 * math.sin(0.7 * t), 0.25 * math.sin(1.4 * t), 0.35 * math.sin(0.5 * t)])
        R = _rot(0.03 * math.sin(0.9 * t), 0.12 * math.sin(0.7 * t + 0.3), 0.05 * math.sin(0.6 * t))
        return R, C

    def frame(self, k):
        R, C = self.pose(k)
        th, tw = self.tex.shape
        # world plane Z = d0 ; texture coords (U,V) px -> world (X,Y,d0): X=(U-tw/2)*pitch, Y=(V-th/2)*pitch
        A = np.array([[self.pitch, 0, -tw / 2 * self.pitch], [0, self.pitch, -th / 2 * self.pitch], [0, 0, self.d0]])
        # x_img ~ K R (Xw - C) = K R (A [U V 1]^T - C)
        M = self.K @ R @ (A - np.outer(C, [0, 0, 1]))        # texture px -> image px homography
        Minv = np.linalg.inv(M)
        if cv2 is not None:
            img = cv2.warpPerspective(self.tex, M, (self.w, self.h), flags=cv2.INTER_LINEAR,
                                      borderMode=cv2.BORDER_REFLECT_101).astype(np.float32)
        else:  # numpy bilinear
            ys, xs = np.mgrid[0:self.h, 0:self.w]
            p = Minv @ np.stack([xs.ravel(), ys.ravel(), np.ones(xs.size)])
            u, v = p[0] / p[2], p[1] / p[2]
            u = np.clip(u, 0, tw - 1.001); v = np.clip(v, 0, th - 1.001)
            u0, v0 = np.floor(u).astype(int), np.floor(v).astype(int)
            fu, fv = u - u0, v - v0
            T = self.tex.astype(np.float32)
            img = ((1 - fu) * (1 - fv) * T[v0, u0] + fu * (1 - fv) * T[v0, u0 + 1] +
                   (1 - fu) * fv * T[v0 + 1, u0] + fu * fv * T[v0 + 1, u0 + 1]).reshape(self.h, self.w)
        if self.noise_sigma > 0:
            rng = np.random.default_rng((self.seed + 1) * 1000003 + k)
            img = img + rng.standard_normal(img.shape).astype(np.float32) * self.noise_sigma
        gray = np.clip(np.rint(img), 0, 255).astype(np.uint8)
        # depth: ray dir (camera) = K^-1 [x y 1]; world dir = R^T dir; hit plane Z=d0: s = (d0 - Cz)/dir_w.z ; depth = s * dir_c.z = s
        ys, xs = np.mgrid[0:self.h, 0:self.w]
        dc = np.linalg.inv(self.K) @ np.stack([xs.ravel(), ys.ravel(), np.ones(xs.size)])
        dw = R.T @ dc
        s = (self.d0 - C[2]) / dw[2]
        depth_m = (s * dc[2]).reshape(self.h, self.w)
        depth = np.where((depth_m > 0) & (depth_m < 10.0), np.rint(depth_m * 1000.0), 0).astype(np.uint16)
        return k / self.fps, gray, depth
