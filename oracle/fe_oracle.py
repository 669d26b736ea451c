"""oracle/fe_oracle.py -- TEST INFRASTRUCTURE ONLY (CPU oracle for the front end).

Line-by-line restatement of the reference's FeatureTracker
(/root/reference/vins_estimator/src/featureTracker/feature_tracker.cpp) on top of the *same three
OpenCV entry points with the same arguments* (cv2 4.13.0: calcOpticalFlowPyrLK, goodFeaturesToTrack,
circle).  The reference itself cannot be compiled here (ROS/Eigen/OpenCV C++ headers absent); the
reference holds no test vectors for this path, so parity is pinned by the fixtures this oracle
generates (tests/golden/, made by tests/golden/make_fe_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.

Defined policy for the one undefined behaviour on the path (feature_tracker.cpp:160-163 reads
cur_img.at<uchar>((int)x, (int)y), i.e. row=x, col=y, which is out of bounds for x >= rows): the
transposed read is reproduced when it is in bounds and treated as "not saturated" otherwise.
"""
import ctypes
import math
import os
import subprocess

import numpy as np

try:  # cv2 is the oracle's arithmetic; the product never imports it
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

_HERE = os.path.dirname(os.path.abspath(__file__))
_SORT_SO = os.path.join(_HERE, "_build", "libgf_oracle_sort.so")
_sort_lib = None


def _sortlib():
    global _sort_lib
    if _sort_lib is None:
        src = os.path.join(_HERE, "stdsort_helper.cpp")
        if not os.path.exists(_SORT_SO) or os.path.getmtime(_SORT_SO) < os.path.getmtime(src):
            os.makedirs(os.path.dirname(_SORT_SO), exist_ok=True)
            subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", _SORT_SO, src])
        L = ctypes.CDLL(_SORT_SO)
        L.gfo_setmask_order.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.gfo_setmask_order.restype = None
        L.gfo_setmask.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 4 + [ctypes.c_void_p] * 2
        L.gfo_setmask.restype = ctypes.c_int
        _sort_lib = L
    return _sort_lib


_glue_lib = None


def _gluelib():
    global _glue_lib
    if _glue_lib is None:
        so = os.path.join(_HERE, "_build", "libgf_oracle_glue.so")
        src = os.path.join(_HERE, "fe_glue.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            os.makedirs(os.path.dirname(so), exist_ok=True)
            subprocess.check_call(["gcc", "-O3", "-ffp-contract=off", "-shared", "-fPIC", "-fvisibility=hidden", "-o", so, src, "-lm"])
        L = ctypes.CDLL(so)
        vp, i = ctypes.c_void_p, ctypes.c_int
        L.gfo_status_rules.argtypes = [vp, vp, vp, vp, vp, i, i, vp, i, i, vp]
        L.gfo_status_rules.restype = None
        L.gfo_finalize.argtypes = [vp, vp, i, vp, vp, vp, i, ctypes.c_double, vp, i, i, i, vp, vp]
        L.gfo_finalize.restype = None
        _glue_lib = L
    return _glue_lib


def setmask_order(track_cnt):
    """Permutation produced by the reference's std::sort (feature_tracker.cpp:66-67)."""
    tc = np.ascontiguousarray(track_cnt, np.int32)
    perm = np.empty(len(tc), np.int32)
    if len(tc):
        _sortlib().gfo_setmask_order(tc.ctypes.data, len(tc), perm.ctypes.data)
    return perm


def cv_round(v):
    """cvRound: round-half-to-even (SSE cvtss2si)."""
    return int(np.rint(np.float32(v)))


def c_round(v):
    """C round(): half away from zero (feature_tracker.cpp:360)."""
    v = float(v)
    return int(math.floor(v + 0.5)) if v >= 0 else -int(math.floor(-v + 0.5))


class PinholeCamera:
    """camodocal::PinholeCamera subset (camera_models/src/camera_models/PinholeCamera.cc:450-541,646-662)."""

    def __init__(self, fx, fy, cx, cy, k1=0.0, k2=0.0, p1=0.0, p2=0.0):
        self.fx, self.fy, self.cx, self.cy = float(fx), float(fy), float(cx), float(cy)
        self.k1, self.k2, self.p1, self.p2 = float(k1), float(k2), float(p1), float(p2)
        self.no_distortion = (k1 == 0.0 and k2 == 0.0 and p1 == 0.0 and p2 == 0.0)
        self.inv_K11 = 1.0 / self.fx
        self.inv_K13 = -self.cx / self.fx
        self.inv_K22 = 1.0 / self.fy
        self.inv_K23 = -self.cy / self.fy

    def params8(self):
        return [self.fx, self.fy, self.cx, self.cy, self.k1, self.k2, self.p1, self.p2]

    def distortion(self, x, y):
        mx2 = x * x
        my2 = y * y
        mxy = x * y
        rho2 = mx2 + my2
        rad = self.k1 * rho2 + self.k2 * rho2 * rho2
        dx = x * rad + 2.0 * self.p1 * mxy + self.p2 * (rho2 + 2.0 * mx2)
        dy = y * rad + 2.0 * self.p2 * mxy + self.p1 * (rho2 + 2.0 * my2)
        return dx, dy

    def lift_projective(self, u, v):
        mx_d = self.inv_K11 * u + self.inv_K13
        my_d = self.inv_K22 * v + self.inv_K23
        if self.no_distortion:
            return mx_d, my_d, 1.0
        dx, dy = self.distortion(mx_d, my_d)
        mx_u = mx_d - dx
        my_u = my_d - dy
        for _ in range(1, 8):
            dx, dy = self.distortion(mx_u, my_u)
            mx_u = mx_d - dx
            my_u = my_d - dy
        return mx_u, my_u, 1.0

    def space_to_plane(self, X, Y, Z):
        x, y = X / Z, Y / Z
        if not self.no_distortion:
            dx, dy = self.distortion(x, y)
            x, y = x + dx, y + dy
        return self.fx * x + self.cx, self.fy * y + self.cy


from ground_fusion_b200.synth import IDC_CAM  # noqa: E402  (config/realsense/idc_cam.yaml; shared with bench.py's GPU arm)


class FeatureTrackerOracle:
    """State and methods named as in feature_tracker.h:43-99."""

    def __init__(self, camera, max_cnt=150, min_dist=30, flow_back=1, depth_cam=1):
        assert cv2 is not None, "the FE oracle needs cv2 (opencv 4.13.0)"
        self.cam = camera
        self.MAX_CNT, self.MIN_DIST, self.FLOW_BACK = int(max_cnt), int(min_dist), int(flow_back)
        self.depth_cam = int(depth_cam)
        self.n_id = 0
        self.hasPrediction = False
        self.prev_img = None
        self.prev_pts = np.zeros((0, 2), np.float32)
        self.cur_pts = np.zeros((0, 2), np.float32)
        self.predict_pts = np.zeros((0, 2), np.float32)
        self.ids = []
        self.track_cnt = []
        self.prev_un_pts_map = {}
        self.cur_un_pts_map = {}
        self.prev_time = 0.0
        self.cur_time = 0.0
        self.last_status = None      # combined LK status of the last call (the "inlier mask")
        self.last_n_pts = None       # corners returned by goodFeaturesToTrack in the last call

    # feature_tracker.cpp:14-20
    def inBorder(self, pt):
        x, y = cv_round(pt[0]), cv_round(pt[1])
        return 1 <= x < self.col - 1 and 1 <= y < self.row - 1

    @staticmethod
    def _lk(prev, cur, p0, p1, max_level, initial):
        flags = cv2.OPTFLOW_USE_INITIAL_FLOW if initial else 0
        crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01)
        nxt = p1.reshape(-1, 1, 2).copy() if initial else None
        q, st, _ = cv2.calcOpticalFlowPyrLK(prev, cur, p0.reshape(-1, 1, 2), nxt, winSize=(21, 21),
                                            maxLevel=max_level, criteria=crit, flags=flags)
        return q.reshape(-1, 2).astype(np.float32), st.ravel().astype(np.uint8)

    # feature_tracker.cpp:56-83
    def setMask(self):
        self.mask = np.full((self.row, self.col), 255, np.uint8)
        perm = setmask_order(self.track_cnt)
        pts, ids, cnt = self.cur_pts, self.ids, self.track_cnt
        k_pts, k_ids, k_cnt = [], [], []
        for j in perm:
            px, py = cv_round(pts[j][0]), cv_round(pts[j][1])
            if self.mask[py, px] == 255:
                k_pts.append(pts[j]); k_ids.append(ids[j]); k_cnt.append(cnt[j])
                cv2.circle(self.mask, (px, py), self.MIN_DIST, 0, -1)
        self.cur_pts = np.array(k_pts, np.float32).reshape(-1, 2)
        self.ids, self.track_cnt = k_ids, k_cnt

    # feature_tracker.cpp:103-372
    def trackImage(self, cur_time, img, depth=None):
        self.cur_time = float(cur_time)
        cur_img = np.ascontiguousarray(img, np.uint8)
        self.row, self.col = cur_img.shape
        self.cur_pts = np.zeros((0, 2), np.float32)
        self.last_status = np.zeros(0, np.uint8)

        if len(self.prev_pts) > 0:
            if self.hasPrediction:
                self.cur_pts, status = self._lk(self.prev_img, cur_img, self.prev_pts, self.predict_pts, 1, True)
                if int(status.sum()) < 10:
                    self.cur_pts, status = self._lk(self.prev_img, cur_img, self.prev_pts, None, 3, False)
            else:
                self.cur_pts, status = self._lk(self.prev_img, cur_img, self.prev_pts, None, 3, False)
            if self.FLOW_BACK:
                reverse_pts, reverse_status = self._lk(cur_img, self.prev_img, self.cur_pts, self.prev_pts, 1, True)
                for i in range(len(status)):
                    dx = float(np.float32(self.prev_pts[i][0]) - np.float32(reverse_pts[i][0]))
                    dy = float(np.float32(self.prev_pts[i][1]) - np.float32(reverse_pts[i][1]))
                    ok = status[i] and reverse_status[i] and math.sqrt(dx * dx + dy * dy) <= 0.5
                    status[i] = 1 if ok else 0
            for i in range(len(self.cur_pts)):
                if status[i] and not self.inBorder(self.cur_pts[i]):
                    status[i] = 0
                if status[i]:
                    p_u, p_v = int(self.cur_pts[i][0]), int(self.cur_pts[i][1])   # (int) truncation
                    grey = cur_img[p_u, p_v] if (0 <= p_u < self.row and 0 <= p_v < self.col) else 0
                    if grey > 250:
                        status[i] = 0
            self.last_status = status.copy()
            keep = status.astype(bool)
            self.prev_pts = self.prev_pts[keep]
            self.cur_pts = self.cur_pts[keep]
            self.ids = [v for v, k in zip(self.ids, keep) if k]
            self.track_cnt = [v for v, k in zip(self.track_cnt, keep) if k]

        self.track_cnt = [n + 1 for n in self.track_cnt]

        self.setMask()
        n_max_cnt = self.MAX_CNT - len(self.cur_pts)
        if n_max_cnt > 0:
            c = cv2.goodFeaturesToTrack(cur_img, n_max_cnt, 0.01, self.MIN_DIST, mask=self.mask)
            n_pts = np.zeros((0, 2), np.float32) if c is None else c.reshape(-1, 2).astype(np.float32)
        else:
            n_pts = np.zeros((0, 2), np.float32)
        self.last_n_pts = n_pts
        # addPoints (:85-93)
        if len(n_pts):
            self.cur_pts = np.concatenate([self.cur_pts, n_pts], 0)
            for _ in range(len(n_pts)):
                self.ids.append(self.n_id); self.n_id += 1
                self.track_cnt.append(1)

        # undistortedPts (:797-808)
        cur_un_pts = np.zeros((len(self.cur_pts), 2), np.float32)
        for i, p in enumerate(self.cur_pts):
            x, y, z = self.cam.lift_projective(float(p[0]), float(p[1]))
            cur_un_pts[i] = (np.float32(x / z), np.float32(y / z))
        # ptsVelocity (:810-847)
        self.cur_un_pts_map = {}
        for i, fid in enumerate(self.ids):
            self.cur_un_pts_map.setdefault(fid, cur_un_pts[i].copy())
        pts_velocity = np.zeros((len(self.cur_pts), 2), np.float32)
        if self.prev_un_pts_map:
            dt = self.cur_time - self.prev_time
            for i, fid in enumerate(self.ids):
                pv = self.prev_un_pts_map.get(fid)
                if pv is not None:
                    vx = float(np.float32(cur_un_pts[i][0]) - np.float32(pv[0])) / dt
                    vy = float(np.float32(cur_un_pts[i][1]) - np.float32(pv[1])) / dt
                    pts_velocity[i] = (np.float32(vx), np.float32(vy))

        self.prev_img = cur_img
        self.prev_pts = self.cur_pts.copy()
        self.prev_un_pts_map = self.cur_un_pts_map
        self.prev_time = self.cur_time
        self.hasPrediction = False

        # feature_tracker.cpp:318-366: depth_cam == 0 -> depth "-2.4 for debug" (:338); depth_cam != 0 -> entries only when
        # the depth image is present (:342); depth_cam != 0 with an empty depth image returns an EMPTY frame
        featureFrame = {}
        if self.depth_cam == 0 or depth is not None:
            for i, fid in enumerate(self.ids):
                if self.depth_cam:
                    r, c = c_round(self.cur_pts[i][1]), c_round(self.cur_pts[i][0])
                    depth_value = int(depth[r, c]) / 1000
                else:
                    depth_value = -2.4
                featureFrame[fid] = np.array([float(cur_un_pts[i][0]), float(cur_un_pts[i][1]), 1.0,
                                              float(self.cur_pts[i][0]), float(self.cur_pts[i][1]),
                                              float(pts_velocity[i][0]), float(pts_velocity[i][1]),
                                              depth_value], np.float64)
        return featureFrame

    # feature_tracker.cpp:1006-1027
    def setPrediction(self, predictPts):
        self.hasPrediction = True
        pp = []
        for i, fid in enumerate(self.ids):
            if fid in predictPts:
                X, Y, Z = predictPts[fid]
                u, v = self.cam.space_to_plane(float(X), float(Y), float(Z))
                pp.append((np.float32(u), np.float32(v)))
            else:
                pp.append(tuple(self.prev_pts[i]))
        self.predict_pts = np.array(pp, np.float32).reshape(-1, 2)

    # feature_tracker.cpp:1029-1045
    def removeOutliers(self, removePtsIds):
        keep = np.array([fid not in removePtsIds for fid in self.ids], bool)
        self.prev_pts = self.prev_pts[keep]
        self.ids = [v for v, k in zip(self.ids, keep) if k]
        self.track_cnt = [v for v, k in zip(self.track_cnt, keep) if k]


class FeatureTrackerOracleFast(FeatureTrackerOracle):
    """The same restatement with the per-feature glue vectorised (NumPy) or moved to C (setMask: oracle/stdsort_helper.cpp),
    so that a timing of it is dominated by the three OpenCV calls like the compiled C++ reference, not by the Python
    interpreter.  tests/test_fe_oracle.py::test_fast_oracle_equals_loop_oracle keeps it equal to the loop version above.
    t_cv accumulates the seconds spent inside cv2 (calcOpticalFlowPyrLK, goodFeaturesToTrack) and the C setMask."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.ids = np.zeros(0, np.int64)
        self.track_cnt = np.zeros(0, np.int32)
        self.prev_un_ids = np.zeros(0, np.int64)
        self.prev_un = np.zeros((0, 2), np.float32)
        self.t_cv = 0.0
        self._mask = None
        self._cam8 = np.array(self.cam.params8(), np.float64)

    def _lk_t(self, *a):
        import time
        t0 = time.perf_counter()
        r = self._lk(*a)
        self.t_cv += time.perf_counter() - t0
        return r

    def _lift(self, u, v):
        c = self.cam
        mx_d = c.inv_K11 * u + c.inv_K13
        my_d = c.inv_K22 * v + c.inv_K23
        if c.no_distortion:
            return mx_d, my_d
        dx, dy = c.distortion(mx_d, my_d)          # elementwise float64: the same operations in the same order
        mx_u, my_u = mx_d - dx, my_d - dy
        for _ in range(1, 8):
            dx, dy = c.distortion(mx_u, my_u)
            mx_u, my_u = mx_d - dx, my_d - dy
        return mx_u, my_u

    def trackImage(self, cur_time, img, depth=None):
        import time
        G = _gluelib()
        self.cur_time = float(cur_time)
        cur_img = np.ascontiguousarray(img, np.uint8)
        self.row, self.col = cur_img.shape
        self.cur_pts = np.zeros((0, 2), np.float32)
        self.last_status = np.zeros(0, np.uint8)
        if len(self.prev_pts) > 0:
            if self.hasPrediction:
                self.cur_pts, status = self._lk_t(self.prev_img, cur_img, self.prev_pts, self.predict_pts, 1, True)
                if int(status.sum()) < 10:
                    self.cur_pts, status = self._lk_t(self.prev_img, cur_img, self.prev_pts, None, 3, False)
            else:
                self.cur_pts, status = self._lk_t(self.prev_img, cur_img, self.prev_pts, None, 3, False)
            reverse_pts, reverse_status = self.cur_pts, status
            if self.FLOW_BACK:
                reverse_pts, reverse_status = self._lk_t(cur_img, self.prev_img, self.cur_pts, self.prev_pts, 1, True)
            st = np.empty(len(status), np.uint8)
            G.gfo_status_rules(self.prev_pts.ctypes.data, self.cur_pts.ctypes.data, reverse_pts.ctypes.data, status.ctypes.data,
                               reverse_status.ctypes.data, len(status), self.FLOW_BACK, cur_img.ctypes.data, self.row, self.col, st.ctypes.data)
            self.last_status = st
            ok = st.view(np.bool_)
            self.prev_pts = self.prev_pts[ok]; self.cur_pts = self.cur_pts[ok]
            self.ids = self.ids[ok]; self.track_cnt = self.track_cnt[ok]
        self.track_cnt = self.track_cnt + 1
        # setMask (:56-83) in C: std::sort + walk + filled circles
        t0 = time.perf_counter()
        if self._mask is None or self._mask.shape != cur_img.shape:
            self._mask = np.empty(cur_img.shape, np.uint8)
        self._mask.fill(255)
        n = len(self.cur_pts)
        r = np.rint(self.cur_pts).astype(np.int32).T.copy() if n else np.zeros((2, 0), np.int32)     # cvRound
        tc = np.ascontiguousarray(self.track_cnt, np.int32)
        keep = np.empty(max(n, 1), np.int32)
        nk = _sortlib().gfo_setmask(tc.ctypes.data, r[0].ctypes.data, r[1].ctypes.data, n, self.MIN_DIST, self.row, self.col,
                                    self._mask.ctypes.data, keep.ctypes.data)
        keep = keep[:nk]
        self.cur_pts = self.cur_pts[keep]; self.ids = self.ids[keep]; self.track_cnt = self.track_cnt[keep]
        self.mask = self._mask
        n_max_cnt = self.MAX_CNT - len(self.cur_pts)
        n_pts = np.zeros((0, 2), np.float32)
        if n_max_cnt > 0:
            c = cv2.goodFeaturesToTrack(cur_img, n_max_cnt, 0.01, self.MIN_DIST, mask=self.mask)
            if c is not None:
                n_pts = c.reshape(-1, 2)
        self.t_cv += time.perf_counter() - t0
        self.last_n_pts = n_pts
        if len(n_pts):                                                           # addPoints (:85-93)
            self.cur_pts = np.concatenate([self.cur_pts.reshape(-1, 2), n_pts], 0)
            self.ids = np.concatenate([self.ids, np.arange(self.n_id, self.n_id + len(n_pts), dtype=np.int64)])
            self.n_id += len(n_pts)
            self.track_cnt = np.concatenate([self.track_cnt, np.ones(len(n_pts), np.int32)])
        # undistortedPts (:797-808), ptsVelocity (:810-847), observation vectors (:318-366): oracle/fe_glue.c
        n = len(self.cur_pts)
        self.cur_pts = np.ascontiguousarray(self.cur_pts, np.float32)
        cur_un = np.empty((n, 2), np.float32)
        obs = np.empty((n, 8), np.float64)
        if depth is not None:
            depth = np.ascontiguousarray(depth, np.uint16)
        G.gfo_finalize(self.cur_pts.ctypes.data, self.ids.ctypes.data, n, self._cam8.ctypes.data, self.prev_un_ids.ctypes.data,
                       self.prev_un.ctypes.data, len(self.prev_un_ids), self.cur_time - self.prev_time,
                       depth.ctypes.data if depth is not None else None, self.row, self.col, self.depth_cam,
                       cur_un.ctypes.data, obs.ctypes.data)
        self.prev_img = cur_img
        self.prev_pts = self.cur_pts
        self.prev_un_ids, self.prev_un = self.ids, cur_un
        self.prev_un_pts_map = self.cur_un_pts_map = None
        self.prev_time = self.cur_time
        self.hasPrediction = False
        if not (self.depth_cam == 0 or depth is not None):
            return {}
        return dict(zip(self.ids.tolist(), obs))

    def setPrediction(self, predictPts):
        self.hasPrediction = True
        pp = self.prev_pts.copy().reshape(-1, 2)
        for i, fid in enumerate(self.ids.tolist()):
            if fid in predictPts:
                X, Y, Z = predictPts[fid]
                u, v = self.cam.space_to_plane(float(X), float(Y), float(Z))
                pp[i] = (np.float32(u), np.float32(v))
        self.predict_pts = pp

    def removeOutliers(self, removePtsIds):
        keep = ~np.isin(self.ids, np.fromiter(removePtsIds, np.int64, len(removePtsIds)))
        self.prev_pts = self.prev_pts[keep]; self.ids = self.ids[keep]; self.track_cnt = self.track_cnt[keep]


def cv_calls_only(prev_img, cur_img, prev_pts, mask, n_new, min_dist, flow_back=True):
    """The reference's three OpenCV calls for one frame, nothing else (CPU baseline timing aid)."""
    q, st, _ = cv2.calcOpticalFlowPyrLK(prev_img, cur_img, prev_pts.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3)
    if flow_back:
        crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01)
        cv2.calcOpticalFlowPyrLK(cur_img, prev_img, q, prev_pts.reshape(-1, 1, 2).copy(), winSize=(21, 21), maxLevel=1,
                                 criteria=crit, flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
    if n_new > 0:
        cv2.goodFeaturesToTrack(cur_img, n_new, 0.01, min_dist, mask=mask)
