"""Host-side mirror of the reference's FeatureManager (vins_estimator/src/estimator/feature_manager.{h,cpp}) over the C ABI.

The observation lists (std::list<FeaturePerId> in the reference) are plain Python lists here -- bookkeeping only; every
per-landmark computation (triangulateWithDepth, triangulate, the keyframe parallax, the depth transfer of
removeBackShiftDepth) runs in libgf_b200.so (csrc/fm_kernels.cu).  There is no CPU fallback.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import check

INIT_DEPTH = 5.0            # parameters.cpp:478
FOCAL_LENGTH = 600.0        # parameters.h:23
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)


def _d(a):
    return a.ctypes.data_as(_dp)


class _Feature:             # FeaturePerId (feature_manager.h:66-88); observations as rows [x y z u v vx vy depth td]
    __slots__ = ("feature_id", "start_frame", "obs", "used_num", "estimated_depth", "estimate_flag", "solve_flag")

    def __init__(self, fid, start):
        self.feature_id, self.start_frame, self.obs = int(fid), int(start), []
        self.used_num = 0; self.estimated_depth = -1.0; self.estimate_flag = 0; self.solve_flag = 0

    def endFrame(self):
        return self.start_frame + len(self.obs) - 1


class FeatureManager:
    """Method names and argument meaning as in feature_manager.h:90-150."""

    def __init__(self, min_parallax_px=10.0, depth_threshold=3.0, window_size=10, device=0):
        self.L = _lib.lib()
        self.L.gf_fm_triangulate.argtypes = [ctypes.c_int, ctypes.c_int, _ip, _ip, _ip, ctypes.c_int, _dp, _dp, _dp, _ip, ctypes.c_int, _dp, _dp, _dp, _dp,
                                             ctypes.c_double, ctypes.c_double]
        self.L.gf_fm_parallax.argtypes = [ctypes.c_int, ctypes.c_int, _dp, _dp, _dp]
        self.L.gf_fm_back_shift_depth.argtypes = [ctypes.c_int, ctypes.c_int, _dp, _dp, _dp, _dp, _dp, _dp, ctypes.c_double]
        self.L.gf_fm_reprojection_errors.argtypes = [ctypes.c_int, ctypes.c_int, _ip, _ip, _ip, ctypes.c_int, _dp, _dp, ctypes.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _ip]
        self.L.gf_fm_predict_next.argtypes = [ctypes.c_int, ctypes.c_int, _ip, _dp, _dp, ctypes.c_int, ctypes.c_int, _dp, _dp, _dp, _dp, _dp]
        self.device = int(device)
        self.feature = []
        self.MIN_PARALLAX = min_parallax_px / FOCAL_LENGTH       # parameters.cpp:345-346
        self.depth_threshold = float(depth_threshold)             # parameters.cpp:172
        self.WINDOW_SIZE = window_size
        self.last_track_num = self.new_feature_num = self.long_track_num = 0
        self.last_average_parallax = 0.0

    def getFeatureCount(self):
        cnt = 0
        for it in self.feature:
            it.used_num = len(it.obs)
            cnt += it.used_num >= 4
        return cnt

    def addFeatureCheckParallax(self, frame_count, image, td):
        """feature_manager.cpp:57-116; image: {id: v[8]} as trackImage returns it."""
        self.last_track_num = self.new_feature_num = self.long_track_num = 0
        self.last_average_parallax = 0.0
        index = {it.feature_id: it for it in self.feature}
        for fid in sorted(image):
            row = np.concatenate([np.asarray(image[fid], float)[:8], [float(td)]])
            it = index.get(fid)
            if it is None:
                it = _Feature(fid, frame_count); self.feature.append(it); index[fid] = it
                it.obs.append(row); self.new_feature_num += 1
            else:
                it.obs.append(row); self.last_track_num += 1
                self.long_track_num += len(it.obs) >= 4
        if frame_count < 2 or self.last_track_num < 20 or self.long_track_num < 40 or self.new_feature_num > 0.5 * self.last_track_num:
            return True
        sel = [it for it in self.feature if it.start_frame <= frame_count - 2 and it.endFrame() >= frame_count - 1]
        if not sel:
            return True
        pi = np.ascontiguousarray([it.obs[frame_count - 2 - it.start_frame][:3] for it in sel], np.float64)
        pj = np.ascontiguousarray([it.obs[frame_count - 1 - it.start_frame][:3] for it in sel], np.float64)
        s = ctypes.c_double(0)
        check(self.L.gf_fm_parallax(self.device, len(sel), _d(pi), _d(pj), ctypes.byref(s)))
        self.last_average_parallax = s.value / len(sel) * FOCAL_LENGTH
        return s.value / len(sel) >= self.MIN_PARALLAX

    def iter_ba_features(self):
        """The landmarks Estimator::optimization() instantiates factors for (estimator.cpp:3268-3297): used_num >= 4, in list
        order; yields (start_frame, [(point, velocity, cur_td), ...], estimated_depth, estimate_flag)."""
        for it in self.feature:
            it.used_num = len(it.obs)
            if it.used_num >= 4:
                yield it.start_frame, [(o[0:3], o[5:7], o[8]) for o in it.obs], it.estimated_depth, it.estimate_flag

    def getDepthVector(self):
        out = []
        for it in self.feature:
            it.used_num = len(it.obs)
            if it.used_num >= 4:
                out.append(1.0 / it.estimated_depth)
        return np.array(out)

    def setDepth(self, x):
        k = -1
        for it in self.feature:
            it.used_num = len(it.obs)
            if it.used_num < 4:
                continue
            k += 1
            it.estimated_depth = 1.0 / x[k]
            it.solve_flag = 2 if it.estimated_depth < 0 else 1

    def removeFailures(self):
        self.feature = [it for it in self.feature if it.solve_flag != 2]

    def triangulateAll(self, frameCnt, Ps, Rs, tic, ric):
        """triangulateWithDepth followed by triangulate, as processImage calls them (estimator.cpp:1090-1102), in one kernel."""
        feats = self.feature
        n = len(feats)
        if n == 0:
            return
        nobs = np.array([len(it.obs) for it in feats], np.int32)
        off = np.zeros(n, np.int32); off[1:] = np.cumsum(nobs)[:-1]
        rows = np.concatenate([np.asarray(it.obs, np.float64).reshape(-1, 9) for it in feats], 0) if nobs.sum() else np.zeros((0, 9))
        pts = np.ascontiguousarray(rows[:, 0:3]); dep = np.ascontiguousarray(rows[:, 7])
        start = np.array([it.start_frame for it in feats], np.int32)
        est = np.array([it.estimated_depth for it in feats], np.float64)
        flag = np.array([it.estimate_flag for it in feats], np.int32)
        Ps_ = np.ascontiguousarray(Ps, np.float64).reshape(-1, 3); Rs_ = np.ascontiguousarray(Rs, np.float64).reshape(-1, 9)
        tic_ = np.ascontiguousarray(tic, np.float64); ric_ = np.ascontiguousarray(ric, np.float64).reshape(9)
        check(self.L.gf_fm_triangulate(self.device, n, start.ctypes.data_as(_ip), nobs.ctypes.data_as(_ip), off.ctypes.data_as(_ip), int(nobs.sum()),
                                       _d(pts), _d(dep), _d(est), flag.ctypes.data_as(_ip), len(Ps_), _d(Ps_), _d(Rs_), _d(tic_), _d(ric_),
                                       self.depth_threshold, INIT_DEPTH))
        for it, e, f in zip(feats, est, flag):
            it.used_num = len(it.obs)
            it.estimated_depth, it.estimate_flag = float(e), int(f)

    # ---- the per-landmark loops of the estimator that feed the front end back (estimator.cpp:3853-4011) ----
    def _reprojection_errors(self, feats, Ps, Rs, tic, ric):
        n = len(feats)
        nobs = np.array([len(it.obs) for it in feats], np.int32)
        off = np.zeros(n, np.int32); off[1:] = np.cumsum(nobs)[:-1]
        pts = np.ascontiguousarray(np.concatenate([np.asarray(it.obs, np.float64).reshape(-1, 9)[:, 0:3] for it in feats], 0))
        start = np.array([it.start_frame for it in feats], np.int32)
        est = np.array([it.estimated_depth for it in feats], np.float64)
        Ps_ = np.ascontiguousarray(Ps, np.float64).reshape(-1, 3); Rs_ = np.ascontiguousarray(Rs, np.float64).reshape(-1, 9)
        tic_ = np.ascontiguousarray(tic, np.float64); ric_ = np.ascontiguousarray(ric, np.float64).reshape(9)
        e2 = np.zeros(n); e3 = np.zeros(n); cnt = np.zeros(n, np.int32)
        check(self.L.gf_fm_reprojection_errors(self.device, n, start.ctypes.data_as(_ip), nobs.ctypes.data_as(_ip), off.ctypes.data_as(_ip), int(nobs.sum()),
                                               _d(pts), _d(est), len(Ps_), _d(Ps_), _d(Rs_), _d(tic_), _d(ric_), _d(e2), _d(e3), cnt.ctypes.data_as(_ip)))
        return e2, e3, cnt

    def outliersRejection(self, Ps, Rs, tic, ric):
        """Estimator::outliersRejection (estimator.cpp:3909-3966): ids whose mean reprojection error exceeds 3 px at FOCAL_LENGTH."""
        feats = [it for it in self.feature if len(it.obs) >= 4]
        for it in self.feature:
            it.used_num = len(it.obs)
        if not feats:
            return set()
        e2, _, cnt = self._reprojection_errors(feats, Ps, Rs, tic, ric)
        with np.errstate(invalid="ignore", divide="ignore"):
            bad = (e2 / cnt) * FOCAL_LENGTH > 3
        return {it.feature_id for it, b in zip(feats, bad) if b}

    def movingConsistencyCheckW(self, Ps, Rs, tic, ric):
        """Estimator::movingConsistencyCheckW (estimator.cpp:3968-4011)."""
        feats = [it for it in self.feature if len(it.obs) >= 2 and it.start_frame < self.WINDOW_SIZE - 2 and not it.estimated_depth < 0]
        if not feats:
            return set()
        e2, e3, cnt = self._reprojection_errors(feats, Ps, Rs, tic, ric)
        out = set()
        for it, a, b, c in zip(feats, e2, e3, cnt):
            if c > 0 and (FOCAL_LENGTH * a / c > 10 or b / c > 2.0):
                out.add(it.feature_id)
        return out

    def predictPtsInNextFrame(self, frame_count, Ps, Rs, tic, ric):
        """Estimator::predictPtsInNextFrame (estimator.cpp:3853-3886): {feature_id: xyz in the predicted next camera}, the argument
        of FeatureTracker::setPrediction."""
        if frame_count < 2:
            return {}
        feats = [it for it in self.feature if it.estimated_depth > 0 and len(it.obs) >= 2 and it.start_frame + len(it.obs) - 1 == frame_count]
        if not feats:
            return {}
        n = len(feats)
        first = np.array([it.start_frame for it in feats], np.int32)
        uv = np.ascontiguousarray([np.asarray(it.obs[0], np.float64)[0:3] for it in feats])
        dep = np.array([it.estimated_depth for it in feats], np.float64)
        Ps_ = np.ascontiguousarray(Ps, np.float64).reshape(-1, 3); Rs_ = np.ascontiguousarray(Rs, np.float64).reshape(-1, 9)
        tic_ = np.ascontiguousarray(tic, np.float64); ric_ = np.ascontiguousarray(ric, np.float64).reshape(9)
        out = np.zeros((n, 3))
        check(self.L.gf_fm_predict_next(self.device, n, first.ctypes.data_as(_ip), _d(uv), _d(dep), len(Ps_), int(frame_count), _d(Ps_), _d(Rs_), _d(tic_), _d(ric_), _d(out)))
        return {it.feature_id: out[k].copy() for k, it in enumerate(feats)}

    def removeOutlier(self, outlierIndex):
        self.feature = [it for it in self.feature if it.feature_id not in outlierIndex]

    def removeBackShiftDepth(self, marg_R, marg_P, new_R, new_P):
        """feature_manager.cpp:818-856"""
        keep, moved, uv = [], [], []
        for it in self.feature:
            if it.start_frame != 0:
                it.start_frame -= 1; keep.append(it); continue
            first = it.obs.pop(0)
            if len(it.obs) < 2:
                continue
            moved.append(it); uv.append(first[:3]); keep.append(it)
        if moved:
            uv = np.ascontiguousarray(uv, np.float64); est = np.array([it.estimated_depth for it in moved], np.float64)
            a = [np.ascontiguousarray(v, np.float64).ravel() for v in (marg_R, marg_P, new_R, new_P)]
            check(self.L.gf_fm_back_shift_depth(self.device, len(moved), _d(uv), _d(est), _d(a[0]), _d(a[1]), _d(a[2]), _d(a[3]), INIT_DEPTH))
            for it, e in zip(moved, est):
                it.estimated_depth = float(e)
        self.feature = keep

    def removeBack(self):
        keep = []
        for it in self.feature:
            if it.start_frame != 0:
                it.start_frame -= 1; keep.append(it)
            else:
                it.obs.pop(0)
                if it.obs:
                    keep.append(it)
        self.feature = keep

    def removeFront(self, frame_count):
        keep = []
        for it in self.feature:
            if it.start_frame == frame_count:
                it.start_frame -= 1; keep.append(it); continue
            j = self.WINDOW_SIZE - 1 - it.start_frame
            if it.endFrame() < frame_count - 1:
                keep.append(it); continue
            del it.obs[j]
            if it.obs:
                keep.append(it)
        self.feature = keep
