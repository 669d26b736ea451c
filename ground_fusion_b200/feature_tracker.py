"""Host-side mirror of the reference's FeatureTracker class
(/root/reference/vins_estimator/src/featureTracker/feature_tracker.h:43-99) over the C ABI of
libgf_b200.so.  Method names and argument meaning follow the reference; all arithmetic runs in the
CUDA library (no OpenCV, no CPU fallback).
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import OBS_DTYPE, TrackInfo, TrackerCfg, check

# stage-level wrappers -----------------------------------------------------------------------------


def _u8(a):
    a = np.ascontiguousarray(a, np.uint8)
    assert a.ndim == 2
    return a


def pyr_down(img, device=0):
    img = _u8(img)
    h, w = img.shape
    out = np.empty(((h + 1) // 2, (w + 1) // 2), np.uint8)
    check(_lib.lib().gf_stage_pyr_down(device, img.ctypes.data, w, h, out.ctypes.data))
    return out


def corner_min_eigen_val(img, device=0):
    img = _u8(img)
    h, w = img.shape
    out = np.empty((h, w), np.float32)
    nfix = ctypes.c_int(0)
    check(_lib.lib().gf_stage_min_eig(device, img.ctypes.data, w, h, out.ctypes.data, ctypes.byref(nfix)))
    return out, nfix.value


def calc_optical_flow_pyr_lk(prev, nxt, prev_pts, max_level=3, init=None, device=0):
    prev, nxt = _u8(prev), _u8(nxt)
    h, w = prev.shape
    p = np.ascontiguousarray(prev_pts, np.float32).reshape(-1, 2)
    q = np.ascontiguousarray(init, np.float32).reshape(-1, 2).copy() if init is not None else np.zeros_like(p)
    st = np.zeros(len(p), np.uint8)
    check(_lib.lib().gf_stage_lk(device, prev.ctypes.data, nxt.ctypes.data, w, h, p.ctypes.data, q.ctypes.data,
                                 len(p), int(max_level), int(init is not None), st.ctypes.data))
    return q, st


def good_features_to_track(img, max_corners, min_dist, kept_pts=None, device=0):
    img = _u8(img)
    h, w = img.shape
    kp = np.zeros((0, 2), np.float32) if kept_pts is None else np.ascontiguousarray(kept_pts, np.float32).reshape(-1, 2)
    out = np.zeros((max(max_corners, 1), 2), np.float32)
    n = ctypes.c_int(0)
    info = TrackInfo()
    check(_lib.lib().gf_stage_gftt(device, img.ctypes.data, w, h, kp.ctypes.data if len(kp) else None, len(kp),
                                   int(max_corners), int(min_dist), out.ctypes.data, ctypes.byref(n), ctypes.byref(info)))
    return out[:n.value].copy(), info.as_dict()


def setmask_order(track_cnt, device=0):
    tc = np.ascontiguousarray(track_cnt, np.int32)
    perm = np.empty(len(tc), np.int32)
    check(_lib.lib().gf_stage_setmask_order(device, tc.ctypes.data, len(tc), perm.ctypes.data))
    return perm


# the class ----------------------------------------------------------------------------------------


class FeatureTracker:
    """trackImage / setPrediction / removeOutliers as in feature_tracker.h:47,70,72.

    readIntrinsicParameter is folded into the constructor (pinhole = fx fy cx cy k1 k2 p1 p2)."""

    def __init__(self, width, height, pinhole, max_cnt=150, min_dist=30, flow_back=1, depth_cam=1, device=0):
        self.L = _lib.lib()
        cfg = TrackerCfg(int(max_cnt), int(min_dist), int(flow_back), int(depth_cam), (ctypes.c_double * 8)(*pinhole))
        self._h = ctypes.c_void_p()
        check(self.L.gf_tracker_create(ctypes.byref(self._h), int(device), int(width), int(height), ctypes.byref(cfg)))
        self.width, self.height, self.max_cnt = width, height, int(max_cnt)
        self._obs = np.zeros(self.max_cnt, OBS_DTYPE)
        self._status = np.zeros(self.max_cnt, np.uint8)
        self.last_status = np.zeros(0, np.uint8)
        self.last_info = {}
        g, d = ctypes.c_void_p(), ctypes.c_void_p()
        check(self.L.gf_tracker_host_buffers(self._h, ctypes.byref(g), ctypes.byref(d)))
        self.pinned_gray = np.ctypeslib.as_array(ctypes.cast(g, ctypes.POINTER(ctypes.c_uint8)), (height, width))
        self.pinned_depth = np.ctypeslib.as_array(ctypes.cast(d, ctypes.POINTER(ctypes.c_uint16)), (height, width))

    def close(self):
        if self._h:
            self.L.gf_tracker_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _collect(self, n, info):
        self.last_info = info.as_dict()
        self.last_status = self._status[:info.n_prev].copy()
        return self._obs[:n.value].copy()

    def trackImageRaw(self, cur_time, img, depth=None):
        """Returns the structured array of gf_obs in the tracker's internal (cur_pts) order."""
        img = np.ascontiguousarray(img, np.uint8)
        assert img.shape == (self.height, self.width)
        dp, dpitch = None, 0
        if depth is not None:
            depth = np.ascontiguousarray(depth, np.uint16)
            dp, dpitch = depth.ctypes.data, depth.strides[0]
        n, info = ctypes.c_int(0), TrackInfo()
        check(self.L.gf_tracker_track(self._h, float(cur_time), img.ctypes.data, img.strides[0], dp, dpitch,
                                      self._obs.ctypes.data, ctypes.byref(n), self._status.ctypes.data, ctypes.byref(info)))
        return self._collect(n, info)

    def trackImage(self, cur_time, img, depth=None):
        """map<int, vector<pair<int, Matrix<double,8,1>>>> as {id: v[8]} (camera id is always 0)."""
        obs = self.trackImageRaw(cur_time, img, depth)
        return {int(o["id"]): o["v"].copy() for o in obs}

    def trackDevice(self, cur_time, d_gray_ptr, d_depth_ptr=None):
        n, info = ctypes.c_int(0), TrackInfo()
        check(self.L.gf_tracker_track_device(self._h, float(cur_time), d_gray_ptr, d_depth_ptr, self._obs.ctypes.data,
                                             ctypes.byref(n), self._status.ctypes.data, ctypes.byref(info)))
        return self._collect(n, info)

    @staticmethod
    def _frame(img, dtype, shape):
        if not (isinstance(img, np.ndarray) and img.dtype == dtype and img.shape == shape and img.strides[1] == img.itemsize):
            raise ValueError("frame must be a %s array of shape %s with contiguous rows" % (np.dtype(dtype).name, shape))
        return img

    def submit(self, cur_time, img, depth=None):
        """Asynchronous trackImage: up to two frames may be in flight; collect them in order with wait().
        The arrays are read by the copy engine after this call returns: keep them alive and unchanged until wait()."""
        img = self._frame(img, np.uint8, (self.height, self.width))
        if depth is not None:
            depth = self._frame(depth, np.uint16, (self.height, self.width))
        dp, dpitch = (depth.ctypes.data, depth.strides[0]) if depth is not None else (None, 0)
        check(self.L.gf_tracker_submit(self._h, float(cur_time), img.ctypes.data, img.strides[0], dp, dpitch))

    def trackBatch(self, times, gray_ptrs, depth_ptrs=None, on_device=False, gray_pitch=None, depth_pitch=None, want=True):
        """n consecutive trackImage calls in ONE library call (gf_tracker_track_batch: two frames in flight inside the
        library).  gray_ptrs / depth_ptrs: integer addresses of the frames (host, ideally pinned, or device).  Returns a
        list of (obs, status, info) per frame when `want`, else None (bench loops that only need the work done)."""
        n = len(times)
        t = np.ascontiguousarray(times, np.float64)
        g = (ctypes.c_void_p * n)(*[int(a) for a in gray_ptrs])
        d = (ctypes.c_void_p * n)(*[(int(a) if a else None) for a in depth_ptrs]) if depth_ptrs is not None else None
        obs = np.zeros((n, self.max_cnt), OBS_DTYPE) if want else None
        st = np.zeros((n, self.max_cnt), np.uint8) if want else None
        cnt = np.zeros(n, np.int32)
        info = (TrackInfo * n)()
        check(self.L.gf_tracker_track_batch(self._h, n, t.ctypes.data, g, int(gray_pitch or self.width), d,
                                            int(depth_pitch or 2 * self.width), int(bool(on_device)),
                                            obs.ctypes.data if want else None, cnt.ctypes.data, st.ctypes.data if want else None, info))
        self.last_info = info[n - 1].as_dict() if n else {}
        self.batch_infos = [info[k].as_dict() for k in range(n)]
        if not want:
            return None
        return [(obs[k, :cnt[k]].copy(), st[k, :info[k].n_prev].copy(), info[k].as_dict()) for k in range(n)]

    @staticmethod
    def trackBatchMulti(trackers, times, gray_ptrs, depth_ptrs=None, on_device=False, gray_pitch=None, depth_pitch=None, want=True):
        """gf_tracker_track_batch_multi: n frames on each of S independent trackers in one call, fed by one host thread.
        times / gray_ptrs / depth_ptrs: per tracker lists of n entries.  Returns per tracker what trackBatch returns."""
        S, n = len(trackers), len(times[0])
        a = trackers[0]
        assert all(len(x) == n for x in times) and all(tr.max_cnt == a.max_cnt and tr.width == a.width for tr in trackers)
        t = np.ascontiguousarray(times, np.float64).reshape(S * n)
        g = (ctypes.c_void_p * (S * n))(*[int(p) for row in gray_ptrs for p in row])
        d = (ctypes.c_void_p * (S * n))(*[(int(p) if p else None) for row in depth_ptrs for p in row]) if depth_ptrs is not None else None
        h = (ctypes.c_void_p * S)(*[tr._h for tr in trackers])
        obs = np.zeros((S, n, a.max_cnt), OBS_DTYPE) if want else None
        st = np.zeros((S, n, a.max_cnt), np.uint8) if want else None
        cnt = np.zeros((S, n), np.int32)
        info = (TrackInfo * (S * n))()
        check(a.L.gf_tracker_track_batch_multi(h, S, n, t.ctypes.data, g, int(gray_pitch or a.width), d, int(depth_pitch or 2 * a.width),
                                               int(bool(on_device)), obs.ctypes.data if want else None, cnt.ctypes.data,
                                               st.ctypes.data if want else None, info))
        for i, tr in enumerate(trackers):
            tr.batch_infos = [info[i * n + k].as_dict() for k in range(n)]
            tr.last_info = tr.batch_infos[-1] if n else {}
        if not want:
            return None
        return [[(obs[i, k, :cnt[i, k]].copy(), st[i, k, :info[i * n + k].n_prev].copy(), info[i * n + k].as_dict()) for k in range(n)] for i in range(S)]

    def submitDevice(self, cur_time, d_gray_ptr, d_depth_ptr=None):
        check(self.L.gf_tracker_submit_device(self._h, float(cur_time), d_gray_ptr, d_depth_ptr))

    def wait(self):
        n, info = ctypes.c_int(0), TrackInfo()
        check(self.L.gf_tracker_wait(self._h, self._obs.ctypes.data, ctypes.byref(n), self._status.ctypes.data, ctypes.byref(info)))
        return self._collect(n, info)

    def timer_start(self):
        check(self.L.gf_tracker_timer_start(self._h))

    def timer_stop(self):
        ms = ctypes.c_float(0)
        check(self.L.gf_tracker_timer_stop(self._h, ctypes.byref(ms)))
        return float(ms.value)

    def setPrediction(self, predictPts):
        ids = np.array(sorted(predictPts.keys()), np.int32)
        xyz = np.array([predictPts[int(i)] for i in ids], np.float64).reshape(-1, 3)
        check(self.L.gf_tracker_set_prediction(self._h, ids.ctypes.data if len(ids) else None,
                                               xyz.ctypes.data if len(ids) else None, len(ids)))

    def removeOutliers(self, removePtsIds):
        ids = np.array(sorted(removePtsIds), np.int32)
        check(self.L.gf_tracker_remove_ids(self._h, ids.ctypes.data if len(ids) else None, len(ids)))

    def set_profiling(self, on=True):
        check(self.L.gf_tracker_set_profiling(self._h, int(bool(on))))

    STAGES = ("upload", "pyramid", "lk", "setmask", "gftt_select", "finalize", "download", "min_eig_aux")

    def last_stage_ms(self):
        ms = (ctypes.c_float * 8)()
        check(self.L.gf_tracker_last_stage_ms(self._h, ms))
        return dict(zip(self.STAGES, [float(v) for v in ms]))

    def last_device_ms(self):
        ms = ctypes.c_float(0)
        check(self.L.gf_tracker_last_device_ms(self._h, ctypes.byref(ms)))
        return ms.value
