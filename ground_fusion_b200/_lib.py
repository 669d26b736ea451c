"""ctypes binding of libgf_b200.so (the C ABI declared in include/gf_b200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is visible every entry
point raises.  The library is built in-tree by `make` / `__graft_entry__.build()`.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgf_b200.so")
_LIB = None


class GfError(RuntimeError):
    pass


class TrackerCfg(ctypes.Structure):
    _fields_ = [("max_cnt", ctypes.c_int), ("min_dist", ctypes.c_int), ("flow_back", ctypes.c_int),
                ("depth_cam", ctypes.c_int), ("pinhole", ctypes.c_double * 8)]


class Obs(ctypes.Structure):
    _fields_ = [("id", ctypes.c_int32), ("track_cnt", ctypes.c_int32), ("v", ctypes.c_double * 8)]


OBS_DTYPE = np.dtype([("id", np.int32), ("track_cnt", np.int32), ("v", np.float64, (8,))])
assert OBS_DTYPE.itemsize == ctypes.sizeof(Obs) == 72


class TrackInfo(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("n_prev", "n_tracked", "n_kept", "n_new", "n_candidates",
                                                "nms_rounds", "eig_fixups", "reserved")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "reserved"}


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise GfError("libgf_b200.so not built (%s): run `make` or __graft_entry__.build(); "
                          "there is no CPU fallback" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        vp, i, d, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_size_t
        L.gf_last_error.restype = ctypes.c_char_p
        L.gf_version.restype = ctypes.c_char_p
        L.gf_kernel_launch_count.restype = ctypes.c_uint64
        L.gf_tracker_create.argtypes = [ctypes.POINTER(vp), i, i, i, ctypes.POINTER(TrackerCfg)]
        L.gf_tracker_destroy.argtypes = [vp]
        L.gf_tracker_destroy.restype = None
        L.gf_tracker_host_buffers.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(vp)]
        L.gf_tracker_track.argtypes = [vp, d, vp, sz, vp, sz, vp, ctypes.POINTER(i), vp, ctypes.POINTER(TrackInfo)]
        L.gf_tracker_submit.argtypes = [vp, d, vp, sz, vp, sz]
        L.gf_tracker_wait.argtypes = [vp, vp, ctypes.POINTER(i), vp, ctypes.POINTER(TrackInfo)]
        L.gf_tracker_track_device.argtypes = [vp, d, vp, vp, vp, ctypes.POINTER(i), vp, ctypes.POINTER(TrackInfo)]
        L.gf_tracker_set_prediction.argtypes = [vp, vp, vp, i]
        L.gf_tracker_remove_ids.argtypes = [vp, vp, i]
        L.gf_tracker_last_device_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_float)]
        L.gf_stage_pyr_down.argtypes = [i, vp, i, i, vp]
        L.gf_stage_min_eig.argtypes = [i, vp, i, i, vp, ctypes.POINTER(i)]
        L.gf_stage_lk.argtypes = [i, vp, vp, i, i, vp, vp, i, i, i, vp]
        L.gf_stage_gftt.argtypes = [i, vp, i, i, vp, i, i, i, vp, ctypes.POINTER(i), ctypes.POINTER(TrackInfo)]
        L.gf_stage_setmask_order.argtypes = [i, vp, i, vp]
        _LIB = L
    return _LIB


def check(rc):
    if rc != 0:
        raise GfError("libgf_b200 error %d: %s" % (rc, lib().gf_last_error().decode()))
