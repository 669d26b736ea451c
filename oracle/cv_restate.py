"""ctypes loader for oracle/fe_cv_restate.c -- TEST INFRASTRUCTURE ONLY.

The C file restates OpenCV's pyrDown / calcOpticalFlowPyrLK / goodFeaturesToTrack (the three
upstream calls made by the reference's feature_tracker.cpp:118-153,198).  Only tests/, smoke() and
bench.py's cpu_baseline leg may import this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libgf_oracle_fe.so")
_LIB = None

ORDER_SEQUENTIAL = 0
ORDER_SSE_HADD = 1
ORDER_SSE_MOVEHL = 2
# pinned against cv2 4.13.0 (tests/test_fe_oracle.py)
ORDER_CV2 = ORDER_SSE_MOVEHL
EIG_VARIANT_CV2 = 0


def build(force=False):
    src = os.path.join(_HERE, "fe_cv_restate.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-msse2", "-mfpmath=sse", "-shared",
                               "-fPIC", "-o", _SO, src, "-lm"])
    return _SO


def lib():
    global _LIB
    if _LIB is None:
        build()
        L = ctypes.CDLL(_SO)
        vp, i, d = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
        L.gfo_pyr_down_u8.argtypes = [vp, i, i, vp]
        L.gfo_pyr_down_u8.restype = None
        L.gfo_lk.argtypes = [vp, vp, i, i, vp, vp, i, i, i, vp, i]
        L.gfo_lk.restype = None
        L.gfo_min_eig.argtypes = [vp, i, i, vp, i]
        L.gfo_min_eig.restype = None
        L.gfo_gftt.argtypes = [vp, i, i, vp, i, d, d, vp, i]
        L.gfo_gftt.restype = i
        _LIB = L
    return _LIB


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    assert a.ndim == 2
    return a


def pyr_down(img):
    img = _u8(img)
    h, w = img.shape
    out = np.empty(((h + 1) // 2, (w + 1) // 2), np.uint8)
    lib().gfo_pyr_down_u8(img.ctypes.data, w, h, out.ctypes.data)
    return out


def lk(prev, nxt, prev_pts, max_level, init=None, order=ORDER_CV2):
    """Returns (next_pts [n,2] float32, status [n] uint8)."""
    prev, nxt = _u8(prev), _u8(nxt)
    h, w = prev.shape
    p = np.ascontiguousarray(prev_pts, np.float32).reshape(-1, 2)
    n = len(p)
    q = (np.ascontiguousarray(init, np.float32).reshape(-1, 2).copy() if init is not None
         else np.zeros_like(p))
    st = np.zeros(n, np.uint8)
    lib().gfo_lk(prev.ctypes.data, nxt.ctypes.data, w, h, p.ctypes.data, q.ctypes.data, n,
                 int(max_level), int(init is not None), st.ctypes.data, int(order))
    return q, st


def min_eig(img, variant=EIG_VARIANT_CV2):
    img = _u8(img)
    h, w = img.shape
    out = np.empty((h, w), np.float32)
    lib().gfo_min_eig(img.ctypes.data, w, h, out.ctypes.data, int(variant))
    return out


def gftt(img, max_corners, quality, min_distance, mask=None, variant=EIG_VARIANT_CV2):
    img = _u8(img)
    h, w = img.shape
    cap = max(int(max_corners), 1) if max_corners > 0 else w * h
    out = np.zeros((cap, 2), np.float32)
    m = None
    if mask is not None:
        m = _u8(mask)
    n = lib().gfo_gftt(img.ctypes.data, w, h, m.ctypes.data if m is not None else None,
                       int(max_corners), float(quality), float(min_distance), out.ctypes.data,
                       int(variant))
    return out[:n].copy()
