"""Host-side mirror of the reference's Estimator::optimization() (estimator.h:147, estimator.cpp:2890-3636)
over the C ABI (gf_ba_*): the ceres::Solve it performs runs on the GPU; there is no CPU fallback.
"""
import ctypes

from . import _lib
from ._lib import BaProblem, BaSummary, check


class BundleAdjuster:
    """One solver workspace bound to one GPU (gf_ba).  optimization(problem) updates the para_* arrays of
    `problem` (ground_fusion_b200.ba_problem.Problem) in place, like Estimator::optimization() does with
    its para_Pose / para_SpeedBias / para_Feature members, and returns the solver summary."""

    def __init__(self, device=0):
        self.L = _lib.lib()
        self.L.gf_ba_create.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]
        self.L.gf_ba_destroy.argtypes = [ctypes.c_void_p]
        self.L.gf_ba_destroy.restype = None
        self.L.gf_ba_solve.argtypes = [ctypes.c_void_p, ctypes.POINTER(BaProblem), ctypes.POINTER(BaSummary)]
        self._h = ctypes.c_void_p()
        check(self.L.gf_ba_create(ctypes.byref(self._h), int(device)))

    def optimization(self, problem):
        p = problem.struct()
        s = BaSummary()
        check(self.L.gf_ba_solve(self._h, ctypes.byref(p), ctypes.byref(s)))
        return s.as_dict()

    def solve_struct(self, p_struct):
        """Same, for a pre-built ctypes gf_ba_problem (avoids rebuilding it in timing loops)."""
        s = BaSummary()
        check(self.L.gf_ba_solve(self._h, ctypes.byref(p_struct), ctypes.byref(s)))
        return s

    def close(self):
        if self._h:
            self.L.gf_ba_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
