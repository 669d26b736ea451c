"""Host-side mirror of the reference's Estimator::optimization() (estimator.h:147, estimator.cpp:2890-3636)
over the C ABI (gf_ba_*): the ceres::Solve it performs runs on the GPU; there is no CPU fallback.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import BaPrior, BaProblem, BaSummary, check


def double2vector(problem, R0_before, P0_before, use_imu=True):
    """Estimator::double2vector (estimator.cpp:2440-2494): Rs [F,3,3], Ps [F,3], Vs [F,3] from the solved para_* arrays,
    with frame 0's yaw and position restored to their values before the solve.  Host-only (no GPU needed)."""
    L = _lib.lib()
    dp = ctypes.POINTER(ctypes.c_double)
    L.gf_ba_double2vector.argtypes = [ctypes.POINTER(BaProblem), dp, dp, ctypes.c_int, dp, dp, dp]
    F = problem.n_frames
    R0 = np.ascontiguousarray(R0_before, np.float64).reshape(3, 3); P0 = np.ascontiguousarray(P0_before, np.float64).reshape(3)
    Rs, Ps, Vs = np.zeros((F, 3, 3)), np.zeros((F, 3)), np.zeros((F, 3))
    p = problem.struct()
    check(L.gf_ba_double2vector(ctypes.byref(p), R0.ctypes.data_as(dp), P0.ctypes.data_as(dp), 1 if use_imu else 0,
                                Rs.ctypes.data_as(dp), Ps.ctypes.data_as(dp), Vs.ctypes.data_as(dp)))
    return Rs, Ps, Vs


def spd_solve(A, b, device=0, tile_cap=-1):
    """x = A^-1 b through the back end's own dense solver (gf_stage_spd_solve: 8x8-tile Cholesky on FP64 tensor-core MMAs)."""
    L = _lib.lib()
    dp = ctypes.POINTER(ctypes.c_double)
    L.gf_stage_spd_solve.argtypes = [ctypes.c_int, dp, dp, ctypes.c_int, dp, ctypes.c_int]
    A = np.ascontiguousarray(A, np.float64); b = np.ascontiguousarray(b, np.float64)
    n = len(b)
    assert A.shape == (n, n)
    x = np.zeros(n)
    check(L.gf_stage_spd_solve(int(device), A.ctypes.data_as(dp), b.ctypes.data_as(dp), n, x.ctypes.data_as(dp), int(tile_cap)))
    return x


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

    def _marginalize(self, problem, fn_name):
        from .ba_problem import Prior
        dp = ctypes.POINTER(ctypes.c_double)
        fn = getattr(self.L, fn_name)
        fn.argtypes = [ctypes.c_void_p, ctypes.POINTER(BaProblem), ctypes.POINTER(BaPrior), dp, dp, dp, ctypes.POINTER(ctypes.c_float)]
        p = problem.struct()
        cap = 16 * problem.n_frames + 24
        x0 = np.zeros(cap); J = np.zeros(cap * cap); r = np.zeros(cap)
        out = BaPrior(); ms = ctypes.c_float(0)
        n = fn(self._h, ctypes.byref(p), ctypes.byref(out), x0.ctypes.data_as(dp), J.ctypes.data_as(dp), r.ctypes.data_as(dp), ctypes.byref(ms))
        if n < 0:
            check(n)
        if n == 0:
            return None
        self.last_marg_ms = float(ms.value)
        nb = out.n_blocks
        return Prior(list(out.block_kind)[:nb], list(out.block_index)[:nb], list(out.block_idx)[:nb], x0.copy(),
                     J[:n * n].reshape(n, n).copy(), r[:n].copy())

    def marginalize_old(self, problem):
        """MARGIN_OLD on the GPU (estimator.cpp:3334-3535): returns the ba_problem.Prior for the next window (block indices
        already shifted by one frame).  last_marg_ms holds the CUDA-event time."""
        return self._marginalize(problem, "gf_ba_marginalize_old")

    def marginalize_second_new(self, problem):
        """MARGIN_SECOND_NEW on the GPU (estimator.cpp:3536-3631): the last prior with para_Pose[WINDOW_SIZE - 1] marginalised;
        None when the prior does not hold that pose (the reference then keeps the prior as it is)."""
        return self._marginalize(problem, "gf_ba_marginalize_second_new")

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
