"""Small driver for ncu: a few gf_ba_solve calls on C2 windows."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ground_fusion_b200.estimator import BundleAdjuster
from ground_fusion_b200.synth_ba import make_window
ba = BundleAdjuster(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for k in range(n):
    pb, _ = make_window(seed=100 + k)
    print(ba.optimization(pb)["device_ms"])

pb, _ = make_window(seed=100)
ba.optimization(pb)
for _ in range(2):
    ba.marginalize_old(pb)
print('marginalize_old device ms', ba.last_marg_ms)

import ctypes
prof = (ctypes.c_longlong * 32)()
ba.L.gf_ba_debug_profile.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_longlong)]
ba.L.gf_ba_debug_profile(ba._h, prof)
names = ["adopt+bookkeeping", "-", "wait for the TMA load", "cholesky", "backsubst + finite check", "landmark backsubst", "dogleg+model", "candidate", "-", "-", "-", "-", "-", "eval: prior CTA max", "eval: visual CTA max", "eval: imu CTA max"]
tot = sum(list(prof)[:9])
for n_, v in zip(names, prof):
    print("%-20s %9d cycles %5.1f%%" % (n_, v, 100.0 * v / max(tot, 1)))

print("k_ba_step total cycles", prof[30], "launches", prof[31])
