"""Generates tests/golden/ba_*.npz from the CPU oracle (oracle/ba_oracle.c).

The reference holds no test vectors for Estimator::optimization() and Ceres is not installable here,
so these fixtures do not pin the oracle against the reference -- the oracle stays "parity unpinned" in that sense.  What they
pin is the oracle (and the CUDA solver) against regressions: for each seeded synthetic window the solver summary, the solved
parameter blocks, and the gauge-invariant parts of the MARGIN_OLD prior.  Run from the repo root:
    python tests/golden/make_ba_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ground_fusion_b200.synth_ba import make_window  # noqa: E402
from oracle import ba_oracle as O  # noqa: E402

CASES = {
    "ba_c2_seed0": dict(seed=0),
    "ba_c3_wheel_seed1": dict(seed=1, with_wheel=True),
    "ba_plane_seed2": dict(seed=2, with_plane=True),
}


def run_case(kw):
    pb, _ = make_window(**kw)
    s = O.solve(pb)
    out = dict(cost=np.array(s["cost"]), radius=np.array(s["radius"]), iterations=np.int32(s["iterations"]),
               termination=np.int32(s["termination"]), reduced_dim=np.int32(s["reduced_dim"]),
               para_pose=pb.para_pose.copy(), para_speed_bias=pb.para_speed_bias.copy(), para_feature=pb.para_feature.copy(),
               para_ex_pose=pb.para_ex_pose.copy(), para_td=pb.para_td.copy(), para_ex_wheel=pb.para_ex_wheel.copy(),
               para_ix_wheel=pb.para_ix_wheel.copy(), para_td_wheel=pb.para_td_wheel.copy(), para_plane_R=pb.para_plane_R.copy(),
               para_plane_Z=pb.para_plane_Z.copy())
    if not kw.get("with_plane"):
        pr = O.marginalize_old(pb)
        out.update(prior_kinds=np.array(pr.kinds, np.int32), prior_indices=np.array(pr.indices, np.int32), prior_idx=np.array(pr.idx, np.int32),
                   prior_H=pr.J.T @ pr.J, prior_b=pr.J.T @ pr.r)
    return out


if __name__ == "__main__":
    for name, kw in CASES.items():
        np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), name + ".npz"), **run_case(kw))
        print("wrote", name)
