"""Generates tests/golden/fe_*.npz from the cv2-based oracle (oracle/fe_oracle.py).

The reference holds no golden vectors for the front end, so parity is pinned by
these fixtures: per frame the id list (in tracker order), the combined LK status vector ("inlier
mask"), the new-corner coordinates, and the 8-vector of every feature.  Run from the repo root:
    python tests/golden/make_fe_golden.py
cv2 version and CPU features are recorded because the eig map bits depend on OpenCV's dispatch path.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import cv2  # noqa: E402

from ground_fusion_b200.synth import SyntheticStream  # noqa: E402
from oracle.fe_oracle import IDC_CAM, FeatureTrackerOracle, PinholeCamera  # noqa: E402

CASES = {
    "fe_c2_seed0": dict(seed=0, w=640, h=480, max_cnt=150, min_dist=30, frames=40),
    "fe_c3_seed1": dict(seed=1, w=640, h=480, max_cnt=300, min_dist=20, frames=16),
}


def run_case(c):
    sc = c["w"] / 640.0
    cam = PinholeCamera(IDC_CAM["fx"] * sc, IDC_CAM["fy"] * sc, IDC_CAM["cx"] * sc, IDC_CAM["cy"] * sc,
                        IDC_CAM["k1"], IDC_CAM["k2"], IDC_CAM["p1"], IDC_CAM["p2"])
    st = SyntheticStream(seed=c["seed"], width=c["w"], height=c["h"])
    ft = FeatureTrackerOracle(cam, c["max_cnt"], c["min_dist"], 1, 1)
    out = {}
    h = hashlib.sha256()
    for k in range(c["frames"]):
        t, g, d = st.frame(k)
        h.update(g.tobytes()); h.update(d.tobytes())
        ff = ft.trackImage(t, g, d)
        out["ids_%d" % k] = np.array(ft.ids, np.int32)
        out["cnt_%d" % k] = np.array(ft.track_cnt, np.int32)
        out["status_%d" % k] = ft.last_status.astype(np.uint8)
        out["npts_%d" % k] = ft.last_n_pts.astype(np.float32)
        out["obs_%d" % k] = np.array([ff[i] for i in ft.ids], np.float64).reshape(-1, 8)
    out["frames_sha256"] = np.frombuffer(h.digest(), np.uint8)
    out["meta"] = np.array([c["seed"], c["w"], c["h"], c["max_cnt"], c["min_dist"], c["frames"]], np.int64)
    out["cv2_version"] = np.array(cv2.__version__)
    return out


if __name__ == "__main__":
    for name, c in CASES.items():
        np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), name + ".npz"), **run_case(c))
        print("wrote", name)
