import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ground_fusion_b200.estimator import spd_solve
for n in (5, 8, 9, 17, 64, 165):
    rng = np.random.default_rng(n)
    B = rng.standard_normal((n, n + 5)); A = B @ B.T + n * np.eye(n); b = rng.standard_normal(n)
    try:
        x = spd_solve(A, b); w = np.linalg.solve(A, b)
        print(n, "err", np.abs(x - w).max() / np.abs(w).max(), "first bad idx", np.argmax(np.abs(x - w) > 1e-9 * np.abs(w).max()))
    except Exception as e:
        print(n, "EXC", e)
