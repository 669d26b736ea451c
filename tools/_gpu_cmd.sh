timeout 600 python -m pytest tests/test_ba_gpu.py -x -q -s 2>&1 | tail -6
timeout 200 python - <<'PY'
import sys, time; sys.path.insert(0, '.')
from ground_fusion_b200.estimator import BundleAdjuster
from ground_fusion_b200.synth_ba import make_window
ba = BundleAdjuster(0)
pb, _ = make_window(seed=100)
ba.optimization(pb)
for k in range(4):
    t = time.perf_counter(); pr = ba.marginalize_old(pb); print("marg device ms %.3f wall ms %.3f n %d" % (ba.last_marg_ms, 1e3 * (time.perf_counter() - t), pr.n))
PY
