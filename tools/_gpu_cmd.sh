timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python - <<'PY'
import sys, time; sys.path.insert(0, '.')
from ground_fusion_b200.estimator import BundleAdjuster
from ground_fusion_b200.synth_ba import make_window
ba = BundleAdjuster(0)
pb, _ = make_window(seed=100)
ba.optimization(pb)
for k in range(3):
    pr = ba.marginalize_old(pb); print("marg device ms %.3f" % ba.last_marg_ms)
PY
