"""Small driver for ncu: a few trackImage calls on the C2 stream."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ground_fusion_b200.feature_tracker import FeatureTracker
from ground_fusion_b200.synth import SyntheticStream
from ground_fusion_b200.synth import idc_params8
st = SyntheticStream(seed=0)
tr = FeatureTracker(640, 480, idc_params8(), 150, 30, 1, 1)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
for k in range(n):
    t, g, d = st.frame(k)
    tr.trackImageRaw(t, g, d)
    print(k, tr.last_device_ms(), tr.last_info)
