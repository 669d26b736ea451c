"""oracle/replay_adapters.py -- TEST INFRASTRUCTURE ONLY: the CPU oracles behind the interfaces ground_fusion_b200.replay expects,
so that tests / bench.py can run the reference pipeline (OpenCV front end + FeatureManager + CPU solver) through the same loop."""
from oracle import ba_oracle as O


class OracleBA:
    def optimization(self, pb):
        return O.solve(pb)

    def marginalize_old(self, pb):
        return O.marginalize_old(pb)

    def marginalize_second_new(self, pb):
        return O.marginalize_second_new(pb)


def oracle_components(cam_kwargs, max_cnt=150, min_dist=30, depth_threshold=3.0, fast=True):
    from oracle.fe_oracle import FeatureTrackerOracle, FeatureTrackerOracleFast, PinholeCamera
    from oracle.fm_oracle import FeatureManagerOracle
    cls = FeatureTrackerOracleFast if fast else FeatureTrackerOracle
    return cls(PinholeCamera(**cam_kwargs), max_cnt, min_dist, 1, 1), FeatureManagerOracle(depth_threshold=depth_threshold), OracleBA()
