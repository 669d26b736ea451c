"""oracle/fm_oracle.py -- TEST INFRASTRUCTURE ONLY (CPU oracle of FeatureManager).

Plain-Python / NumPy restatement of the reference's FeatureManager
(/root/reference/vins_estimator/src/estimator/feature_manager.cpp; class layout feature_manager.h:30-106), written
independently of ground_fusion_b200/feature_manager.py (whose per-landmark arithmetic runs in CUDA).  The reference holds
no tests for this class: parity is pinned by geometric properties (tests/test_fm_oracle.py) and by the GPU mirror agreeing
with this restatement on seeded inputs (tests/test_fm_gpu.py).
"""
import numpy as np

INIT_DEPTH = 5.0            # parameters.cpp:478
FOCAL_LENGTH = 600.0        # parameters.h:23


class FeaturePerFrame:      # feature_manager.h:30-64
    def __init__(self, v8, td):
        self.point = np.array(v8[0:3], float); self.uv = np.array(v8[3:5], float); self.velocity = np.array(v8[5:7], float)
        self.depth = float(v8[7]); self.cur_td = float(td)


class FeaturePerId:         # feature_manager.h:66-88
    def __init__(self, feature_id, start_frame):
        self.feature_id, self.start_frame = int(feature_id), int(start_frame)
        self.feature_per_frame = []
        self.used_num = 0; self.estimated_depth = -1.0; self.estimate_flag = 0; self.solve_flag = 0

    def endFrame(self):
        return self.start_frame + len(self.feature_per_frame) - 1


class FeatureManagerOracle:
    def __init__(self, min_parallax_px=10.0, depth_threshold=3.0, window_size=10):
        self.feature = []
        self.MIN_PARALLAX = min_parallax_px / FOCAL_LENGTH       # parameters.cpp:345-346
        self.depth_threshold = depth_threshold                    # parameters.cpp:172
        self.WINDOW_SIZE = window_size
        self.last_track_num = self.new_feature_num = self.long_track_num = 0
        self.last_average_parallax = 0.0

    # feature_manager.cpp:43-55
    def getFeatureCount(self):
        cnt = 0
        for it in self.feature:
            it.used_num = len(it.feature_per_frame)
            if it.used_num >= 4:
                cnt += 1
        return cnt

    # feature_manager.cpp:57-116
    def addFeatureCheckParallax(self, frame_count, image, td):
        parallax_sum, parallax_num = 0.0, 0
        self.last_track_num = self.new_feature_num = self.long_track_num = 0
        self.last_average_parallax = 0.0
        index = {it.feature_id: it for it in self.feature}
        for fid in sorted(image):                      # std::map iterates by id
            f = FeaturePerFrame(image[fid], td)
            it = index.get(fid)
            if it is None:
                it = FeaturePerId(fid, frame_count); self.feature.append(it); index[fid] = it
                it.feature_per_frame.append(f); self.new_feature_num += 1
            else:
                it.feature_per_frame.append(f); self.last_track_num += 1
                if len(it.feature_per_frame) >= 4:
                    self.long_track_num += 1
        if frame_count < 2 or self.last_track_num < 20 or self.long_track_num < 40 or self.new_feature_num > 0.5 * self.last_track_num:
            return True
        for it in self.feature:
            if it.start_frame <= frame_count - 2 and it.start_frame + len(it.feature_per_frame) - 1 >= frame_count - 1:
                parallax_sum += self.compensatedParallax2(it, frame_count); parallax_num += 1
        if parallax_num == 0:
            return True
        self.last_average_parallax = parallax_sum / parallax_num * FOCAL_LENGTH
        return parallax_sum / parallax_num >= self.MIN_PARALLAX

    # feature_manager.cpp:977-1011
    @staticmethod
    def compensatedParallax2(it, frame_count):
        fi = it.feature_per_frame[frame_count - 2 - it.start_frame]
        fj = it.feature_per_frame[frame_count - 1 - it.start_frame]
        p_j, p_i = fj.point, fi.point
        u_j, v_j = p_j[0], p_j[1]
        dep_i = p_i[2]
        u_i, v_i = p_i[0] / dep_i, p_i[1] / dep_i
        du, dv = u_i - u_j, v_i - v_j
        du_comp, dv_comp = du, dv                       # p_i_comp = p_i (:989)
        return max(0.0, float(np.sqrt(min(du * du + dv * dv, du_comp * du_comp + dv_comp * dv_comp))))

    def iter_ba_features(self):
        """estimator.cpp:3268-3297: the landmarks optimization() builds factors for."""
        for it in self.feature:
            it.used_num = len(it.feature_per_frame)
            if it.used_num >= 4:
                yield it.start_frame, [(f.point, f.velocity, f.cur_td) for f in it.feature_per_frame], it.estimated_depth, it.estimate_flag

    def triangulateAll(self, frameCnt, Ps, Rs, tic, ric):
        self.triangulateWithDepth(frameCnt, Ps, Rs, tic, ric)      # estimator.cpp:1090-1102
        self.triangulate(frameCnt, Ps, Rs, tic, ric)

    # feature_manager.cpp:286-302, 249-267, 269-278
    def getDepthVector(self):
        out = []
        for it in self.feature:
            it.used_num = len(it.feature_per_frame)
            if it.used_num >= 4:
                out.append(1.0 / it.estimated_depth)
        return np.array(out)

    def setDepth(self, x):
        k = -1
        for it in self.feature:
            it.used_num = len(it.feature_per_frame)
            if it.used_num < 4:
                continue
            k += 1
            it.estimated_depth = 1.0 / x[k]
            it.solve_flag = 2 if it.estimated_depth < 0 else 1

    def removeFailures(self):
        self.feature = [it for it in self.feature if it.solve_flag != 2]

    # feature_manager.cpp:726-799
    def triangulateWithDepth(self, frameCnt, Ps, Rs, tic, ric):
        for it in self.feature:
            it.used_num = len(it.feature_per_frame)
            if it.used_num < 4 or it.estimated_depth > 0:
                continue
            s0 = it.start_frame
            verified = []
            tr = Ps[s0] + Rs[s0] @ tic; Rr = Rs[s0] @ ric
            for i, fi in enumerate(it.feature_per_frame):
                t0 = Ps[s0 + i] + Rs[s0 + i] @ tic; R0 = Rs[s0 + i] @ ric
                if fi.depth < 0.1 or fi.depth > self.depth_threshold:
                    continue
                point0 = fi.point * fi.depth
                t2r = Rr.T @ (t0 - tr); R2r = Rr.T @ R0
                for j, fj in enumerate(it.feature_per_frame):
                    if i == j:
                        continue
                    t1 = Ps[s0 + j] + Rs[s0 + j] @ tic; R1 = Rs[s0 + j] @ ric
                    t20 = R0.T @ (t1 - t0); R20 = R0.T @ R1
                    pp = R20.T @ point0 - R20.T @ t20
                    res = fj.point[:2] - np.array([pp[0] / pp[2], pp[1] / pp[2]])
                    if np.linalg.norm(res) < 10.0 / 460:
                        verified.append((R2r @ point0 + t2r)[2])
            if not verified:
                continue
            it.estimated_depth = sum(verified) / len(verified); it.estimate_flag = 1
            if it.estimated_depth < 0.1:
                it.estimated_depth = INIT_DEPTH; it.estimate_flag = 0

    # feature_manager.cpp:668-723
    def triangulate(self, frameCnt, Ps, Rs, tic, ric):
        for it in self.feature:
            if it.estimated_depth > 0:
                continue
            it.used_num = len(it.feature_per_frame)
            if it.used_num < 4:
                continue
            imu_i = it.start_frame
            A = np.zeros((2 * len(it.feature_per_frame), 4))
            t0 = Ps[imu_i] + Rs[imu_i] @ tic; R0 = Rs[imu_i] @ ric
            for k, fpf in enumerate(it.feature_per_frame):
                imu_j = imu_i + k
                t1 = Ps[imu_j] + Rs[imu_j] @ tic; R1 = Rs[imu_j] @ ric
                t = R0.T @ (t1 - t0); R = R0.T @ R1
                P = np.zeros((3, 4)); P[:, :3] = R.T; P[:, 3] = -R.T @ t
                f = fpf.point / np.linalg.norm(fpf.point)
                A[2 * k] = f[0] * P[2] - f[2] * P[0]
                A[2 * k + 1] = f[1] * P[2] - f[2] * P[1]
            V = np.linalg.svd(A, full_matrices=False)[2][-1]      # JacobiSVD(...).matrixV().rightCols<1>()
            it.estimated_depth = V[2] / V[3]; it.estimate_flag = 2
            if it.estimated_depth < 0.1:
                it.estimated_depth = INIT_DEPTH; it.estimate_flag = 0

    # feature_manager.cpp:801-816
    # ---- the estimator's per-landmark loops after optimization() (estimator.cpp:3853-4011), with 4x4 homogeneous transforms ----
    @staticmethod
    def _T(R, P):
        T = np.eye(4); T[:3, :3] = R; T[:3, 3] = P
        return T

    def _errors(self, it, Ps, Rs, tic, ric):
        """(sum of reprojectionError, sum of reprojectionError3D, count) of a landmark's first observation into the later ones."""
        Tbc = self._T(ric, tic)
        i = it.start_frame
        d = it.estimated_depth
        pw = self._T(Rs[i], Ps[i]) @ Tbc @ np.append(d * it.feature_per_frame[0].point, 1.0)
        e2 = e3 = 0.0; cnt = 0
        for k, fr in enumerate(it.feature_per_frame):
            if k == 0:
                continue
            j = i + k
            pc = (np.linalg.inv(self._T(Rs[j], Ps[j]) @ Tbc) @ pw)[:3]
            e2 += float(np.hypot(pc[0] / pc[2] - fr.point[0], pc[1] / pc[2] - fr.point[1]))
            e3 += float(np.linalg.norm(pc - fr.point)) / d
            cnt += 1
        return e2, e3, cnt

    def outliersRejection(self, Ps, Rs, tic, ric):
        out = set()
        for it in self.feature:
            it.used_num = len(it.feature_per_frame)
            if it.used_num < 4:
                continue
            e2, _, cnt = self._errors(it, Ps, Rs, tic, ric)
            if e2 / cnt * FOCAL_LENGTH > 3:
                out.add(it.feature_id)
        return out

    def movingConsistencyCheckW(self, Ps, Rs, tic, ric):
        out = set()
        for it in self.feature:
            it.used_num = len(it.feature_per_frame)
            if not (it.used_num >= 2 and it.start_frame < self.WINDOW_SIZE - 2) or it.estimated_depth < 0:
                continue
            e2, e3, cnt = self._errors(it, Ps, Rs, tic, ric)
            if cnt > 0 and (FOCAL_LENGTH * e2 / cnt > 10 or e3 / cnt > 2.0):
                out.add(it.feature_id)
        return out

    def predictPtsInNextFrame(self, frame_count, Ps, Rs, tic, ric):
        if frame_count < 2:
            return {}
        cur, prev = self._T(Rs[frame_count], Ps[frame_count]), self._T(Rs[frame_count - 1], Ps[frame_count - 1])
        nxt = cur @ (np.linalg.inv(prev) @ cur)
        Tbc = self._T(ric, tic)
        out = {}
        for it in self.feature:
            if it.estimated_depth > 0 and len(it.feature_per_frame) >= 2 and it.start_frame + len(it.feature_per_frame) - 1 == frame_count:
                i = it.start_frame
                pw = self._T(Rs[i], Ps[i]) @ Tbc @ np.append(it.estimated_depth * it.feature_per_frame[0].point, 1.0)
                out[it.feature_id] = (np.linalg.inv(nxt @ Tbc) @ pw)[:3]
        return out

    def removeOutlier(self, outlierIndex):
        self.feature = [it for it in self.feature if it.feature_id not in outlierIndex]

    # feature_manager.cpp:818-856
    def removeBackShiftDepth(self, marg_R, marg_P, new_R, new_P):
        keep = []
        for it in self.feature:
            if it.start_frame != 0:
                it.start_frame -= 1; keep.append(it); continue
            uv_i = it.feature_per_frame[0].point
            del it.feature_per_frame[0]
            if len(it.feature_per_frame) < 2:
                continue
            pts_i = uv_i * it.estimated_depth
            w_pts_i = marg_R @ pts_i + marg_P
            pts_j = new_R.T @ (w_pts_i - new_P)
            it.estimated_depth = pts_j[2] if pts_j[2] > 0 else INIT_DEPTH
            keep.append(it)
        self.feature = keep

    # feature_manager.cpp:858-874
    def removeBack(self):
        keep = []
        for it in self.feature:
            if it.start_frame != 0:
                it.start_frame -= 1; keep.append(it)
            else:
                del it.feature_per_frame[0]
                if len(it.feature_per_frame):
                    keep.append(it)
        self.feature = keep

    # feature_manager.cpp:913-931
    def removeFront(self, frame_count):
        keep = []
        for it in self.feature:
            if it.start_frame == frame_count:
                it.start_frame -= 1; keep.append(it); continue
            j = self.WINDOW_SIZE - 1 - it.start_frame
            if it.endFrame() < frame_count - 1:
                keep.append(it); continue
            del it.feature_per_frame[j]
            if len(it.feature_per_frame):
                keep.append(it)
        self.feature = keep
