#!/usr/bin/env python
"""bench.py -- headline benchmark of the two hot paths (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one FeatureTracker::trackImage call on one 640x480 RGB-D frame of a synthetic stream with
150 features (BASELINE.json configs[1], "C2").  Rank r tracks its own stream (seed r): weak scaling, no
data-path collective; the timed region is bracketed by barrier + synchronize and the
maximum over ranks is reported.

  value      frames/s with the frame already resident in HBM (gf_tracker_track_device)
  e2e        frames/s through gf_tracker_track with pinned HOST buffers (H2D of gray+depth and D2H of
             the observations inside the timed region)
  roofline   of the dominant kernel (k_track: forward+reverse pyramidal LK), algorithmic bytes / duration
  cpu_baseline  the cv2-based oracle of the same call on this box's host cores (bounded sample)
  ba         (when the back end is built) sliding-window solves/s next to the CPU oracle

--impl reference times the reference's CPU path: the reference cannot be compiled here (ROS/Eigen/Ceres/
OpenCV C++ absent), so this is the line-by-line restatement on the same three OpenCV entry points
(oracle/fe_oracle.py; kind "port").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, MAX_CNT, MIN_DIST = 640, 480, 150, 30
RING = 160            # distinct frames per stream: 160 x 0.92 MB = 147 MB > 126 MB L2


def tri(k, n):
    """ping-pong index so the ring stays temporally coherent"""
    p = k % (2 * n - 2)
    return p if p < n else 2 * n - 2 - p


class ClockSampler(threading.Thread):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[1]) for r in self.rows)
        reasons = []
        for i, name in ((4, "hw_slowdown"), (5, "hw_thermal_slowdown"), (6, "sw_thermal_slowdown"), (7, "sw_power_cap")):
            if any(r[i].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][2]), "reasons": reasons, "samples": len(self.rows)}


def make_frames(seed, n):
    from ground_fusion_b200.synth import SyntheticStream
    st = SyntheticStream(seed=seed, width=W, height=H)
    gray = np.empty((n, H, W), np.uint8)
    depth = np.empty((n, H, W), np.uint16)
    for k in range(n):
        _, gray[k], depth[k] = st.frame(k)
    return gray, depth


def cpu_reference_fps(gray, depth, budget_s, warm=3):
    """frames/s of the cv2-based oracle (the reference's three OpenCV calls + its glue) on host cores."""
    import cv2
    from oracle.fe_oracle import IDC_CAM, FeatureTrackerOracle, PinholeCamera
    ft = FeatureTrackerOracle(PinholeCamera(**IDC_CAM), MAX_CNT, MIN_DIST, 1, 1)
    n = len(gray)
    k = 0
    for _ in range(warm):
        ft.trackImage(k / 30.0, gray[tri(k, n)], depth[tri(k, n)]); k += 1
    t0 = time.perf_counter(); done = 0
    while time.perf_counter() - t0 < budget_s:
        ft.trackImage(k / 30.0, gray[tri(k, n)], depth[tri(k, n)]); k += 1; done += 1
    dt = time.perf_counter() - t0
    return done / dt, done, cv2.getNumThreads()


def ba_roofline(nc, n_lm, iterations, device_ms, step_share=0.77):
    """Algorithmic FP64 work of k_ba_step (the dominant kernel: 77 % of a solve's kernel time in
    profiles/r1_ncu_launches_ba_v5.csv) against the FP64 rate of the ONE SM a single-CTA kernel can use (58 DFMA/clk measured
    on a B200 SM).  Per launch: Cholesky of the (nc+1)-row reduced system (nc+1)^3/3 FMA, the two triangular solves nc^2 FMA,
    landmark back-substitution and dogleg / model terms ~3*n_lm*nc FMA."""
    launches = iterations + 1
    fma = iterations * ((nc + 1) ** 3 / 3.0 + nc * nc + 3.0 * n_lm * nc)
    flops = 2.0 * fma
    sm_peak = 58.0 * 2 * 1.965e9
    t = step_share * device_ms * 1e-3
    ach = flops / t if t > 0 else None
    return {"kernel": "k_ba_step (single CTA: adoption, register-blocked Cholesky with look-ahead, back substitution, dogleg)",
            "bound": "latency: 43 dependent panels per factorisation, FP64 pipe 8 % busy (profiles/r1_ncu_k_ba_step_full.txt)",
            "algorithmic_flops_per_launch": flops / launches, "launches_per_solve": launches,
            "time_share_of_solve": step_share, "achieved": ach / 1e9 if ach else None, "unit": "GFLOP/s (FP64)",
            "peak": sm_peak / 1e9, "peak_source": "58 DFMA/clk measured on one B200 SM x 1.965 GHz x 2 flop",
            "frac": (ach / sm_peak) if ach else None}


def ba_bench(device, n_windows=8, reps=40, cpu_seconds=8.0, with_cpu=True):
    """Sliding-window solves/s (C2: 11 frames, ~1.5 k visual factors, 10 IMU factors, 8 iterations max) on the GPU
    through gf_ba_solve (host descriptor in, optimised blocks out: this IS the end-to-end call) next to the CPU oracle."""
    from ground_fusion_b200.estimator import BundleAdjuster
    from ground_fusion_b200.synth_ba import make_window
    wins = [make_window(seed=100 + k)[0] for k in range(n_windows)]
    saved = [(w.para_pose.copy(), w.para_speed_bias.copy(), w.para_feature.copy(), w.para_ex_pose.copy(), w.para_td.copy()) for w in wins]
    structs = [w.struct() for w in wins]

    def restore(k):
        w, sv = wins[k], saved[k]
        w.para_pose[:] = sv[0]; w.para_speed_bias[:] = sv[1]; w.para_feature[:] = sv[2]; w.para_ex_pose[:] = sv[3]; w.para_td[:] = sv[4]
    ba = BundleAdjuster(device)
    for k in range(n_windows):
        restore(k); ba.solve_struct(structs[k])
    dev_ms = 0.0; iters = 0
    t0 = time.perf_counter()
    for r in range(reps):
        k = r % n_windows
        restore(k)
        sm = ba.solve_struct(structs[k])
        dev_ms += sm.device_ms; iters += sm.iterations
    el = time.perf_counter() - t0
    out = {"metric": "ba_solves_per_sec", "value": reps / el, "unit": "solves/s", "ms_per_solve": 1e3 * el / reps,
           "device_ms_per_solve": dev_ms / reps, "iterations_per_solve": iters / reps,
           "workload": "C2 window: 11 frames, %d visual factors, %d IMU factors, reduced system %d + %d free landmarks, max 8 iterations"
                       % (wins[0].n_visual, wins[0].n_imu, sm.reduced_dim, sm.n_free_landmarks),
           "e2e": "value already includes the host->device upload of the problem and the download of the blocks"}
    out["roofline"] = ba_roofline(int(sm.reduced_dim), int(sm.n_free_landmarks), iters / reps, dev_ms / reps)
    # the marginalisation that ends Estimator::optimization() on a keyframe (MARGIN_OLD), on the solved window
    restore(0); ba.solve_struct(structs[0])
    ba.marginalize_old(wins[0])
    mms = []
    for _ in range(5):
        ba.marginalize_old(wins[0]); mms.append(ba.last_marg_ms)
    out["marginalize_old"] = {"device_ms": float(np.median(mms)), "note": "gf_ba_marginalize_old on the solved C2 window (191 marginalised + 76 kept dimensions)"}
    ba.close()
    # independent windows (several estimators sharing one GPU): one gf_ba handle and one host thread per stream; a solve
    # occupies 1-121 CTAs for microseconds at a time, so concurrent solves spread over the SMs
    import threading
    n_str = 4
    sets = []
    for t in range(n_str):
        ws = [make_window(seed=200 + 10 * t + k)[0] for k in range(2)]
        sets.append((ws, [w.struct() for w in ws], [(w.para_pose.copy(), w.para_speed_bias.copy(), w.para_feature.copy()) for w in ws], BundleAdjuster(device)))
    def worker(t, n):
        ws, st, sv, b = sets[t]
        for r in range(n):
            k = r % len(ws)
            ws[k].para_pose[:] = sv[k][0]; ws[k].para_speed_bias[:] = sv[k][1]; ws[k].para_feature[:] = sv[k][2]
            b.solve_struct(st[k])
    for t in range(n_str):
        worker(t, 2)
    th = [threading.Thread(target=worker, args=(t, reps)) for t in range(n_str)]
    t0 = time.perf_counter()
    for x in th:
        x.start()
    for x in th:
        x.join()
    elc = time.perf_counter() - t0
    out["concurrent_streams"] = {"streams": n_str, "value": n_str * reps / elc, "unit": "solves/s",
                                 "note": "aggregate of %d independent windows solved concurrently on one GPU (one gf_ba handle + host thread each)" % n_str}
    for s_ in sets:
        s_[3].close()
    if with_cpu:
        from oracle import ba_oracle
        t0 = time.perf_counter(); done = 0
        while time.perf_counter() - t0 < cpu_seconds:
            k = done % n_windows
            restore(k); ba_oracle.solve(wins[k]); done += 1
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": done / dt, "unit": "solves/s", "cores": 1, "kind": "port",
                               "sample": "%d solves of the same windows, oracle/ba_oracle.c (block-sparse, -O2, 1 thread)" % done}
        restore(0); ba_oracle.solve(wins[0])
        t0 = time.perf_counter(); ba_oracle.marginalize_old(wins[0]); 
        out["marginalize_old"]["cpu_oracle_ms"] = 1e3 * (time.perf_counter() - t0)
        out["marginalize_old"]["cpu_note"] = "oracle uses a plain cyclic Jacobi eigensolver (slower than Eigen's tridiagonal QR the reference calls)"
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ring", type=int, default=RING)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = {"workload": "C2: synthetic 640x480 RGB-D stream, 150 features, min_dist 30, flow_back 1 (BASELINE.json configs[1])",
           "frames_ring": args.ring, "l2_policy": "inputs larger than L2 (ring of %d distinct frames = %.0f MB per stream)" % (args.ring, args.ring * W * H * 3 / 1e6),
           "streams_per_gpu": 1, "parallelism": "one independent stream per GPU"}

    if args.impl == "reference":
        if rank != 0:
            return
        gray, depth = make_frames(0, min(args.ring, 60))
        vals = []
        for _ in range(max(1, min(args.steps, 3))):          # each step = a bounded sample of the workload
            fps, done, cores = cpu_reference_fps(gray, depth, min(args.cpu_seconds, 10.0))
            vals.append(fps)
        v = float(np.median(vals))
        print(json.dumps({"impl": "reference", "metric": "tracker_frames_per_sec", "value": v, "unit": "frames/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / v, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "u8/f32", "data": "synthetic", "config": cfg,
                          "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": "port",
                                           "sample": "%d frames of the C2 stream per sample, cv2 %s threads" % (done, cores)},
                          "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from ground_fusion_b200 import _lib
    from ground_fusion_b200.feature_tracker import FeatureTracker
    from ground_fusion_b200.synth import idc_params8      # config/realsense/idc_cam.yaml

    gray, depth = make_frames(rank, args.ring)
    d_gray = torch.from_numpy(gray).cuda()
    d_depth = torch.from_numpy(depth.view(np.int16)).cuda()
    h_gray = torch.from_numpy(gray).pin_memory(); h_depth = torch.from_numpy(depth.view(np.int16)).pin_memory()
    hg, hd = h_gray.numpy(), h_depth.numpy().view(np.uint16)
    tr = FeatureTracker(W, H, idc_params8(), MAX_CNT, MIN_DIST, 1, 1, device=local)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def run(n, mode, k0):
        """n frames through the asynchronous API, two in flight (submit k+1 before collecting k).  Returns the
        device time of the whole run (CUDA events around the first copy and the last result copy)."""
        tr.timer_start()
        pending = 0
        for k in range(k0, k0 + n):
            i = tri(k, args.ring)
            if mode == "device":
                tr.submitDevice(k / 30.0, d_gray[i].data_ptr(), d_depth[i].data_ptr())
            else:
                tr.submit(k / 30.0, hg[i], hd[i])      # pinned ring: unchanged until collected
            pending += 1
            if pending == 2:
                tr.wait(); pending -= 1
        while pending:
            tr.wait(); pending -= 1
        return tr.timer_stop()

    sampler = ClockSampler(local); sampler.start()
    launches0 = _lib.lib().gf_kernel_launch_count()
    # ---- kernel-side value: frames already in HBM ----
    run(args.warmup, "device", 0)
    barrier(); t0 = time.perf_counter()
    dev_ms = run(args.steps, "device", args.warmup)
    barrier(); el_dev = time.perf_counter() - t0
    launches = _lib.lib().gf_kernel_launch_count() - launches0
    # ---- stage breakdown (separate pass; event records only) ----
    tr.set_profiling(True)
    stage = {}; iters = 0; nprev = 0
    for k in range(args.warmup + args.steps, args.warmup + args.steps + 50):
        i = tri(k, args.ring)
        tr.trackDevice(k / 30.0, d_gray[i].data_ptr(), d_depth[i].data_ptr())
        for s, v in tr.last_stage_ms().items():
            stage[s] = stage.get(s, 0.0) + v / 50.0
        iters += tr.last_info["lk_iterations"] / 50.0; nprev += tr.last_info["n_prev"] / 50.0
    tr.set_profiling(False)
    # ---- end to end: host buffers in, observations out ----
    k1 = args.warmup + args.steps + 50
    run(args.warmup, "host", k1)
    barrier(); t0 = time.perf_counter()
    dev_ms_e2e = run(args.steps, "host", k1 + args.warmup)
    barrier(); el_e2e = time.perf_counter() - t0
    sampler.stop_flag = True; sampler.join(timeout=2)

    if dist is not None:
        tt = torch.tensor([el_dev, el_e2e, dev_ms, dev_ms_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        el_dev, el_e2e, dev_ms, dev_ms_e2e = float(tt[0]), float(tt[1]), float(tt[2]), float(tt[3])
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    # algorithmic bytes of k_track per launch: prev-window gathers (23x23 per level: 4 fwd + 2 bwd
    # levels per feature) + next-image window per LK iteration (22x22)
    lk_bytes = nprev * 6 * 23 * 23 + iters * 22 * 22
    lk_s = stage.get("lk", 0.0) / 1e3
    roof = {"kernel": "k_track (fwd 4-level + reverse 2-level LK, one 8-warp CTA per feature)", "bound": "hbm",
            "achieved": (lk_bytes / lk_s / 1e9) if lk_s > 0 else None, "peak": hbm,
            "unit": "GB/s", "frac": (lk_bytes / lk_s / 1e9 / hbm) if lk_s > 0 else None,
            "traffic": 876544, "traffic_source": "dram__bytes_read+write of one k_track launch, profiles/r1_ncu_k_track_full.txt",
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
            "algorithmic_bytes_per_launch": lk_bytes, "kernel_ms": stage.get("lk"),
            "note": "latency-bound by construction: per feature a chain of ~22 dependent LK iterations, each 105 dependent FADDs in OpenCV lane order; kernel_ms is the stage time of a profiled (one frame at a time) pass"}
    value = world * args.steps / el_dev
    out = {"metric": "tracker_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1000.0 * el_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "u8/i32 fixed point + f32 (LK, min-eig), f64 (box sums, undistortion)", "data": "synthetic", "config": cfg,
           "device_ms_per_step": dev_ms / args.steps, "device_ms_per_step_e2e": dev_ms_e2e / args.steps,
           "timing": "value/e2e: wall clock between barrier+synchronize, max over ranks; device_ms_per_step*: CUDA events "
                     "around the same K frames on the tracker's streams (first copy .. last result copy), max over ranks",
           "frames_in_flight": 2, "stage_ms": stage,
           "e2e": {"value": world * args.steps / el_e2e, "unit": "frames/s", "h2d_bytes_per_step": W * H * 3,
                   "d2h_bytes_per_step": MAX_CNT * 72 + 40 + MAX_CNT},
           "gpu_launches": int(launches), "roofline": roof, "clocks": sampler.summary()}
    if world == 1:
        fps, done, cores = cpu_reference_fps(gray[:60], depth[:60], args.cpu_seconds)
        out["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                               "sample": "%d frames of the same C2 stream, cv2-based oracle (reference's OpenCV calls + glue)" % done}
        try:
            out["ba"] = ba_bench(local, cpu_seconds=min(args.cpu_seconds, 8.0))
        except Exception as e:      # the FE line must survive a BA problem
            out["ba"] = {"error": repr(e)}
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
