/*
 * gf_b200.h -- C ABI of libgf_b200.so, the B200 (sm_100a) implementation of Ground-Fusion's two
 * data-parallel hot paths.  Plain pointers and sizes only; no exceptions cross this boundary.
 * Every entry point returns 0 on success or a negative gf_status.  Handles are thread-compatible:
 * one host thread per handle at a time (the reference runs FE on sync_thread and BA on processThread).
 *
 * Reference interfaces replaced (paths relative to /root/reference):
 *   gf_tracker_*   <- class FeatureTracker, vins_estimator/src/featureTracker/feature_tracker.h:43-99
 *   gf_ba_*        <- Estimator::optimization(), vins_estimator/src/estimator/estimator.h:147
 *                     (estimator.cpp:2890-3636) and MarginalizationInfo (factor/marginalization_factor.h)
 */
#ifndef GF_B200_H
#define GF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden */
#endif

typedef enum gf_status {
    GF_OK = 0,
    GF_ERR_INVALID_ARG = -1,
    GF_ERR_CUDA = -2,         /* a CUDA call failed; gf_last_error() has the text */
    GF_ERR_NO_DEVICE = -3,    /* no CUDA device: there is NO CPU fallback */
    GF_ERR_CAPACITY = -4,
    GF_ERR_UNSUPPORTED = -5
} gf_status;

/* Text of the last error raised on the calling thread (never NULL). */
const char* gf_last_error(void);
/* Library version string, e.g. "gf_b200 0.1 sm_100a". */
const char* gf_version(void);
/* Number of kernel launches issued by this library since load (all handles). */
uint64_t gf_kernel_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Front end: FeatureTracker (feature_tracker.h:43-99)
 * ---------------------------------------------------------------------------------------------- */
typedef struct gf_tracker gf_tracker; /* opaque: one per camera stream, bound to one GPU */

typedef struct gf_tracker_cfg {
    int max_cnt;        /* MAX_CNT   (parameters.h:130; groundchallenge.yaml:101)            */
    int min_dist;       /* MIN_DIST  (parameters.h:131; groundchallenge.yaml:102)            */
    int flow_back;      /* FLOW_BACK (parameters.h:134; groundchallenge.yaml:106)            */
    int depth_cam;      /* readIntrinsicParameter(calib, depth): RGB-D mode (feature_tracker.cpp:757-758) */
    double pinhole[8];  /* fx fy cx cy k1 k2 p1 p2 (camodocal PINHOLE; config/realsense/idc_cam.yaml) */
} gf_tracker_cfg;

/* One entry of the map trackImage returns (feature_tracker.cpp:344-366):
 * v = [x_n, y_n, 1, u, v, vx_n, vy_n, depth_m]; camera id is always 0 on this path. */
typedef struct gf_obs {
    int32_t id;
    int32_t track_cnt;
    double v[8];
} gf_obs;

/* Per-call diagnostics the parity tests compare with the oracle ("inlier masks"). */
typedef struct gf_track_info {
    int32_t n_prev;     /* points that entered LK                                           */
    int32_t n_tracked;  /* points that survived status/reverse/border/saturation             */
    int32_t n_kept;     /* ... and survived setMask                                          */
    int32_t n_new;      /* corners added by goodFeaturesToTrack                              */
    int32_t n_candidates; /* GFTT local-maximum candidates before the min-distance pass      */
    int32_t nms_rounds; /* parallel min-distance rounds used                                 */
    int32_t eig_fixups; /* column bands re-run by the box-filter verifier                    */
    int32_t reserved;
} gf_track_info;

/* FeatureTracker::FeatureTracker + readIntrinsicParameter (feature_tracker.cpp:48-54, 745-759). */
int gf_tracker_create(gf_tracker** out, int device, int width, int height, const gf_tracker_cfg* cfg);
void gf_tracker_destroy(gf_tracker* t);

/* Pinned host staging buffers owned by the tracker (width*height u8, width*height u16).  A caller
 * that renders/decodes straight into them avoids one host memcpy; any other pointer is also accepted
 * by gf_tracker_track and is copied into these buffers first. */
int gf_tracker_host_buffers(gf_tracker* t, uint8_t** gray, uint16_t** depth);

/* FeatureTracker::trackImage(t, img, depth) (feature_tracker.cpp:103-372).
 *   gray: H x W u8, row pitch gray_pitch bytes; only read during the call (as cv::Mat in the reference)
 *   depth: H x W u16 millimetres or NULL (then v[7] = -2.4 as feature_tracker.cpp:338)
 *   out: caller-allocated, capacity >= cfg.max_cnt; entries are in the tracker's internal order
 *        (the reference's std::map iterates by id: sort by id on the caller side if needed)
 *   status_out (nullable, capacity >= max_cnt): combined status of the n_prev points that entered LK
 *        (LK status & reverse check & inBorder & grey<=250), i.e. the reference's `status` vector at
 *        feature_tracker.cpp:170 */
int gf_tracker_track(gf_tracker* t, double time, const uint8_t* gray, size_t gray_pitch,
                     const uint16_t* depth, size_t depth_pitch, gf_obs* out, int* n_out,
                     uint8_t* status_out, gf_track_info* info);

/* Asynchronous split of gf_tracker_track for callers that overlap several streams on one GPU:
 * _submit enqueues copies + kernels and returns, _wait blocks for the result of the last _submit. */
int gf_tracker_submit(gf_tracker* t, double time, const uint8_t* gray, size_t gray_pitch,
                      const uint16_t* depth, size_t depth_pitch);
int gf_tracker_wait(gf_tracker* t, gf_obs* out, int* n_out, uint8_t* status_out, gf_track_info* info);

/* Device-resident variant used by bench.py's kernel-only `value`: the frame is already in HBM
 * (device pointers, tightly packed W x H), nothing is copied from the host. */
int gf_tracker_track_device(gf_tracker* t, double time, const void* d_gray, const void* d_depth,
                            gf_obs* out, int* n_out, uint8_t* status_out, gf_track_info* info);

/* FeatureTracker::setPrediction (feature_tracker.cpp:1006-1027): xyz are camera-frame 3-D points. */
int gf_tracker_set_prediction(gf_tracker* t, const int32_t* ids, const double* xyz, int n);
/* FeatureTracker::removeOutliers (feature_tracker.cpp:1029-1045). */
int gf_tracker_remove_ids(gf_tracker* t, const int32_t* ids, int n);

/* Device time (ms, CUDA events on the tracker's stream) of the last completed frame. */
int gf_tracker_last_device_ms(gf_tracker* t, float* ms);

/* ------------------------------------------------------------------------------------------------
 * Stage-level entry points (same kernels as the tracker; used by the parity tests, which read like
 * tests of the three OpenCV calls the reference makes).  All pointers are HOST pointers.
 * ---------------------------------------------------------------------------------------------- */
/* cv::pyrDown on u8 (LK pyramid level). dst is ((w+1)/2) x ((h+1)/2), tightly packed. */
int gf_stage_pyr_down(int device, const uint8_t* src, int w, int h, uint8_t* dst);
/* cv::cornerMinEigenVal(img, blockSize=3, ksize=3) as used by goodFeaturesToTrack. */
int gf_stage_min_eig(int device, const uint8_t* img, int w, int h, float* eig, int* n_fixups);
/* cv::calcOpticalFlowPyrLK(prev, next, prev_pts, next_pts, status, err, Size(21,21), max_level,
 * TermCriteria(COUNT+EPS,30,0.01), use_initial_flow ? OPTFLOW_USE_INITIAL_FLOW : 0). */
int gf_stage_lk(int device, const uint8_t* prev, const uint8_t* next, int w, int h,
                const float* prev_pts, float* next_pts, int n, int max_level, int use_initial_flow,
                uint8_t* status);
/* cv::goodFeaturesToTrack(img, max_corners, 0.01, min_dist, mask) where mask = 255 everywhere except
 * the integer disks d^2 <= min_dist^2 around cvRound(kept_pts[i]) (setMask, feature_tracker.cpp:56-83).
 * Returns the corners in OpenCV's output order. */
int gf_stage_gftt(int device, const uint8_t* img, int w, int h, const float* kept_pts, int n_kept,
                  int max_corners, int min_dist, float* corners, int* n_corners, gf_track_info* info);
/* The permutation libstdc++'s std::sort produces for setMask's comparator (device replica). */
int gf_stage_setmask_order(int device, const int32_t* track_cnt, int n, int32_t* perm);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* GF_B200_H */
