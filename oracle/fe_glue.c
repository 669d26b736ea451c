/* oracle/fe_glue.c -- TEST INFRASTRUCTURE ONLY.
 * The per-feature glue of FeatureTracker::trackImage (reference vins_estimator/src/featureTracker/feature_tracker.cpp)
 * in plain C, for the vectorised CPU arm bench.py times (oracle/fe_oracle.py::FeatureTrackerOracleFast): with it the arm
 * spends its time in the three OpenCV calls like the compiled reference, not in the Python interpreter.
 * tests/test_fe_oracle.py::test_fast_oracle_equals_loop_oracle keeps it equal to the loop restatement.
 * Built with -ffp-contract=off: the reference is compiled without FMA contraction (-O3, no -march). */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

#define API __attribute__((visibility("default")))

/* feature_tracker.cpp:137-168: reverse check (distance(), :95-101), inBorder (:14-20, cvRound), grey <= 250 with the
 * transposed read cur_img.at<uchar>((int)x, (int)y); out of bounds => "not saturated" (policy in fe_oracle.py) */
API void gfo_status_rules(const float* prev_pts, const float* cur_pts, const float* rev_pts, const uint8_t* status,
                          const uint8_t* rev_status, int n, int flow_back, const uint8_t* img, int rows, int cols, uint8_t* out)
{
    for (int i = 0; i < n; i++) {
        int ok = status[i] != 0;
        if (flow_back) {
            double dx = (double)(prev_pts[2 * i] - rev_pts[2 * i]), dy = (double)(prev_pts[2 * i + 1] - rev_pts[2 * i + 1]);
            ok = ok && rev_status[i] && sqrt(dx * dx + dy * dy) <= 0.5;
        }
        if (ok) {
            long ix = lrintf(cur_pts[2 * i]), iy = lrintf(cur_pts[2 * i + 1]);
            if (!(1 <= ix && ix < cols - 1 && 1 <= iy && iy < rows - 1)) ok = 0;
        }
        if (ok) {
            int pu = (int)cur_pts[2 * i], pv = (int)cur_pts[2 * i + 1];
            int grey = (pu >= 0 && pu < rows && pv >= 0 && pv < cols) ? img[(size_t)pu * cols + pv] : 0;
            if (grey > 250) ok = 0;
        }
        out[i] = (uint8_t)ok;
    }
}

/* camodocal PinholeCamera::distortion / liftProjective (PinholeCamera.cc:450-510, 646-664) */
static void distortion(const double* k, double x, double y, double* dx, double* dy)
{
    double mx2 = x * x, my2 = y * y, mxy = x * y, rho2 = mx2 + my2;
    double rad = k[4] * rho2 + k[5] * rho2 * rho2;
    *dx = x * rad + 2.0 * k[6] * mxy + k[7] * (rho2 + 2.0 * mx2);
    *dy = y * rad + 2.0 * k[7] * mxy + k[6] * (rho2 + 2.0 * my2);
}

/* undistortedPts (:797-808), ptsVelocity (:810-847), observation vectors (:318-366).
 * cam = fx fy cx cy k1 k2 p1 p2.  prev_ids/prev_un: the previous frame's cur_un_pts_map.  depth nullable.
 * obs: n x 8 doubles [x_n, y_n, 1, u, v, vx, vy, depth]. */
API void gfo_finalize(const float* cur_pts, const int64_t* ids, int n, const double* cam, const int64_t* prev_ids,
                      const float* prev_un, int n_prev, double dt, const uint16_t* depth, int rows, int cols, int depth_cam,
                      float* cur_un, double* obs)
{
    const double ik11 = 1.0 / cam[0], ik13 = -cam[2] / cam[0], ik22 = 1.0 / cam[1], ik23 = -cam[3] / cam[1];
    const int nodist = cam[4] == 0.0 && cam[5] == 0.0 && cam[6] == 0.0 && cam[7] == 0.0;
    (void)rows;
    for (int i = 0; i < n; i++) {
        const double u = (double)cur_pts[2 * i], v = (double)cur_pts[2 * i + 1];
        double mxd = ik11 * u + ik13, myd = ik22 * v + ik23, mxu = mxd, myu = myd;
        if (!nodist) {
            double dx, dy;
            distortion(cam, mxd, myd, &dx, &dy);
            mxu = mxd - dx; myu = myd - dy;
            for (int it = 1; it < 8; it++) { distortion(cam, mxu, myu, &dx, &dy); mxu = mxd - dx; myu = myd - dy; }
        }
        const float ux = (float)(mxu / 1.0), uy = (float)(myu / 1.0);
        cur_un[2 * i] = ux; cur_un[2 * i + 1] = uy;
        float vx = 0.f, vy = 0.f;
        if (n_prev > 0)
            for (int j = 0; j < n_prev; j++)
                if (prev_ids[j] == ids[i]) {
                    vx = (float)((double)(ux - prev_un[2 * j]) / dt);
                    vy = (float)((double)(uy - prev_un[2 * j + 1]) / dt);
                    break;
                }
        double dval = -2.4;
        if (depth_cam && depth) {
            const long r = lround((double)cur_pts[2 * i + 1]), c = lround((double)cur_pts[2 * i]);
            dval = (double)(int)depth[(size_t)r * cols + c] / 1000;
        }
        double* o = obs + 8 * (size_t)i;
        o[0] = ux; o[1] = uy; o[2] = 1.0; o[3] = cur_pts[2 * i]; o[4] = cur_pts[2 * i + 1]; o[5] = vx; o[6] = vy; o[7] = dval;
    }
}
