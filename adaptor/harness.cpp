// Drives the adaptor classes from flat binary dumps written by tests/test_adaptor.py; the outputs go back as flat binaries.
//   harness fe <in> <out>   frames through FeatureTracker::trackImage
//   harness ba <in> <out>   one window through Estimator::optimization()
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "estimator_ba.h"
#include "feature_tracker.h"

int MAX_CNT = 150, MIN_DIST = 30, FLOW_BACK = 1;
int USE_IMU = 1, NUM_ITERATIONS = 8, ESTIMATE_TD = 0;
double FOCAL_LENGTH = 600.0;
Eigen::Vector3d G;

static void rd(FILE *f, void *p, size_t n) { if (fread(p, 1, n, f) != n) { std::fprintf(stderr, "short read\n"); std::exit(2); } }
template <class T> static T rd1(FILE *f) { T v; rd(f, &v, sizeof(T)); return v; }

static int run_fe(const char *in, const char *out)
{
    FILE *f = fopen(in, "rb"), *o = fopen(out, "wb");
    if (!f || !o) return 2;
    const int n = rd1<int>(f), w = rd1<int>(f), h = rd1<int>(f), with_depth = rd1<int>(f);
    double p[8]; rd(f, p, sizeof(p));
    MAX_CNT = rd1<int>(f); MIN_DIST = rd1<int>(f); FLOW_BACK = rd1<int>(f);
    FeatureTracker tracker;
    tracker.setPinhole(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], with_depth);
    std::vector<unsigned char> gray((size_t)w * h); std::vector<unsigned short> depth((size_t)w * h);
    for (int k = 0; k < n; k++) {
        const double t = rd1<double>(f);
        rd(f, gray.data(), gray.size());
        if (with_depth) rd(f, depth.data(), depth.size() * 2);
        cv::Mat img(h, w, CV_8UC1, gray.data()), dimg;
        if (with_depth) dimg = cv::Mat(h, w, CV_16UC1, depth.data());
        auto frame = tracker.trackImage(t, img, dimg);
        const int m = (int)frame.size();
        fwrite(&m, 4, 1, o);
        for (auto &kv : frame) {                                   // std::map: ascending feature id, as featureBuf consumers see it
            fwrite(&kv.first, 4, 1, o);
            fwrite(kv.second[0].second.data(), 8, 8, o);
        }
    }
    fclose(f); fclose(o);
    return 0;
}

static int run_ba(const char *in, const char *out)
{
    FILE *f = fopen(in, "rb"), *o = fopen(out, "wb");
    if (!f || !o) return 2;
    static Estimator est;
    const int n_frames = rd1<int>(f), n_windows = rd1<int>(f);
    est.frame_count = n_frames - 1;
    NUM_ITERATIONS = rd1<int>(f);
    rd(f, G.data(), 24);
    for (int w = 0; w < n_windows; w++) {
        est.marginalization_flag = rd1<int>(f) ? Estimator::MARGIN_SECOND_NEW : Estimator::MARGIN_OLD;
        const int n_feat = rd1<int>(f);
        rd(f, est.para_Pose, sizeof(double) * 7 * n_frames);
        rd(f, est.para_SpeedBias, sizeof(double) * 9 * n_frames);
        rd(f, est.para_Ex_Pose[0], 56);
        for (int k = 0; k < n_frames; k++) for (int c = 0; c < 3; c++) est.Vs[k](c) = est.para_SpeedBias[k][c];
        for (int j = 1; j < n_frames; j++) {
            delete est.pre_integrations[j];
            IntegrationBase *b = est.pre_integrations[j] = new IntegrationBase;
            b->sum_dt = rd1<double>(f);
            rd(f, b->delta_p.data(), 24); rd(f, b->delta_q.coeffs().data(), 32); rd(f, b->delta_v.data(), 24);
            rd(f, b->linearized_ba.data(), 24); rd(f, b->linearized_bg.data(), 24);
            double m[225];
            rd(f, m, sizeof(m)); for (int i = 0; i < 15; i++) for (int c = 0; c < 15; c++) b->jacobian(i, c) = m[i * 15 + c];
            rd(f, m, sizeof(m)); for (int i = 0; i < 15; i++) for (int c = 0; c < 15; c++) b->covariance(i, c) = m[i * 15 + c];
        }
        est.f_manager.feature.clear();
        for (int k = 0; k < n_feat; k++) {
            FeaturePerId it;
            it.feature_id = k; it.start_frame = rd1<int>(f);
            const int n_obs = rd1<int>(f);
            it.estimate_flag = rd1<int>(f);
            est.para_Feature[k][0] = rd1<double>(f);
            for (int q = 0; q < n_obs; q++) {
                FeaturePerFrame fr;
                rd(f, fr.point.data(), 24); rd(f, fr.velocity.data(), 16); fr.cur_td = rd1<double>(f);
                it.feature_per_frame.push_back(fr);
            }
            est.f_manager.feature.push_back(it);
        }
        est.optimization();
        fwrite(est.para_Pose, sizeof(double) * 7, n_frames, o);
        fwrite(est.para_SpeedBias, sizeof(double) * 9, n_frames, o);
        fwrite(&est.para_Feature[0][0], sizeof(double), n_feat, o);
        const double tail[4] = {(double)est.last_summary.iterations, est.last_summary.final_cost, (double)est.prior_dim(), (double)est.last_summary.termination};
        fwrite(tail, 8, 4, o);
    }
    fclose(f); fclose(o);
    return 0;
}

int main(int argc, char **argv)
{
    if (argc != 4) { std::fprintf(stderr, "usage: harness fe|ba <in> <out>\n"); return 2; }
    return argv[1][0] == 'f' ? run_fe(argv[2], argv[3]) : run_ba(argv[2], argv[3]);
}
