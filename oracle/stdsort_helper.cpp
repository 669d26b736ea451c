// oracle/stdsort_helper.cpp -- TEST INFRASTRUCTURE ONLY.
// FeatureTracker::setMask (reference feature_tracker.cpp:56-83) orders features with
//   std::sort(cnt_pts_id.begin(), cnt_pts_id.end(), [](a, b){ return a.first > b.first; });
// std::sort is unstable, so the order of equal track counts is whatever libstdc++'s introsort does.
// This helper runs the real std::sort (this image: gcc 13 libstdc++) on (track_cnt, index) pairs with
// the same comparator and returns the permutation, so the oracle reproduces the reference order.
#include <algorithm>
#include <utility>
#include <vector>

extern "C" __attribute__((visibility("default")))
void gfo_setmask_order(const int* track_cnt, int n, int* perm)
{
    std::vector<std::pair<int, std::pair<std::pair<float, float>, int>>> v;
    v.reserve(n);
    for (int i = 0; i < n; i++) v.push_back(std::make_pair(track_cnt[i], std::make_pair(std::make_pair(0.f, 0.f), i)));
    std::sort(v.begin(), v.end(),
              [](const std::pair<int, std::pair<std::pair<float, float>, int>>& a,
                 const std::pair<int, std::pair<std::pair<float, float>, int>>& b) { return a.first > b.first; });
    for (int i = 0; i < n; i++) perm[i] = v[i].second.second;
}

// FeatureTracker::setMask as a whole (feature_tracker.cpp:56-83) for the vectorised oracle (bench.py's CPU arm): the same
// std::sort, then the walk "keep a point iff mask(pt) == 255, then cv::circle(mask, pt, MIN_DIST, 0, -1)".  cv::circle's
// filled circle is the integer disk d^2 <= r^2 clipped to the image (tests/test_fe_oracle.py::test_circle_is_integer_disk).
// px, py: cvRound-ed point coordinates.  mask: rows x cols, filled with 255 by the caller.  Returns the number kept.
extern "C" __attribute__((visibility("default")))
int gfo_setmask(const int* track_cnt, const int* px, const int* py, int n, int min_dist, int rows, int cols,
                unsigned char* mask, int* keep_idx)
{
    std::vector<int> perm(n > 0 ? n : 1);
    gfo_setmask_order(track_cnt, n, perm.data());
    int nk = 0;
    const int r = min_dist, r2 = r * r;
    for (int k = 0; k < n; k++) {
        const int j = perm[k], cx = px[j], cy = py[j];
        if (mask[(size_t)cy * cols + cx] != 255) continue;
        keep_idx[nk++] = j;
        const int y0 = cy - r < 0 ? 0 : cy - r, y1 = cy + r >= rows ? rows - 1 : cy + r;
        for (int y = y0; y <= y1; y++) {
            const int dy = y - cy;
            int hw = 0;
            while ((hw + 1) * (hw + 1) + dy * dy <= r2) hw++;
            const int x0 = cx - hw < 0 ? 0 : cx - hw, x1 = cx + hw >= cols ? cols - 1 : cx + hw;
            for (int x = x0; x <= x1; x++) mask[(size_t)y * cols + x] = 0;
        }
    }
    return nk;
}
