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
