// FeatureTracker with the reference's class surface (vins_estimator/src/featureTracker/feature_tracker.h:43-99) whose
// trackImage / setPrediction / removeOutliers bodies call libgf_b200.so.  Callers (Estimator::inputImage,
// estimator.cpp:182-221; sync_process in rosNodeTest.cpp) compile against it unchanged.  The members the reference's other
// code reads after a frame (drawTrack: cur_pts, ids, track_cnt; prev/cur time) are kept and refreshed from the result.
#pragma once
#include <map>
#include <set>
#include <string>
#include <utility>
#include <vector>

#include <eigen3/Eigen/Dense>
#include <opencv2/opencv.hpp>

#include "gf_b200.h"

// parameters.h:130-134 (read from the YAML by readParameters): defined by the node, declared here as the reference does
extern int MAX_CNT;
extern int MIN_DIST;
extern int FLOW_BACK;

class FeatureTracker
{
  public:
    FeatureTracker();
    ~FeatureTracker();
    FeatureTracker(const FeatureTracker &) = delete;
    FeatureTracker &operator=(const FeatureTracker &) = delete;

    std::map<int, std::vector<std::pair<int, Eigen::Matrix<double, 8, 1>>>> trackImage(double _cur_time, const cv::Mat &_img, const cv::Mat &_img1 = cv::Mat());
    // readIntrinsicParameter(calib_file, depth) (feature_tracker.cpp:745-762) parses the camodocal YAML; the parsed PINHOLE
    // parameters are all this path needs.  With camodocal present: m_camera[0]->writeParameters(p) -> setPinhole(p[4..7], p[0..3]).
    void setPinhole(double fx, double fy, double cx, double cy, double k1, double k2, double p1, double p2, int depth);
    void setPrediction(std::map<int, Eigen::Vector3d> &predictPts);
    void removeOutliers(std::set<int> &removePtsIds);

    int row = 0, col = 0;
    std::vector<cv::Point2f> cur_pts;
    std::vector<int> ids;
    std::vector<int> track_cnt;
    double cur_time = 0, prev_time = 0;
    bool stereo_cam = false;
    bool depth_cam = false;
    bool hasPrediction = false;
    int device = 0;          // CUDA device of this camera stream

  private:
    gf_tracker *gf_ = nullptr;          // created on the first frame: needs its size
    gf_tracker_cfg cfg_{};
    std::vector<gf_obs> out_;
};
