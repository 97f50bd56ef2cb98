// dropin/BFMatcher.h - a GPU-backed stand-in for the ONE cv::BFMatcher the reference holds:
//   R/include/Frame.h:288   static cv::BFMatcher BFmatcher;
//   R/src/Frame.cc:26       cv::BFMatcher Frame::BFmatcher = cv::BFMatcher(cv::NORM_HAMMING);
//   R/src/Frame.cc:1130     BFmatcher.knnMatch(stereoDescLeft, stereoDescRight, matches, 2);    (Frame::ComputeStereoFishEyeMatches)
// Optional optimisation: change the type in those two declarations to ORB_SLAM3::BFMatcherB200 and the brute-force
// search of the two-camera (KannalaBrandt8) frames runs on the tensor cores (orbx_bf_knn2); the ratio test and the
// triangulation of ComputeStereoFishEyeMatches stay the reference's own code.  Same results as cv::BFMatcher(NORM_HAMMING).knnMatch(k = 2):
// per query row up to two train rows ordered by (distance, train index), fewer when the train set has fewer rows.
#pragma once
#include <vector>
#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>

namespace ORB_SLAM3
{

class BFMatcherB200
{
public:
    explicit BFMatcherB200(int normType = cv::NORM_HAMMING, bool crossCheck = false);
    // only what the reference calls: NORM_HAMMING, k = 2, 32-byte descriptor rows; anything else throws std::runtime_error
    void knnMatch(cv::InputArray queryDescriptors, cv::InputArray trainDescriptors, std::vector<std::vector<cv::DMatch> >& matches, int k) const;

private:
    int norm_;
};

}  // namespace ORB_SLAM3
