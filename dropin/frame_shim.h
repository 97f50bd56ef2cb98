// frame_shim.h - the members of the reference's Frame / MapPoint / GeometricCamera that the drop-in ORBmatcher
// touches (R/orb_slam3/include/Frame.h, MapPoint.h, CameraModels/GeometricCamera.h), for compile-checking in an image
// without OpenCV/Eigen.  In the reference tree the real headers are used (-DORBX_USE_REAL_OPENCV).
#pragma once
#ifdef ORBX_USE_REAL_OPENCV
#include "MapPoint.h"
#include "KeyFrame.h"
#include "Frame.h"
#else
#include <map>
#include <set>
#include <vector>
#include "cv_shim.h"
#ifndef ORBX_DBOW2_SHIM
#define ORBX_DBOW2_SHIM
namespace DBoW2 {      // same value types as R/Thirdparty/DBoW2/DBoW2/BowVector.h:23-29, FeatureVector.h
typedef unsigned int WordId;
typedef double WordValue;
typedef unsigned int NodeId;
class BowVector : public std::map<WordId, WordValue> {};
class FeatureVector : public std::map<NodeId, std::vector<unsigned int> > {};
}  // namespace DBoW2
#endif
namespace ORB_SLAM3 {
class MapPoint {
public:
    // tracking state written by Frame::isInFrustum (MapPoint.h)
    float mTrackProjX = 0, mTrackProjY = 0, mTrackDepth = 0, mTrackProjXR = 0, mTrackViewCos = 1;
    int mnTrackScaleLevel = 0;
    bool mbTrackInView = false, mbTrackInViewR = false;
    float mWorldPos[3] = {0, 0, 0};
    cv::Mat mDescriptor;
    int nObs = 1; bool bad = false;
    cv::Mat GetDescriptor() const { return mDescriptor; }
    int Observations() const { return nObs; }
    bool isBad() const { return bad; }
    const float* GetWorldPosPtr() const { return mWorldPos; }
};
struct GeometricCamera { float fx = 1, fy = 1, cx = 0, cy = 0;
    cv::Point2f project(const float p[3]) const { return cv::Point2f(fx * p[0] / p[2] + cx, fy * p[1] / p[2] + cy); } };
class KeyFrame {          // the members SearchByBoW reads (R/include/KeyFrame.h)
public:
    int N = 0, NLeft = -1;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn, mvKeysRight;
    cv::Mat mDescriptors;
    DBoW2::BowVector mBowVec;
    DBoW2::FeatureVector mFeatVec;
    std::vector<MapPoint*> mvpMapPoints;
    GeometricCamera* mpCamera = nullptr; GeometricCamera* mpCamera2 = nullptr;
    std::vector<MapPoint*> GetMapPointMatches() const { return mvpMapPoints; }
};
class Frame {
public:
    int N = 0, Nleft = -1;
    DBoW2::BowVector mBowVec;
    DBoW2::FeatureVector mFeatVec;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn, mvKeysRight;
    cv::Mat mDescriptors;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    std::vector<float> mvuRight, mvScaleFactors;
    float mb = 0, mbf = 0;
    float mRcw[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, mtcw[3] = {0, 0, 0};   // rows of mTcw
    GeometricCamera* mpCamera = nullptr;
    static float mnMinX, mnMaxX, mnMinY, mnMaxY;
};
}  // namespace ORB_SLAM3
#endif
