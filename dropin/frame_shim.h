// frame_shim.h - the members of the reference's Frame / MapPoint / GeometricCamera that the drop-in ORBmatcher
// touches (R/orb_slam3/include/Frame.h, MapPoint.h, CameraModels/GeometricCamera.h), for compile-checking in an image
// without OpenCV/Eigen.  In the reference tree the real headers are used (-DORBX_USE_REAL_OPENCV).
#pragma once
#ifdef ORBX_USE_REAL_OPENCV
#include "MapPoint.h"
#include "KeyFrame.h"
#include "Frame.h"
#else
#include <set>
#include <vector>
#include "cv_shim.h"
namespace ORB_SLAM3 {
class KeyFrame;
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
class Frame {
public:
    int N = 0, Nleft = -1;
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
