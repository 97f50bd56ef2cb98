// orbslam_world.h - stand-ins for the reference classes AROUND the ORB front-end: ORB_SLAM3::Frame, KeyFrame, MapPoint, Map,
// Converter (R/include/Frame.h, KeyFrame.h, MapPoint.h, Map.h, Converter.h).
//
// Why: the real headers pull ROS messages, Eigen, g2o, Boost and Sophus, none of which exist in this image.  The stand-ins
// declare the SAME member names and types for everything the ORB front-end touches (ORBmatcher.cc, Frame::ComputeStereoMatches,
// the grid, BoW, MapPoint bookkeeping), so that
//   * the reference's own R/src/ORBmatcher.cc compiles UNMODIFIED against them (oracle/_ref, the parity oracle of record), and
//   * dropin/ORBmatcher.cc, written against the real headers' names, compiles here against the very same declarations.
// Member functions with logic are only DECLARED: their bodies are the reference's own, extracted by function name from
// Frame.cc / KeyFrame.cc / MapPoint.cc / Converter.cc at build time (oracle/ref/extract_functions.py -> oracle/_ref/gen/), never
// restated.  What is written by hand here is construction (the reference builds these objects from images, ROS messages and
// files) and pose setters stripped of their communication side effects.
//
// This header is force-included (-include) with the include guards of the headers it replaces pre-defined, so the `#include
// "Frame.h"` etc. of the reference's ORBmatcher.h become no-ops.  In the reference tree none of this is used.
#pragma once
#define FRAME_H
#define KEYFRAME_H
#define MAPPOINT_H
#define CONVERTER_H
#define TwoViewReconstruction_H
#define DATATYPES_H_

#include <cmath>
#include <climits>
#include <iostream>
#include <list>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <tuple>
#include <vector>
#include <opencv2/core/core.hpp>
#include "Thirdparty/DBoW2/DBoW2/BowVector.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"
#include "ORBVocabulary.h"
#include "ORBextractor.h"

namespace ORB_SLAM3
{
using namespace std;        // R/include/Datatypes.h:31 does this for the whole namespace; the reference sources rely on it

enum eSystemState { NOTYPE = -1, CLIENT = 0, SERVER = 1 };     // R/include/Datatypes.h:35-39

class MapPoint;
class KeyFrame;
class Frame;
class Map;
class GeometricCamera;

// R/include/TwoViewReconstruction.h: only constructed inside Pinhole::ReconstructWithTwoViews (monocular map initialisation,
// not on the ORB path)
class TwoViewReconstruction {
public:
    TwoViewReconstruction(cv::Mat&, float = 1.0f, int = 200) {}
    bool Reconstruct(const std::vector<cv::KeyPoint>&, const std::vector<cv::KeyPoint>&, const std::vector<int>&, cv::Mat&, cv::Mat&,
                     std::vector<cv::Point3f>&, std::vector<bool>&) { return false; }
};
}  // namespace ORB_SLAM3

#include "CameraModels/GeometricCamera.h"
#include "CameraModels/Pinhole.h"

namespace ORB_SLAM3
{

class Converter {            // R/include/Converter.h
public:
    static std::vector<cv::Mat> toDescriptorVector(const cv::Mat &Descriptors);
};

class Map {                  // R/include/Map.h: the one member MapPoint::Replace calls
public:
    void EraseMapPoint(MapPoint*) {}
};

#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64

class Frame                  // R/include/Frame.h:47-363
{
public:
    Frame() : mpORBvocabulary(NULL), mpORBextractorLeft(NULL), mpORBextractorRight(NULL), mTimeStamp(0), mbf(0), mb(0), mThDepth(0), N(0),
              mnCloseMPs(0), mnId(nNextId++), mpReferenceKF(NULL), mnScaleLevels(0), mfScaleFactor(0), mfLogScaleFactor(0), mnClientId(0),
              mpCamera(NULL), mpCamera2(NULL), Nleft(-1), Nright(-1), monoLeft(-1), monoRight(-1) {}

    // ---- bodies from R/src/Frame.cc ----
    void ExtractORB(int flag, const cv::Mat &im, const int x0, const int x1);
    void ComputeBoW();
    void SetPose(cv::Mat Tcw);
    void UpdatePoseMatrices();
    inline cv::Mat GetCameraCenter() { return mOw.clone(); }
    inline cv::Mat GetRotationInverse() { return mRwc.clone(); }
    bool isInFrustum(MapPoint* pMP, float viewingCosLimit);
    bool PosInGrid(const cv::KeyPoint &kp, int &posX, int &posY);
    vector<size_t> GetFeaturesInArea(const float &x, const float &y, const float &r, const int minLevel = -1, const int maxLevel = -1, const bool bRight = false) const;
    void ComputeStereoMatches();
    bool isInFrustumChecks(MapPoint* pMP, float viewingCosLimit, bool bRight = false);
    void UndistortKeyPoints();
    void AssignFeaturesToGrid();

    cv::Mat mRwc;
    cv::Mat mOw;
    ORBVocabulary* mpORBvocabulary;
    ORBextractor *mpORBextractorLeft, *mpORBextractorRight;
    double mTimeStamp;
    cv::Mat mK;
    static float fx, fy, cx, cy, invfx, invfy;
    cv::Mat mDistCoef;
    float mbf, mb, mThDepth;
    int N;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight;
    std::vector<cv::KeyPoint> mvKeysUn;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<float> mvuRight;
    std::vector<float> mvDepth;
    DBoW2::BowVector mBowVec;
    DBoW2::FeatureVector mFeatVec;
    cv::Mat mDescriptors, mDescriptorsRight;
    std::vector<bool> mvbOutlier;
    int mnCloseMPs;
    static float mfGridElementWidthInv;
    static float mfGridElementHeightInv;
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    cv::Mat mTcw;
    static long unsigned int nNextId;
    long unsigned int mnId;
    KeyFrame* mpReferenceKF;
    int mnScaleLevels;
    float mfScaleFactor;
    float mfLogScaleFactor;
    vector<float> mvScaleFactors;
    vector<float> mvInvScaleFactors;
    vector<float> mvLevelSigma2;
    vector<float> mvInvLevelSigma2;
    static float mnMinX, mnMaxX, mnMinY, mnMaxY;
    static bool mbInitialComputations;
    uint8_t mnClientId;
    // private in the reference (Frame.h:301-311); the extracted bodies are members, so access is the same
    cv::Mat mRcw;
    cv::Mat mtcw;
    GeometricCamera *mpCamera, *mpCamera2;
    int Nleft, Nright;
    int monoLeft, monoRight;
    std::vector<int> mvLeftToRightMatch, mvRightToLeftMatch;
    std::vector<cv::Mat> mvStereo3Dpoints;
    std::vector<std::size_t> mGridRight[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    cv::Mat mTlr, mRlr, mtlr, mTrl;
};

class KeyFrame               // R/include/KeyFrame.h
{
public:
    // Frame -> KeyFrame as R/src/KeyFrame.cc:52-120 copies the members (by hand: the reference constructor also registers the
    // keyframe with the map, the database and the communicator)
    KeyFrame(Frame &F) :
        mnId(nNextId++), mnFrameId(F.mnId), mnClientId(F.mnClientId), mTimeStamp(F.mTimeStamp), mnGridCols(FRAME_GRID_COLS), mnGridRows(FRAME_GRID_ROWS),
        mfGridElementWidthInv(F.mfGridElementWidthInv), mfGridElementHeightInv(F.mfGridElementHeightInv),
        fx(F.fx), fy(F.fy), cx(F.cx), cy(F.cy), invfx(F.invfx), invfy(F.invfy), mbf(F.mbf), mb(F.mb), mThDepth(F.mThDepth), N(F.N),
        mvKeys(F.mvKeys), mvKeysUn(F.mvKeysUn), mvuRight(F.mvuRight), mvDepth(F.mvDepth), mDescriptors(F.mDescriptors.clone()),
        mBowVec(F.mBowVec), mFeatVec(F.mFeatVec), mnScaleLevels(F.mnScaleLevels), mfScaleFactor(F.mfScaleFactor),
        mfLogScaleFactor(F.mfLogScaleFactor), mvScaleFactors(F.mvScaleFactors), mvLevelSigma2(F.mvLevelSigma2),
        mvInvLevelSigma2(F.mvInvLevelSigma2), mnMinX(F.mnMinX), mnMinY(F.mnMinY), mnMaxX(F.mnMaxX), mnMaxY(F.mnMaxY), mK(F.mK),
        mvpMapPoints(F.mvpMapPoints), mpORBvocabulary(F.mpORBvocabulary), mbBad(false), mHalfBaseline(F.mb / 2), mSysState(NOTYPE),
        mpCamera(F.mpCamera), mpCamera2(F.mpCamera2), mTlr(F.mTlr.clone()), mTrl(F.mTrl.clone()), mvKeysRight(F.mvKeysRight),
        NLeft(F.Nleft), NRight(F.Nright)
    {
        mGrid.resize(mnGridCols);
        if (F.Nleft != -1) mGridRight.resize(mnGridCols);
        for (int i = 0; i < mnGridCols; i++) {
            mGrid[i].resize(mnGridRows);
            if (F.Nleft != -1) mGridRight[i].resize(mnGridRows);
            for (int j = 0; j < mnGridRows; j++) {
                mGrid[i][j] = F.mGrid[i][j];
                if (F.Nleft != -1) mGridRight[i][j] = F.mGridRight[i][j];
            }
        }
        if (!F.mTcw.empty()) SetPose(F.mTcw);
    }

    // R/src/KeyFrame.cc:178-227 without the IMU, lock and communication branches
    void SetPose(const cv::Mat &Tcw_)
    {
        unique_lock<mutex> lock(mMutexPose);
        Tcw_.copyTo(Tcw);
        cv::Mat Rcw = Tcw.rowRange(0, 3).colRange(0, 3);
        cv::Mat tcw = Tcw.rowRange(0, 3).col(3);
        cv::Mat Rwc = Rcw.t();
        Ow = -Rwc * tcw;
        Twc = cv::Mat::eye(4, 4, Tcw.type());
        Rwc.copyTo(Twc.rowRange(0, 3).colRange(0, 3));
        Ow.copyTo(Twc.rowRange(0, 3).col(3));
        cv::Mat center = (cv::Mat_<float>(4, 1) << mHalfBaseline, 0, 0, 1);
        Cw = Twc * center;
    }

    // ---- bodies from R/src/KeyFrame.cc ----
    void ComputeBoW();
    cv::Mat GetPose();
    cv::Mat GetPoseInverse();
    cv::Mat GetCameraCenter();
    cv::Mat GetRotation();
    cv::Mat GetTranslation();
    void AddMapPoint(MapPoint* pMP, const size_t &idx);
    void EraseMapPointMatch(const int &idx);
    void EraseMapPointMatch(MapPoint* pMP);
    void ReplaceMapPointMatch(const int &idx, MapPoint* pMP);
    std::set<MapPoint*> GetMapPoints();
    std::vector<MapPoint*> GetMapPointMatches();
    MapPoint* GetMapPoint(const size_t &idx);
    std::vector<size_t> GetFeaturesInArea(const float &x, const float &y, const float &r, const bool bRight = false) const;
    bool IsInImage(const float &x, const float &y) const;
    bool isBad();
    cv::Mat GetRightPose();
    cv::Mat GetRightCameraCenter();
    cv::Mat GetRightRotation();
    cv::Mat GetRightTranslation();

    static long unsigned int nNextId;
    long unsigned int mnId;
    const long unsigned int mnFrameId;
    uint8_t mnClientId;
    const double mTimeStamp;
    const int mnGridCols;
    const int mnGridRows;
    const float mfGridElementWidthInv;
    const float mfGridElementHeightInv;
    const float fx, fy, cx, cy, invfx, invfy, mbf, mb, mThDepth;
    const int N;
    const std::vector<cv::KeyPoint> mvKeys;
    std::vector<cv::KeyPoint> mvKeysUn;
    std::vector<float> mvuRight;
    std::vector<float> mvDepth;
    cv::Mat mDescriptors;
    DBoW2::BowVector mBowVec;
    DBoW2::FeatureVector mFeatVec;
    const int mnScaleLevels;
    const float mfScaleFactor;
    const float mfLogScaleFactor;
    std::vector<float> mvScaleFactors;
    std::vector<float> mvLevelSigma2;
    std::vector<float> mvInvLevelSigma2;
    const int mnMinX, mnMinY, mnMaxX, mnMaxY;
    cv::Mat mK;
    // protected in the reference
    cv::Mat Tcw, Twc, Ow, Cw;
    std::vector<MapPoint*> mvpMapPoints;
    ORBVocabulary* mpORBvocabulary;
    std::vector<std::vector<std::vector<size_t> > > mGrid;
    bool mbBad;
    float mHalfBaseline;
    eSystemState mSysState;
    std::mutex mMutexPose, mMutexConnections, mMutexFeatures;
    GeometricCamera *mpCamera, *mpCamera2;
    cv::Mat mTlr;
    cv::Mat mTrl;
    const std::vector<cv::KeyPoint> mvKeysRight;
    int NLeft, NRight;
    std::vector<std::vector<std::vector<size_t> > > mGridRight;
};

class MapPoint               // R/include/MapPoint.h
{
public:
    // by hand: the reference constructors register the point with the map / communicator (R/src/MapPoint.cc:19-122)
    MapPoint(const cv::Mat &Pos, KeyFrame* pRefKF, Map* pMap) :
        mnId(nNextId++), mnClientId(0), mnFirstKFid(pRefKF ? (long)pRefKF->mnId : -1), mnFirstFrame(pRefKF ? (long)pRefKF->mnFrameId : -1), nObs(0),
        mTrackProjX(0), mTrackProjY(0), mTrackDepth(0), mTrackDepthR(0), mTrackProjXR(0), mTrackProjYR(0), mbTrackInView(false), mbTrackInViewR(false),
        mnTrackScaleLevel(0), mnTrackScaleLevelR(0), mTrackViewCos(0), mTrackViewCosR(0), mnTrackReferenceForFrame(0), mnLastFrameSeen(0),
        mnFuseCandidateForKF(0), mpRefKF(pRefKF), mnVisible(1), mnFound(1), mbBad(false), mpReplaced(NULL), mfMinDistance(0), mfMaxDistance(0),
        mpMap(pMap), mSysState(NOTYPE), mMaxObsKFId(0)
    {
        Pos.copyTo(mWorldPos);
        mNormalVector = cv::Mat::zeros(3, 1, CV_32F);
    }

    // ---- bodies from R/src/MapPoint.cc ----
    cv::Mat GetWorldPos();
    cv::Mat GetNormal();
    KeyFrame* GetReferenceKeyFrame();
    std::map<KeyFrame*, std::tuple<int, int> > GetObservations();
    int Observations();
    void AddObservation(KeyFrame* pKF, int idx);
    std::tuple<int, int> GetIndexInKeyFrame(KeyFrame* pKF);
    bool IsInKeyFrame(KeyFrame* pKF);
    bool isBad();
    void Replace(MapPoint* pMP);
    MapPoint* GetReplaced();
    void IncreaseVisible(int n = 1);
    void IncreaseFound(int n = 1);
    void ComputeDistinctiveDescriptors();
    cv::Mat GetDescriptor();
    void UpdateNormalAndDepth();
    float GetMinDistanceInvariance();
    float GetMaxDistanceInvariance();
    int PredictScale(const float &currentDist, KeyFrame* pKF);
    int PredictScale(const float &currentDist, Frame* pF);

    long unsigned int mnId;
    static long unsigned int nNextId;
    uint8_t mnClientId;
    long int mnFirstKFid;
    long int mnFirstFrame;
    int nObs;
    float mTrackProjX, mTrackProjY, mTrackDepth, mTrackDepthR, mTrackProjXR, mTrackProjYR;
    bool mbTrackInView, mbTrackInViewR;
    int mnTrackScaleLevel, mnTrackScaleLevelR;
    float mTrackViewCos, mTrackViewCosR;
    long unsigned int mnTrackReferenceForFrame;
    long unsigned int mnLastFrameSeen;
    long unsigned int mnFuseCandidateForKF;
    static std::mutex mGlobalMutex;
    // protected in the reference
    cv::Mat mWorldPos;
    std::map<KeyFrame*, std::tuple<int, int> > mObservations;
    cv::Mat mNormalVector;
    cv::Mat mDescriptor;
    KeyFrame* mpRefKF;
    int mnVisible, mnFound;
    bool mbBad;
    MapPoint* mpReplaced;
    float mfMinDistance, mfMaxDistance;
    Map* mpMap;
    eSystemState mSysState;
    size_t mMaxObsKFId;
    std::mutex mMutexPos, mMutexFeatures;
};

}  // namespace ORB_SLAM3

#include "ORBmatcher.h"
