// Drop-in replacement of R/orb_slam3/include/ORBmatcher.h:35-108 (class ORB_SLAM3::ORBmatcher) for the methods on
// the B200 hot path.  Public signatures are the reference's.  Methods not listed here (SearchByBoW, the Sim3 /
// KeyFrame SearchByProjection overloads, SearchForTriangulation, SearchBySim3, Fuse) keep the reference's own
// bodies from ORBmatcher.cc: they are host-side candidate gathering around DescriptorDistance and are marked
// "glue" in SURVEY.md section 8a; see INTEGRATION.md for how both translation units live side by side.
#ifndef ORBMATCHER_H
#define ORBMATCHER_H

#include <vector>
#include "frame_shim.h"

struct orbx_matcher;   // include/orbx.h

namespace ORB_SLAM3
{

class ORBextractor;

class ORBmatcher
{
public:

    ORBmatcher(float nnratio=0.6, bool checkOri=true);

    // Computes the Hamming distance between two ORB descriptors
    static int DescriptorDistance(const cv::Mat &a, const cv::Mat &b);

    // Search matches between Frame keypoints and projected MapPoints. Returns number of matches
    // Used to track the local map (Tracking)
    int SearchByProjection(Frame &F, const std::vector<MapPoint*> &vpMapPoints, const float th=3, const bool bFarPoints = false, const float thFarPoints = 50.0f);

    // Project MapPoints tracked in last frame into the current frame and search matches.
    // Used to track from previous frame (Tracking)
    int SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono);

    // Search matches between MapPoints in a KeyFrame and ORB in a Frame.
    // Brute force constrained to ORB that belong to the same vocabulary node (at a certain level)
    // Used in Relocalisation and Loop Detection
    int SearchByBoW(KeyFrame *pKF, Frame &F, std::vector<MapPoint*> &vpMapPointMatches);
    int SearchByBoW(KeyFrame *pKF1, KeyFrame* pKF2, std::vector<MapPoint*> &vpMatches12);

    // Matching for the Map Initialization (only used in the monocular case)
    int SearchForInitialization(Frame &F1, Frame &F2, std::vector<cv::Point2f> &vbPrevMatched, std::vector<int> &vnMatches12, int windowSize=10);

    // Addition (not in the reference class): the body of Frame::ComputeStereoMatches (R/src/Frame.cc:785-962) on the
    // device-resident results and pyramids of the frame's two extractors, so that mvImagePyramid never leaves the GPU.
    // Frame::ComputeStereoMatches() becomes: ORBmatcher::ComputeStereoMatches(mpORBextractorLeft, mpORBextractorRight,
    // mb, mbf, mvuRight, mvDepth);  (N = mvKeys.size() entries each, -1 where there is no match)
    static void ComputeStereoMatches(ORBextractor* pLeft, ORBextractor* pRight, float mb, float mbf,
                                     std::vector<float> &mvuRight, std::vector<float> &mvDepth);

public:

    static const int TH_LOW;
    static const int TH_HIGH;
    static const int HISTO_LENGTH;

protected:

    float RadiusByViewingCos(const float &viewCos);

    float mfNNratio;
    bool mbCheckOrientation;
};

}// namespace ORB_SLAM

#endif // ORBMATCHER_H
