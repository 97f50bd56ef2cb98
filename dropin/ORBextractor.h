// Drop-in replacement of R/orb_slam3/include/ORBextractor.h:47-113 (class ORB_SLAM3::ORBextractor).
// The PUBLIC interface is the reference's, member for member: constructor, operator(), the six getters and the
// public mvImagePyramid.  The private part is a handle of the B200 C ABI (include/orbx.h) instead of the
// reference's tables; every caller in the reference (Frame.cc:80-86, :397-399, :792-901; Tracking.cc:145-151)
// is compiled from source in the same package, so the changed private layout is safe.
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <list>
#include <vector>
#include "cv_shim.h"

struct orbx_extractor;   // include/orbx.h

namespace ORB_SLAM3
{

class ORBextractor
{
public:

    enum {HARRIS_SCORE=0, FAST_SCORE=1 };

    ORBextractor(int nfeatures, float scaleFactor, int nlevels,
                 int iniThFAST, int minThFAST);

    ~ORBextractor();

    // Compute the ORB features and descriptors on an image.
    // ORB are dispersed on the image using an octree.
    // Mask is ignored in the current implementation.
    int operator()( cv::InputArray _image, cv::InputArray _mask,
                    std::vector<cv::KeyPoint>& _keypoints,
                    cv::OutputArray _descriptors, std::vector<int> &vLappingArea);

    int inline GetLevels(){
        return nlevels;}

    float inline GetScaleFactor(){
        return scaleFactor;}

    std::vector<float> inline GetScaleFactors(){
        return mvScaleFactor;
    }

    std::vector<float> inline GetInverseScaleFactors(){
        return mvInvScaleFactor;
    }

    std::vector<float> inline GetScaleSigmaSquares(){
        return mvLevelSigma2;
    }

    std::vector<float> inline GetInverseScaleSigmaSquares(){
        return mvInvLevelSigma2;
    }

    // The pyramid of the last frame.  It lives on the GPU; the host copies are refreshed lazily by
    // SyncPyramidToHost(), which the stereo SAD refinement (Frame.cc:871-946) must call before it reads them.
    std::vector<cv::Mat> mvImagePyramid;
    void SyncPyramidToHost();

    // device selection for multi-agent boxes: one agent (= one ORBextractor pair) per GPU
    static void SetDevice(int device);
    // C-ABI handle of this extractor (for orbx_stereo_matches, which replaces Frame::ComputeStereoMatches)
    orbx_extractor* handle() { return mpHandle; }

protected:

    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;

    std::vector<int> mnFeaturesPerLevel;

    std::vector<float> mvScaleFactor;
    std::vector<float> mvInvScaleFactor;
    std::vector<float> mvLevelSigma2;
    std::vector<float> mvInvLevelSigma2;

    orbx_extractor* mpHandle;      // created lazily at the first frame (the image size is not a ctor argument)
    int mnHandleW, mnHandleH;
    std::vector<unsigned char> mvTightImage;
};

} //namespace ORB_SLAM

#endif
