/**
* Drop-in replacement of ORB-SLAM3's include/ORBextractor.h (class ORB_SLAM3::ORBextractor) for the B200 front-end.
*
* The PUBLIC interface reproduces the reference's member for member (R/orb_slam3/include/ORBextractor.h:47-113: constructor,
* operator(), the six getters, the public mvImagePyramid); that declaration is part of ORB-SLAM3:
*   Copyright (C) 2017-2020 Carlos Campos, Richard Elvira, Juan J. Gómez Rodríguez, José M.M. Montiel and Juan D. Tardós, University of Zaragoza.
*   Copyright (C) 2014-2016 Raúl Mur-Artal, José M.M. Montiel and Juan D. Tardós, University of Zaragoza.
* ORB-SLAM3 is free software under the GNU General Public License v3 (or later); this interface declaration is used under
* the same license.  tests/test_dropin_signatures.py checks that every public signature here equals the reference's.
*
* The private part is a handle of the B200 C ABI (include/orbx.h) instead of the reference's tables; every caller in the
* reference (Frame.cc:80-86, :397-399, :792-901; Tracking.cc:145-151) is compiled from source in the same package, so the
* changed private layout is safe.
*/
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <vector>
#include <list>
#if (CV_MAJOR_VERSION > 3)
#include <opencv2/opencv.hpp>
#else
#include <opencv/cv.h>
#endif

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

    std::vector<cv::Mat> mvImagePyramid;

    // ---- additions of the B200 drop-in (not in the reference class) ----
    // The pyramid lives on the GPU.  By default operator() also downloads it into mvImagePyramid, exactly as the reference leaves
    // it (Frame::ComputeStereoMatches reads it, R/src/Frame.cc:792, :882-901).  A stereo integration that calls
    // ORBmatcher::ComputeStereoMatches (device-side) turns the download off and saves 0.8 MB of D2H per 752x480 frame (levels 1 .. 7; level 0 is copied from the input on the host);
    // SyncPyramidToHost() then fetches the levels of the last frame on demand.
    static void SetPyramidSync(bool on);
    void SyncPyramidToHost();
    void DownloadLevels(int first);          // levels first .. nlevels-1 of the last frame into mvImagePyramid, one round trip
    // The GPU that extractors constructed from now on live on (multi-agent boxes: one agent per GPU); default 0
    static void SetDevice(int device);
    static int DefaultDevice();
    // the device of the extractor the calling thread used last (what ORBmatcher calls of that thread follow), else the default
    static int ThreadDevice();
    int device() const { return mnDevice; }
    // C-ABI handle of this extractor (NULL before the first frame)
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
    int mnDevice;
};

} //namespace ORB_SLAM

#endif
