// Drop-in body of ORB_SLAM3::ORBextractor over the B200 C ABI.  Replaces R/orb_slam3/src/ORBextractor.cc.
#include "ORBextractor.h"
#include <atomic>
#include <cstring>
#include <stdexcept>
#include <string>
#include "orbx.h"

namespace ORB_SLAM3
{

static std::atomic<int> g_orbx_device(0);
static std::atomic<bool> g_pyramid_sync(true);
static thread_local int tl_last_device = -1;
void ORBextractor::SetDevice(int device) { g_orbx_device = device; }
int ORBextractor::DefaultDevice() { return g_orbx_device; }
int ORBextractor::ThreadDevice() { return tl_last_device >= 0 ? tl_last_device : (int)g_orbx_device; }
void ORBextractor::SetPyramidSync(bool on) { g_pyramid_sync = on; }

static void create_handle(orbx_extractor** out, int device, int nfeatures, float scaleFactor, int nlevels, int ini, int min, int w, int h)
{
    orbx_params p;
    p.nfeatures = nfeatures; p.scale_factor = scaleFactor; p.nlevels = nlevels; p.ini_th_fast = ini; p.min_th_fast = min;
    p.max_width = w; p.max_height = h; p.max_batch = 1; p.device = device;
    p.max_candidates_per_level = 1 << 30;      // clamped to the geometric NMS bound of every level: a corner-dense frame can never overflow
    if (orbx_extractor_create(&p, out) != ORBX_OK)
        throw std::runtime_error(std::string("ORBextractor (B200): ") + orbx_last_error());   // no CPU fallback exists
}

// R/src/ORBextractor.cc:408-468: the tables come from the library so that both sides agree bit for bit
ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST):
    nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels),
    iniThFAST(_iniThFAST), minThFAST(_minThFAST), mpHandle(nullptr), mnHandleW(0), mnHandleH(0), mnDevice(g_orbx_device)
{
    // a small probe handle gives the tables without knowing the camera resolution yet
    orbx_extractor* probe = nullptr;
    create_handle(&probe, mnDevice, nfeatures, (float)scaleFactor, nlevels, iniThFAST, minThFAST, 64, 64);
    mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels);
    mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels); mnFeaturesPerLevel.resize(nlevels);
    orbx_extractor_tables(probe, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(),
                          mvInvLevelSigma2.data(), mnFeaturesPerLevel.data());
    orbx_extractor_destroy(probe);
    mvImagePyramid.resize(nlevels);
}

ORBextractor::~ORBextractor()
{
    if (mpHandle) orbx_extractor_destroy(mpHandle);
}

// R/src/ORBextractor.cc:1068-1150
int ORBextractor::operator()( cv::InputArray _image, cv::InputArray _mask, std::vector<cv::KeyPoint>& _keypoints,
                              cv::OutputArray _descriptors, std::vector<int> &vLappingArea)
{
    (void)_mask;                                       // ignored by the reference too (ORBextractor.h:60)
    if(_image.empty())
        return -1;

    cv::Mat image = _image.getMat();
    // assert(image.type() == CV_8UC1) in the reference (:1076)

    if (!mpHandle || image.cols > mnHandleW || image.rows > mnHandleH) {
        if (mpHandle) orbx_extractor_destroy(mpHandle);
        mpHandle = nullptr;
        mnHandleW = image.cols; mnHandleH = image.rows;
        create_handle(&mpHandle, mnDevice, nfeatures, (float)scaleFactor, nlevels, iniThFAST, minThFAST, mnHandleW, mnHandleH);
    }
    tl_last_device = mnDevice;                         // ORBmatcher calls of this thread follow the extractor's device
    const int cap = orbx_extractor_max_keypoints(mpHandle);
    _keypoints.resize(cap);                            // cv::KeyPoint is layout-compatible with orbx_keypoint
    std::vector<unsigned char> desc((size_t)cap * 32);
    int n = 0, mono = 0;
    const int rc = orbx_extract(mpHandle, image.ptr(0), image.cols, image.rows, (int)image.step,
                                vLappingArea[0], vLappingArea[1],
                                reinterpret_cast<orbx_keypoint*>(_keypoints.data()), desc.data(), cap, &n, &mono);
    if (rc == ORBX_E_EMPTY) return -1;
    if (rc != ORBX_OK) throw std::runtime_error(std::string("ORBextractor (B200): ") + orbx_last_error());
    _keypoints.resize(n);
    if (n == 0) _descriptors.release();
    else {
        _descriptors.create(n, 32, CV_8U);
        cv::Mat d = _descriptors.getMat();
        for (int i = 0; i < n; i++) std::memcpy(d.ptr(i), desc.data() + (size_t)i * 32, 32);
    }
    // mvImagePyramid as the reference leaves it (Frame::ComputeStereoMatches reads it), unless the integration keeps it on the GPU
    if (g_pyramid_sync) {
        // level 0 is the input itself (the reference keeps a bordered copy of it, :1165-1172): host copy; levels 1.. in one round trip
        mvImagePyramid[0].create(image.rows, image.cols, CV_8UC1);
        for (int y = 0; y < image.rows; y++) std::memcpy(mvImagePyramid[0].ptr(y), image.ptr(y), (size_t)image.cols);
        DownloadLevels(1);
    }
    else for (auto& m : mvImagePyramid) m.release();   // stale until SyncPyramidToHost()
    return mono;
}

// explicit download of mvImagePyramid (R/include/ORBextractor.h:88); only the stereo SAD refinement reads it
void ORBextractor::SyncPyramidToHost()
{
    DownloadLevels(0);
}

// levels first .. nlevels-1 of the last frame into mvImagePyramid: one queue of copies into the handle's pinned staging, one
// synchronisation, and cv::Mat HEADERS over that staging (no second copy).  The headers stay valid until the next frame of this
// extractor, like the reference's own entries, which ComputePyramid re-creates every frame (R/src/ORBextractor.cc:1150-1177) and
// which only Frame::ComputeStereoMatches reads, right after the extraction (R/src/Frame.cc:792, :882-901).
void ORBextractor::DownloadLevels(int first)
{
    if (!mpHandle || first >= nlevels) return;
    const int cnt = nlevels - first;
    std::vector<const uint8_t*> ptr(cnt); std::vector<int> stride(cnt);
    if (orbx_pyramid_levels_staged(mpHandle, 0, first, cnt, ptr.data(), stride.data()) != ORBX_OK)
        throw std::runtime_error(std::string("ORBextractor (B200): ") + orbx_last_error());
    for (int l = first; l < nlevels; l++) {
        int w = 0, h = 0;
        if (orbx_pyramid_level_size(mpHandle, l, &w, &h) != ORBX_OK) return;
        mvImagePyramid[l] = cv::Mat(h, w, CV_8UC1, const_cast<uint8_t*>(ptr[l - first]), (size_t)stride[l - first]);
    }
}

} //namespace ORB_SLAM
