// ref_extractor_api.cc - flat-array entry points around the reference's ORB_SLAM3::ORBextractor (TEST INFRASTRUCTURE).
// The class itself comes from /root/reference/.../src/ORBextractor.cc, compiled unmodified (see Makefile).
#include <list>
#include <vector>
#include <opencv2/core/core.hpp>
#include "ORBextractor.h"
#include "ref_alloc.h"
#include "ref_api.h"

namespace {
// the protected stages, made callable
class Extractor : public ORB_SLAM3::ORBextractor {
public:
    Extractor(int nf, float sf, int nl, int ini, int mn) : ORB_SLAM3::ORBextractor(nf, sf, nl, ini, mn) {}
    using ORB_SLAM3::ORBextractor::ComputePyramid;
    using ORB_SLAM3::ORBextractor::ComputeKeyPointsOctTree;
    using ORB_SLAM3::ORBextractor::DistributeOctTree;
    const std::vector<int>& quotas() const { return mnFeaturesPerLevel; }
    const std::vector<int>& umaxv() const { return umax; }
};
// the node type whose addresses ORBextractor.cc:682 sorts by
constexpr size_t kNodeBytes = sizeof(std::_List_node<ORB_SLAM3::ExtractorNode>);
struct ArenaScope { ArenaScope() { ref_arena_begin(kNodeBytes); } ~ArenaScope() { ref_arena_end(); } };
}

struct RefExtractor { Extractor ex; RefExtractor(int a, float b, int c, int d, int e) : ex(a, b, c, d, e) {} };

extern "C" int ref_uses_arena(void)
{
#ifdef REF_NO_ARENA
    return 0;
#else
    return 1;
#endif
}

extern "C" RefExtractor* ref_extractor_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th)
{
    return new RefExtractor(nfeatures, scale_factor, nlevels, ini_th, min_th);
}
extern "C" void ref_extractor_destroy(RefExtractor* e) { delete e; }

extern "C" void ref_extractor_tables(RefExtractor* e, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                                     int32_t* fpl, int32_t* umax16)
{
    const int nl = e->ex.GetLevels();
    std::vector<float> a = e->ex.GetScaleFactors(), b = e->ex.GetInverseScaleFactors(), c = e->ex.GetScaleSigmaSquares(),
                       d = e->ex.GetInverseScaleSigmaSquares();
    for (int i = 0; i < nl; i++) {
        if (scale) scale[i] = a[i];
        if (inv_scale) inv_scale[i] = b[i];
        if (sigma2) sigma2[i] = c[i];
        if (inv_sigma2) inv_sigma2[i] = d[i];
        if (fpl) fpl[i] = e->ex.quotas()[i];
    }
    if (umax16) for (int i = 0; i < 16; i++) umax16[i] = e->ex.umaxv()[i];
}

extern "C" int ref_extract(RefExtractor* e, const uint8_t* img, int w, int h, int stride, int lap0, int lap1,
                           OrcKeyPoint* kps, uint8_t* desc, int cap, int* n_out)
{
    ArenaScope scope;
    cv::Mat image = (img && w > 0 && h > 0) ? cv::Mat(h, w, CV_8UC1, (void*)img, (size_t)stride) : cv::Mat();
    std::vector<cv::KeyPoint> keys;
    cv::Mat descriptors;
    std::vector<int> lap = {lap0, lap1};
    const int mono = e->ex(image, cv::Mat(), keys, descriptors, lap);
    const int n = (int)keys.size();
    if (n_out) *n_out = n;
    if (mono < 0) { if (n_out) *n_out = 0; return mono; }
    for (int i = 0; i < n && i < cap; i++) {
        static_assert(sizeof(cv::KeyPoint) == sizeof(OrcKeyPoint), "keypoint layouts");
        memcpy(&kps[i], &keys[i], sizeof(OrcKeyPoint));
        memcpy(desc + (size_t)i * 32, descriptors.ptr(i), 32);
    }
    return mono;
}

extern "C" int ref_level_size(RefExtractor* e, int level, int* w, int* h)
{
    if (level < 0 || level >= e->ex.GetLevels() || e->ex.mvImagePyramid[level].empty()) return -1;
    if (w) *w = e->ex.mvImagePyramid[level].cols;
    if (h) *h = e->ex.mvImagePyramid[level].rows;
    return 0;
}

extern "C" int ref_level_image(RefExtractor* e, int level, uint8_t* dst, int dst_stride)
{
    if (level < 0 || level >= e->ex.GetLevels() || e->ex.mvImagePyramid[level].empty()) return -1;
    const cv::Mat& m = e->ex.mvImagePyramid[level];
    for (int y = 0; y < m.rows; y++) memcpy(dst + (size_t)y * dst_stride, m.ptr(y), m.cols);
    return 0;
}

extern "C" int ref_octree_keypoints(RefExtractor* e, const uint8_t* img, int w, int h, int stride, int level, OrcKeyPoint* kps, int cap)
{
    ArenaScope scope;
    cv::Mat image(h, w, CV_8UC1, (void*)img, (size_t)stride);
    e->ex.ComputePyramid(image);
    std::vector<std::vector<cv::KeyPoint> > all;
    e->ex.ComputeKeyPointsOctTree(all);
    if (level < 0 || level >= (int)all.size()) return -1;
    const int n = (int)all[level].size();
    for (int i = 0; i < n && i < cap; i++) memcpy(&kps[i], &all[level][i], sizeof(OrcKeyPoint));
    return n;
}

extern "C" int ref_distribute_octree(RefExtractor* e, const float* xyr, int n, int minX, int maxX, int minY, int maxY, int N, int level,
                                     float* out_xyr, int cap)
{
    ArenaScope scope;
    std::vector<cv::KeyPoint> in(n);
    for (int i = 0; i < n; i++) in[i] = cv::KeyPoint(xyr[3 * i], xyr[3 * i + 1], 7.f, -1.f, xyr[3 * i + 2]);
    std::vector<cv::KeyPoint> out = e->ex.DistributeOctTree(in, minX, maxX, minY, maxY, N, level);
    for (int i = 0; i < (int)out.size() && i < cap; i++) { out_xyr[3 * i] = out[i].pt.x; out_xyr[3 * i + 1] = out[i].pt.y; out_xyr[3 * i + 2] = out[i].response; }
    return (int)out.size();
}

// for ref_matcher_api.cc (Frame::ComputeStereoMatches reads mpORBextractorLeft/Right->mvImagePyramid)
ORB_SLAM3::ORBextractor* ref_extractor_object(RefExtractor* e) { return &e->ex; }
