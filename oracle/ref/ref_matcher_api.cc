// ref_matcher_api.cc - flat-array entry points around the reference's OWN ORBmatcher / Frame / MapPoint / DBoW2 code
// (TEST INFRASTRUCTURE).  Every function builds the reference's objects (stand-in declarations of dropin/shim/orbslam_world.h,
// bodies from /root/reference), calls the reference method, and flattens the result into the same layout as the C restatement
// oracle/orb_oracle.h uses, so that tests can demand  _ref == oracle == GPU.
#include <cstdio>
#include <fstream>
#include <memory>
#include <unistd.h>
#include "orbslam_world.h"
#include "ref_api.h"

using namespace ORB_SLAM3;

namespace {

static_assert(sizeof(cv::KeyPoint) == sizeof(OrcKeyPoint), "keypoint layouts");

void set_bounds(float minX, float maxX, float minY, float maxY)
{
    // R/src/Frame.cc:314-327 (first-frame computations of the constructors)
    Frame::mnMinX = minX; Frame::mnMaxX = maxX; Frame::mnMinY = minY; Frame::mnMaxY = maxY;
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(Frame::mnMaxX - Frame::mnMinX);
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(Frame::mnMaxY - Frame::mnMinY);
}

void fill_frame(Frame& F, const OrcKeyPoint* k, const uint8_t* d, int n)
{
    F.N = n;
    F.mvKeys.resize(n);
    if (n) memcpy((void*)F.mvKeys.data(), k, sizeof(OrcKeyPoint) * (size_t)n);
    F.mvKeysUn = F.mvKeys;
    F.mDescriptors = cv::Mat(n, 32, CV_8U);
    if (n) memcpy(F.mDescriptors.data, d, (size_t)32 * n);
    F.mvpMapPoints.assign(n, static_cast<MapPoint*>(NULL));
    F.mvbOutlier.assign(n, false);
    F.mvuRight.assign(n, -1.f);
    F.mvDepth.assign(n, -1.f);
    F.AssignFeaturesToGrid();
}

void set_scale_tables(Frame& F, const float* scale, int nlevels)
{
    F.mnScaleLevels = nlevels;
    F.mvScaleFactors.assign(scale, scale + nlevels);
    F.mvInvScaleFactors.resize(nlevels); F.mvLevelSigma2.resize(nlevels); F.mvInvLevelSigma2.resize(nlevels);
    for (int i = 0; i < nlevels; i++) {
        F.mvInvScaleFactors[i] = 1.0f / scale[i];
        F.mvLevelSigma2[i] = scale[i] * scale[i];
        F.mvInvLevelSigma2[i] = 1.0f / F.mvLevelSigma2[i];
    }
    F.mfScaleFactor = nlevels > 1 ? scale[1] : 1.2f;
    F.mfLogScaleFactor = std::log(F.mfScaleFactor);
}

void fill_feature_vector(DBoW2::FeatureVector& fv, const int32_t* nodes, const int32_t* start, const int32_t* feat, int nfv)
{
    for (int i = 0; i < nfv; i++)
        for (int j = start[i]; j < start[i + 1]; j++) fv.addFeature((DBoW2::NodeId)nodes[i], (unsigned)feat[j]);
}

}  // namespace

extern "C" int ref_hamming256(const uint8_t* a, const uint8_t* b)
{
    cv::Mat A(1, 32, CV_8U, (void*)a), B(1, 32, CV_8U, (void*)b);
    return ORBmatcher::DescriptorDistance(A, B);
}

// Frame::AssignFeaturesToGrid + Frame::GetFeaturesInArea (R/src/Frame.cc:360-391, 628-709)
extern "C" int ref_features_in_area(const OrcKeyPoint* kps, int n, float minX, float maxX, float minY, float maxY,
                                    float x, float y, float r, int minLevel, int maxLevel, int32_t* out, int cap)
{
    set_bounds(minX, maxX, minY, maxY);
    Frame F;
    std::vector<uint8_t> zero((size_t)32 * (n > 0 ? n : 1));
    fill_frame(F, kps, zero.data(), n);
    const std::vector<size_t> v = F.GetFeaturesInArea(x, y, r, minLevel, maxLevel);
    for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = (int32_t)v[i];
    return (int)v.size();
}

// ORBmatcher::SearchForInitialization (R/src/ORBmatcher.cc:702-817)
extern "C" int ref_search_for_initialization(const OrcKeyPoint* k1, const uint8_t* d1, int n1, const OrcKeyPoint* k2, const uint8_t* d2, int n2,
                                             float minX, float maxX, float minY, float maxY, float* prev_xy, int32_t* matches12, int window,
                                             float nnratio, int check_ori)
{
    set_bounds(minX, maxX, minY, maxY);
    Frame F1, F2;
    fill_frame(F1, k1, d1, n1); fill_frame(F2, k2, d2, n2);
    std::vector<cv::Point2f> prev(n1);
    for (int i = 0; i < n1; i++) prev[i] = cv::Point2f(prev_xy[2 * i], prev_xy[2 * i + 1]);
    std::vector<int> m12;
    ORBmatcher matcher(nnratio, check_ori != 0);
    const int nm = matcher.SearchForInitialization(F1, F2, prev, m12, window);
    for (int i = 0; i < n1; i++) { matches12[i] = m12[i]; prev_xy[2 * i] = prev[i].x; prev_xy[2 * i + 1] = prev[i].y; }
    return nm;
}

// Frame::ComputeStereoMatches (R/src/Frame.cc:785-962) on the pyramids of two reference extractors' last operator() calls
struct RefExtractor;
ORB_SLAM3::ORBextractor* ref_extractor_object(RefExtractor* e);     // ref_extractor_api.cc
extern "C" void ref_compute_stereo_matches(RefExtractor* left, RefExtractor* right, const OrcKeyPoint* kl, const uint8_t* dl, int nl,
                                           const OrcKeyPoint* kr, const uint8_t* dr, int nr, const float* scale, int nlevels,
                                           float mb, float mbf, float* uright, float* depth)
{
    Frame F;
    F.N = nl;
    F.mvKeys.resize(nl); if (nl) memcpy((void*)F.mvKeys.data(), kl, sizeof(OrcKeyPoint) * (size_t)nl);
    F.mvKeysRight.resize(nr); if (nr) memcpy((void*)F.mvKeysRight.data(), kr, sizeof(OrcKeyPoint) * (size_t)nr);
    F.mDescriptors = cv::Mat(nl, 32, CV_8U); if (nl) memcpy(F.mDescriptors.data, dl, (size_t)32 * nl);
    F.mDescriptorsRight = cv::Mat(nr, 32, CV_8U); if (nr) memcpy(F.mDescriptorsRight.data, dr, (size_t)32 * nr);
    set_scale_tables(F, scale, nlevels);
    F.mb = mb; F.mbf = mbf;
    F.mpORBextractorLeft = ref_extractor_object(left);
    F.mpORBextractorRight = ref_extractor_object(right);
    F.ComputeStereoMatches();
    for (int i = 0; i < nl; i++) { uright[i] = F.mvuRight[i]; depth[i] = F.mvDepth[i]; }
}

// ---- DBoW2 vocabulary (R/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h) built from a node table through its own text loader ----
namespace {
struct OpenVocabulary : public ORBVocabulary {         // the per-feature descent is a protected member
    void descend(const cv::Mat& feature, DBoW2::WordId& id, DBoW2::WordValue& weight, DBoW2::NodeId* nid, int levelsup) const
    {
        ORBVocabulary::transform(feature, id, weight, nid, levelsup);
    }
};
}
struct RefVocab { OpenVocabulary voc; };

extern "C" RefVocab* ref_vocab_create(int n_nodes, const int32_t* parent, const uint8_t* is_leaf, const uint8_t* desc, const double* weight,
                                      int k, int L)
{
    // ORBvoc.txt layout (TemplatedVocabulary::loadFromTextFile): "k L scoring weighting", then one line per node after the root:
    // parent, isLeaf, 32 descriptor bytes, weight.  L1_NORM = 0, TF_IDF = 0.
    char path[] = "/tmp/ref_vocab_XXXXXX";
    const int fd = mkstemp(path);
    if (fd < 0) return NULL;
    FILE* f = fdopen(fd, "w");
    fprintf(f, "%d %d 0 0\n", k, L);
    for (int i = 1; i < n_nodes; i++) {
        fprintf(f, "%d %d ", parent[i], is_leaf[i] ? 1 : 0);
        for (int j = 0; j < 32; j++) fprintf(f, "%d ", desc[(size_t)i * 32 + j]);
        fprintf(f, "%.17g%s", weight[i], i + 1 < n_nodes ? "\n" : "");
    }
    fclose(f);
    RefVocab* v = new RefVocab();
    const bool ok = v->voc.loadFromTextFile(path);
    unlink(path);
    if (!ok) { delete v; return NULL; }
    return v;
}
extern "C" void ref_vocab_destroy(RefVocab* v) { delete v; }

// transform(features, BowVector&, FeatureVector&, levelsup) (:1127-1200), flattened like orc_bow_transform
extern "C" int ref_bow_transform(RefVocab* v, const uint8_t* desc, int n, int levelsup, int32_t* bow_words, double* bow_values,
                                 int32_t* fv_nodes, int32_t* fv_start, int32_t* fv_features, int* n_fv)
{
    cv::Mat D(n, 32, CV_8U, (void*)desc);
    std::vector<cv::Mat> feats = Converter::toDescriptorVector(D);
    DBoW2::BowVector bv; DBoW2::FeatureVector fv;
    v->voc.transform(feats, bv, fv, levelsup);
    int i = 0;
    for (DBoW2::BowVector::const_iterator it = bv.begin(); it != bv.end(); ++it, ++i) { bow_words[i] = (int32_t)it->first; bow_values[i] = it->second; }
    int j = 0, pos = 0;
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it, ++j) {
        fv_nodes[j] = (int32_t)it->first; fv_start[j] = pos;
        for (size_t t = 0; t < it->second.size(); t++) fv_features[pos++] = (int32_t)it->second[t];
    }
    fv_start[j] = pos;
    if (n_fv) *n_fv = j;
    return i;
}

// per-feature descent: transform(feature, word id, weight, node id, levelsup) (:1218-1259)
extern "C" void ref_bow_transform_features(RefVocab* v, const uint8_t* desc, int n, int levelsup, int32_t* word_id, double* weight, int32_t* node_id)
{
    for (int i = 0; i < n; i++) {
        cv::Mat D(1, 32, CV_8U, (void*)(desc + (size_t)i * 32));
        DBoW2::WordId id = 0; DBoW2::WordValue w = 0; DBoW2::NodeId nid = 0;
        v->voc.descend(D, id, w, &nid, levelsup);
        word_id[i] = (int32_t)id; weight[i] = w; node_id[i] = (int32_t)nid;
    }
}

// ORBmatcher::SearchByBoW (R/src/ORBmatcher.cc:269-471 KeyFrame-Frame, :819-959 KeyFrame-KeyFrame), flattened like orc_search_by_bow
extern "C" int ref_search_by_bow(int mode, const OrcKeyPoint* k1, const uint8_t* d1, const uint8_t* valid1, int n1,
                                 const int32_t* fv1_nodes, const int32_t* fv1_start, const int32_t* fv1_feat, int nfv1,
                                 const OrcKeyPoint* k2, const uint8_t* d2, const uint8_t* valid2, int n2,
                                 const int32_t* fv2_nodes, const int32_t* fv2_start, const int32_t* fv2_feat, int nfv2,
                                 float nnratio, int check_ori, int32_t* matches12)
{
    set_bounds(0, 4096, 0, 4096);
    Frame F1, F2;
    fill_frame(F1, k1, d1, n1); fill_frame(F2, k2, d2, n2);
    fill_feature_vector(F1.mFeatVec, fv1_nodes, fv1_start, fv1_feat, nfv1);
    fill_feature_vector(F2.mFeatVec, fv2_nodes, fv2_start, fv2_feat, nfv2);
    Map map;
    cv::Mat origin = cv::Mat::zeros(3, 1, CV_32F);
    std::vector<std::unique_ptr<MapPoint> > own;
    std::map<MapPoint*, int> index1;
    for (int i = 0; i < n1; i++)
        if (valid1[i]) { own.emplace_back(new MapPoint(origin, NULL, &map)); F1.mvpMapPoints[i] = own.back().get(); index1[own.back().get()] = i; }
    if (mode == 1)
        for (int i = 0; i < n2; i++)
            if (!valid2 || valid2[i]) { own.emplace_back(new MapPoint(origin, NULL, &map)); F2.mvpMapPoints[i] = own.back().get(); }
    KeyFrame KF1(F1);
    ORBmatcher matcher(nnratio, check_ori != 0);
    for (int i = 0; i < n1; i++) matches12[i] = -1;
    int nm;
    if (mode == 0) {
        // the result is stored per FRAME feature: vpMapPointMatches[i2] = the keyframe's MapPoint
        std::vector<MapPoint*> vp;
        nm = matcher.SearchByBoW(&KF1, F2, vp);
        for (int i2 = 0; i2 < (int)vp.size(); i2++)
            if (vp[i2]) matches12[index1[vp[i2]]] = i2;
    } else {
        KeyFrame KF2(F2);
        std::map<MapPoint*, int> index2;
        for (int i = 0; i < n2; i++) if (F2.mvpMapPoints[i]) index2[F2.mvpMapPoints[i]] = i;
        std::vector<MapPoint*> vp;
        nm = matcher.SearchByBoW(&KF1, &KF2, vp);
        for (int i1 = 0; i1 < (int)vp.size(); i1++)
            if (vp[i1]) matches12[i1] = index2[vp[i1]];
    }
    return nm;
}

// MapPoint::ComputeDistinctiveDescriptors (R/src/MapPoint.cc:448-524) for a batch of points, flattened like
// orc_distinctive_descriptors.  The reference walks std::map<KeyFrame*, ...>, i.e. the observing keyframes in ADDRESS order: the
// keyframes of a point are carved from one buffer in row order so that this order is the row order.
extern "C" void ref_distinctive_descriptors(const uint8_t* desc, const int32_t* offsets, int npoints, int32_t* best)
{
    set_bounds(0, 4096, 0, 4096);
    Map map;
    cv::Mat origin = cv::Mat::zeros(3, 1, CV_32F);
    for (int p = 0; p < npoints; p++) {
        const int n = offsets[p + 1] - offsets[p];
        best[p] = -1;
        if (n <= 0) continue;
        void* raw = ::malloc(sizeof(KeyFrame) * (size_t)n);
        KeyFrame* kfs = static_cast<KeyFrame*>(raw);
        MapPoint mp(origin, NULL, &map);
        for (int i = 0; i < n; i++) {
            Frame F;
            OrcKeyPoint kp = {0, 0, 31.f, 0.f, 1.f, 0, -1};
            fill_frame(F, &kp, desc + (size_t)(offsets[p] + i) * 32, 1);
            new (kfs + i) KeyFrame(F);
            mp.AddObservation(kfs + i, 0);
        }
        mp.ComputeDistinctiveDescriptors();
        const cv::Mat d = mp.GetDescriptor();
        for (int i = 0; i < n && best[p] < 0; i++)
            if (memcmp(d.data, desc + (size_t)(offsets[p] + i) * 32, 32) == 0) best[p] = i;     // first row with the winning bytes
        for (int i = 0; i < n; i++) kfs[i].~KeyFrame();
        ::free(raw);
    }
}
