// Drop-in bodies of the hot-path methods of ORB_SLAM3::ORBmatcher over the B200 C ABI (include/orbx.h).
// Replaces R/orb_slam3/src/ORBmatcher.cc:36-222, 702-817, 1970-2186, 2358-2374.  The geometric part of every
// search (projection, frustum and depth tests) stays on the host exactly as in the reference; what moves to the
// GPU is GetFeaturesInArea + DescriptorDistance + best/second bookkeeping + the rotation histogram.
#include "ORBmatcher.h"
#include <cmath>
#include <mutex>
#include <stdexcept>
#include <string>
#include "../include/orbx.h"
#include "ORBextractor.h"

namespace ORB_SLAM3
{

const int ORBmatcher::TH_HIGH = 100;
const int ORBmatcher::TH_LOW = 50;
const int ORBmatcher::HISTO_LENGTH = 30;

#ifndef ORBX_USE_REAL_OPENCV
float Frame::mnMinX = 0, Frame::mnMaxX = 0, Frame::mnMinY = 0, Frame::mnMaxY = 0;
#endif

// one matcher context per calling thread (Tracking / LocalMapping / LoopClosing call concurrently on different frames)
static orbx_matcher* context()
{
    static thread_local orbx_matcher* ctx = nullptr;
    if (!ctx) {
        orbx_matcher_params p; p.device = 0; p.max_keypoints = 8192; p.max_batch = 1; p.max_candidates = 0;
        if (orbx_matcher_create(&p, &ctx) != ORBX_OK)
            throw std::runtime_error(std::string("ORBmatcher (B200): ") + orbx_last_error());   // no CPU fallback exists
    }
    return ctx;
}
static void check(int rc) { if (rc != ORBX_OK) throw std::runtime_error(std::string("ORBmatcher (B200): ") + orbx_last_error()); }

ORBmatcher::ORBmatcher(float nnratio, bool checkOri): mfNNratio(nnratio), mbCheckOrientation(checkOri) {}

// R/src/ORBmatcher.cc:2358-2374.  A single pair is a scalar accessor (Frame.cc:860, MapPoint.cc:496 call it inside
// host loops); the batched form every search uses is orbx_hamming_pairs / the window kernels.
int ORBmatcher::DescriptorDistance(const cv::Mat &a, const cv::Mat &b)
{
    const uint32_t *pa = a.ptr<uint32_t>(), *pb = b.ptr<uint32_t>();
    int dist = 0;
    for (int i = 0; i < 8; i++) dist += __builtin_popcount(pa[i] ^ pb[i]);
    return dist;
}

float ORBmatcher::RadiusByViewingCos(const float &viewCos) { return viewCos > 0.998 ? 2.5f : 4.0f; }   // :216-222

static const uint8_t* rows32(const cv::Mat& d, std::vector<uint8_t>& tmp)
{
    if (d.isContinuous()) return d.ptr(0);
    tmp.resize((size_t)d.rows * 32);
    for (int i = 0; i < d.rows; i++) std::memcpy(tmp.data() + (size_t)i * 32, d.ptr(i), 32);
    return tmp.data();
}

// R/src/ORBmatcher.cc:702-817
int ORBmatcher::SearchForInitialization(Frame &F1, Frame &F2, std::vector<cv::Point2f> &vbPrevMatched, std::vector<int> &vnMatches12, int windowSize)
{
    const int n1 = (int)F1.mvKeysUn.size(), n2 = (int)F2.mvKeysUn.size();
    vnMatches12 = std::vector<int>(n1, -1);
    if (n1 == 0) return 0;
    const float bounds[4] = {Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY};
    std::vector<uint8_t> t1, t2;
    int nmatches = 0;
    static_assert(sizeof(cv::KeyPoint) == sizeof(orbx_keypoint) && sizeof(cv::Point2f) == 8, "layout");
    check(orbx_search_for_initialization(context(), reinterpret_cast<const orbx_keypoint*>(F1.mvKeysUn.data()), rows32(F1.mDescriptors, t1), n1,
                                         reinterpret_cast<const orbx_keypoint*>(F2.mvKeysUn.data()), rows32(F2.mDescriptors, t2), n2, bounds,
                                         reinterpret_cast<float*>(vbPrevMatched.data()), vnMatches12.data(), windowSize, mfNNratio,
                                         mbCheckOrientation ? 1 : 0, &nmatches));
    return nmatches;
}

// R/src/ORBmatcher.cc:1970-2186, monocular / rectified-stereo branch (Nleft == -1)
int ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono)
{
    const float* Rcw = CurrentFrame.mRcw; const float* tcw = CurrentFrame.mtcw;
    const float* Rlw = LastFrame.mRcw; const float* tlw = LastFrame.mtcw;
    // twc = -Rcw^T tcw ; tlc = Rlw twc + tlw  (:1983-1988)
    float twc[3], tlc[3];
    for (int i = 0; i < 3; i++) twc[i] = -(Rcw[0 * 3 + i] * tcw[0] + Rcw[1 * 3 + i] * tcw[1] + Rcw[2 * 3 + i] * tcw[2]);
    for (int i = 0; i < 3; i++) tlc[i] = Rlw[i * 3] * twc[0] + Rlw[i * 3 + 1] * twc[1] + Rlw[i * 3 + 2] * twc[2] + tlw[i];
    const bool bForward = tlc[2] > CurrentFrame.mb && !bMono;
    const bool bBackward = -tlc[2] > CurrentFrame.mb && !bMono;

    const int nq = LastFrame.N;
    std::vector<orbx_proj_query> q(nq);
    std::vector<uint8_t> qdesc((size_t)nq * 32);
    std::vector<MapPoint*> owner(nq, nullptr);
    for (int i = 0; i < nq; i++) {
        q[i].valid = 0;
        MapPoint* pMP = LastFrame.mvpMapPoints[i];
        if (!pMP || LastFrame.mvbOutlier[i]) continue;
        const float* Xw = pMP->GetWorldPosPtr();
        float Xc[3];
        for (int r = 0; r < 3; r++) Xc[r] = Rcw[r * 3] * Xw[0] + Rcw[r * 3 + 1] * Xw[1] + Rcw[r * 3 + 2] * Xw[2] + tcw[r];
        const float invzc = 1.0f / Xc[2];
        if (invzc < 0) continue;
        const cv::Point2f uv = CurrentFrame.mpCamera->project(Xc);
        if (uv.x < Frame::mnMinX || uv.x > Frame::mnMaxX || uv.y < Frame::mnMinY || uv.y > Frame::mnMaxY) continue;
        const int nLastOctave = LastFrame.mvKeys[i].octave;
        q[i].u = uv.x; q[i].v = uv.y; q[i].r = th * CurrentFrame.mvScaleFactors[nLastOctave];
        if (bForward) { q[i].minl = nLastOctave; q[i].maxl = -1; }
        else if (bBackward) { q[i].minl = 0; q[i].maxl = nLastOctave; }
        else { q[i].minl = nLastOctave - 1; q[i].maxl = nLastOctave + 1; }
        q[i].ur = uv.x - CurrentFrame.mbf * invzc;
        q[i].angle = LastFrame.mvKeysUn[i].angle;
        q[i].valid = 1;
        std::memcpy(qdesc.data() + (size_t)i * 32, pMP->GetDescriptor().ptr(0), 32);
        owner[i] = pMP;
    }
    const int n2 = CurrentFrame.N;
    std::vector<int32_t> assigned(n2, -1);
    for (int i = 0; i < n2; i++)
        if (CurrentFrame.mvpMapPoints[i] && CurrentFrame.mvpMapPoints[i]->Observations() > 0) assigned[i] = nq;   // occupied (:2045-2047)
    std::vector<uint8_t> t2;
    const float bounds[4] = {Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY};
    int nmatches = 0;
    check(orbx_search_by_projection(context(), 0, q.data(), qdesc.data(), nq,
                                    reinterpret_cast<const orbx_keypoint*>(CurrentFrame.mvKeysUn.data()), rows32(CurrentFrame.mDescriptors, t2),
                                    CurrentFrame.mvuRight.empty() ? nullptr : CurrentFrame.mvuRight.data(), n2, bounds, assigned.data(),
                                    mfNNratio, mbCheckOrientation ? 1 : 0, &nmatches));
    for (int i = 0; i < n2; i++)
        if (assigned[i] >= 0 && assigned[i] < nq) CurrentFrame.mvpMapPoints[i] = owner[assigned[i]];
    return nmatches;
}

// R/src/ORBmatcher.cc:44-214, left-image branch
int ORBmatcher::SearchByProjection(Frame &F, const std::vector<MapPoint*> &vpMapPoints, const float th, const bool bFarPoints, const float thFarPoints)
{
    const bool bFactor = th != 1.0;
    const int nq = (int)vpMapPoints.size();
    std::vector<orbx_proj_query> q(nq);
    std::vector<uint8_t> qdesc((size_t)nq * 32);
    for (int i = 0; i < nq; i++) {
        q[i].valid = 0;
        MapPoint* pMP = vpMapPoints[i];
        if (!pMP->mbTrackInView && !pMP->mbTrackInViewR) continue;
        if (bFarPoints && pMP->mTrackDepth > thFarPoints) continue;
        if (pMP->isBad() || !pMP->mbTrackInView) continue;
        const int nPredictedLevel = pMP->mnTrackScaleLevel;
        float r = RadiusByViewingCos(pMP->mTrackViewCos);
        if (bFactor) r *= th;
        q[i].u = pMP->mTrackProjX; q[i].v = pMP->mTrackProjY; q[i].r = r * F.mvScaleFactors[nPredictedLevel];
        q[i].minl = nPredictedLevel - 1; q[i].maxl = nPredictedLevel;
        q[i].ur = pMP->mTrackProjXR; q[i].angle = 0; q[i].valid = 1;
        std::memcpy(qdesc.data() + (size_t)i * 32, pMP->GetDescriptor().ptr(0), 32);
    }
    const int n2 = F.N;
    std::vector<int32_t> assigned(n2, -1);
    for (int i = 0; i < n2; i++)
        if (F.mvpMapPoints[i] && F.mvpMapPoints[i]->Observations() > 0) assigned[i] = nq;          // occupied (:89-91)
    std::vector<uint8_t> t2;
    const float bounds[4] = {Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY};
    int nmatches = 0;
    check(orbx_search_by_projection(context(), 1, q.data(), qdesc.data(), nq,
                                    reinterpret_cast<const orbx_keypoint*>(F.mvKeysUn.data()), rows32(F.mDescriptors, t2),
                                    F.mvuRight.empty() ? nullptr : F.mvuRight.data(), n2, bounds, assigned.data(),
                                    mfNNratio, mbCheckOrientation ? 1 : 0, &nmatches));
    for (int i = 0; i < n2; i++)
        if (assigned[i] >= 0 && assigned[i] < nq) F.mvpMapPoints[i] = vpMapPoints[assigned[i]];
    return nmatches;
}

// FeatureVector (std::map<NodeId, vector<unsigned>>) -> CSR sorted by node id
static void fv_csr(const DBoW2::FeatureVector& fv, std::vector<int32_t>& nodes, std::vector<int32_t>& start, std::vector<int32_t>& feat)
{
    nodes.clear(); start.assign(1, 0); feat.clear();
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it) {
        nodes.push_back((int32_t)it->first);
        for (size_t j = 0; j < it->second.size(); j++) feat.push_back((int32_t)it->second[j]);
        start.push_back((int32_t)feat.size());
    }
}
static void valid_flags(const std::vector<MapPoint*>& mps, int n, std::vector<uint8_t>& valid)
{
    valid.assign(n, 0);
    for (int i = 0; i < n && i < (int)mps.size(); i++) valid[i] = (mps[i] && !mps[i]->isBad()) ? 1 : 0;
}

// R/src/ORBmatcher.cc:269-471, monocular branch (F.Nleft == -1)
int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame &F, std::vector<MapPoint*> &vpMapPointMatches)
{
    const std::vector<MapPoint*> vpMapPointsKF = pKF->GetMapPointMatches();
    vpMapPointMatches = std::vector<MapPoint*>(F.N, static_cast<MapPoint*>(NULL));
    const int n1 = pKF->mDescriptors.rows, n2 = F.N;
    if (n1 == 0 || n2 == 0) return 0;
    std::vector<int32_t> nd1, st1, ft1, nd2, st2, ft2, m12(n1, -1);
    fv_csr(pKF->mFeatVec, nd1, st1, ft1); fv_csr(F.mFeatVec, nd2, st2, ft2);
    std::vector<uint8_t> v1, t1, t2;
    valid_flags(vpMapPointsKF, n1, v1);
    int nmatches = 0;
    check(orbx_search_by_bow(context(), 0, reinterpret_cast<const orbx_keypoint*>(pKF->mvKeysUn.data()), rows32(pKF->mDescriptors, t1), v1.data(), n1,
                             nd1.data(), st1.data(), ft1.data(), (int)nd1.size(),
                             reinterpret_cast<const orbx_keypoint*>(F.mvKeys.data()), rows32(F.mDescriptors, t2), nullptr, n2,
                             nd2.data(), st2.data(), ft2.data(), (int)nd2.size(), mfNNratio, mbCheckOrientation ? 1 : 0, m12.data(), &nmatches));
    for (int i1 = 0; i1 < n1; i1++)
        if (m12[i1] >= 0) vpMapPointMatches[m12[i1]] = vpMapPointsKF[i1];
    return nmatches;
}

// R/src/ORBmatcher.cc:819-959
int ORBmatcher::SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, std::vector<MapPoint *> &vpMatches12)
{
    const std::vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();
    vpMatches12 = std::vector<MapPoint*>(vpMapPoints1.size(), static_cast<MapPoint*>(NULL));
    const int n1 = pKF1->mDescriptors.rows, n2 = pKF2->mDescriptors.rows;
    if (n1 == 0 || n2 == 0) return 0;
    std::vector<int32_t> nd1, st1, ft1, nd2, st2, ft2, m12(n1, -1);
    fv_csr(pKF1->mFeatVec, nd1, st1, ft1); fv_csr(pKF2->mFeatVec, nd2, st2, ft2);
    std::vector<uint8_t> v1, v2, t1, t2;
    valid_flags(vpMapPoints1, n1, v1); valid_flags(vpMapPoints2, n2, v2);
    int nmatches = 0;
    check(orbx_search_by_bow(context(), 1, reinterpret_cast<const orbx_keypoint*>(pKF1->mvKeysUn.data()), rows32(pKF1->mDescriptors, t1), v1.data(), n1,
                             nd1.data(), st1.data(), ft1.data(), (int)nd1.size(),
                             reinterpret_cast<const orbx_keypoint*>(pKF2->mvKeysUn.data()), rows32(pKF2->mDescriptors, t2), v2.data(), n2,
                             nd2.data(), st2.data(), ft2.data(), (int)nd2.size(), mfNNratio, mbCheckOrientation ? 1 : 0, m12.data(), &nmatches));
    for (int i1 = 0; i1 < n1 && i1 < (int)vpMatches12.size(); i1++)
        if (m12[i1] >= 0) vpMatches12[i1] = vpMapPoints2[m12[i1]];
    return nmatches;
}

// R/src/Frame.cc:785-962 (see ORBmatcher.h)
void ORBmatcher::ComputeStereoMatches(ORBextractor* pLeft, ORBextractor* pRight, float mb, float mbf,
                                      std::vector<float> &mvuRight, std::vector<float> &mvDepth)
{
    const int cap = orbx_extractor_max_keypoints(pLeft->handle());
    mvuRight.assign(cap, -1.0f); mvDepth.assign(cap, -1.0f);
    int n = 0;
    check(orbx_stereo_matches(context(), pLeft->handle(), pRight->handle(), 0, 0, 0, 0, mb, mbf, mvuRight.data(), mvDepth.data(),
                              nullptr, cap, &n));
    mvuRight.resize(n); mvDepth.resize(n);
}

} //namespace ORB_SLAM
