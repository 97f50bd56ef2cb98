// Drop-in bodies of ORB_SLAM3::ORBmatcher over the B200 C ABI (include/orbx.h).  Replaces R/orb_slam3/src/ORBmatcher.cc as a whole:
// all 15 public methods and the 4 protected helpers of R/orb_slam3/include/ORBmatcher.h:35-108.
//
// Split of the work, the same for every search:
//   host   the geometry the reference evaluates per MapPoint (pose algebra on cv::Mat, projection through the camera model,
//          image / depth / viewing-angle tests, PredictScale) with the reference's own expressions, so that window centres,
//          radii and level ranges are the reference's to the bit; the map bookkeeping after the search (AddObservation, Replace);
//   device GetFeaturesInArea + DescriptorDistance + best / second bookkeeping in visit order + rotation histogram
//          (orbx_search_by_projection_opts, orbx_search_for_initialization, orbx_search_by_bow, orbx_search_for_triangulation,
//          orbx_hamming_pairs).
// Only members that exist in the reference's headers are used (mTcw.rowRange(), GetWorldPos(), project(cv::Point3f), ...), so the
// one source builds against the real headers in the reference tree and against the stand-ins of dropin/shim here.
// There is no CPU fallback: a failing device call throws std::runtime_error.
//
// Two-camera frames (KannalaBrandt8 rig, Frame::Nleft != -1): SearchByProjection(Frame&, vector<MapPoint*>&), SearchByProjection(
// Frame&, const Frame&), SearchByBoW(KeyFrame*, Frame&), Fuse(..., bRight) and both SearchForTriangulation overloads follow the
// reference's left / right branches (orbx_search_by_projection_rig, orbx_search_by_bow_rig; R/src/ORBmatcher.cc:144-213, :344-431,
// :1395-1560, :2093-2160); SearchByBoW(KeyFrame*, KeyFrame*) skips the right camera's features as the reference does (:854-876).
// The relocalisation overload SearchByProjection(Frame&, KeyFrame*, ...) has no two-camera branch in the reference either (it
// indexes mvKeysUn with right-camera indices there); it throws for such frames, as do mixed one- / two-camera argument pairs the
// reference does not produce.
#include "ORBmatcher.h"

#include <limits.h>
#include <cmath>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>

#include "BFMatcher.h"
#include "ORBextractor.h"
#include "orbx.h"

using namespace std;

namespace ORB_SLAM3
{

const int ORBmatcher::TH_HIGH = 100;
const int ORBmatcher::TH_LOW = 50;
const int ORBmatcher::HISTO_LENGTH = 30;

namespace
{

static_assert(sizeof(cv::KeyPoint) == sizeof(orbx_keypoint), "cv::KeyPoint must have the 28-byte layout of orbx_keypoint");
static_assert(sizeof(cv::Point2f) == 2 * sizeof(float), "cv::Point2f layout");

[[noreturn]] void fail(const char* what)
{
    throw std::runtime_error(std::string("ORBmatcher (B200): ") + what + ": " + orbx_last_error());
}
[[noreturn]] void unsupported(const char* what)
{
    throw std::runtime_error(std::string("ORBmatcher (B200): ") + what + " is not implemented by the B200 drop-in (two-camera rigs: keep the reference's ORBmatcher for them)");
}

// ---- one matcher context per (calling thread, device): Tracking / LocalMapping / LoopClosing call concurrently ----
struct Context { orbx_matcher* m; int K; int pool; };
thread_local std::map<int, Context> tl_contexts;
thread_local int tl_device = -1;

int current_device()
{
    if (tl_device >= 0) return tl_device;
    return ORBextractor::ThreadDevice();
}

// a context that holds at least `need` keypoints / queries; grown (never shrunk) on demand
Context& context(int need, int min_pool = 0)
{
    const int dev = current_device();
    Context& c = tl_contexts[dev];
    int K = c.m ? c.K : 8192;
    while (K < need) K *= 2;
    if (K > 24000) K = 24000;
    int pool = c.m ? c.pool : (1 << 18);
    if (pool < min_pool) pool = min_pool;
    if (!c.m || K != c.K || pool != c.pool) {
        if (c.m) orbx_matcher_destroy(c.m);
        c.m = NULL;
        orbx_matcher_params p; p.device = dev; p.max_keypoints = K; p.max_batch = 2; p.max_candidates = pool;      // two pairs: the halves of a two-camera frame
        if (orbx_matcher_create(&p, &c.m) != ORBX_OK) fail("orbx_matcher_create");
        c.K = K; c.pool = pool;
    }
    return c;
}

// rows of an N x 32 CV_8U matrix as one contiguous block
const uint8_t* rows32(const cv::Mat& d, std::vector<uint8_t>& tmp)
{
    if (d.rows == 0) return NULL;
    if (d.isContinuous()) return d.ptr(0);
    tmp.resize((size_t)d.rows * 32);
    for (int i = 0; i < d.rows; i++) std::memcpy(tmp.data() + (size_t)i * 32, d.ptr(i), 32);
    return tmp.data();
}

inline const orbx_keypoint* kp_ptr(const std::vector<cv::KeyPoint>& v) { return reinterpret_cast<const orbx_keypoint*>(v.data()); }

// queries of one projection search, gathered on the host
struct Queries {
    std::vector<orbx_proj_query> q;
    std::vector<uint8_t> desc;
    std::vector<int> tag;                         // caller's index of the query (MapPoint / keypoint index)
    void add(float u, float v, float r, int minl, int maxl, float ur, float angle, bool occupies, const cv::Mat& d, int t)
    {
        orbx_proj_query e;
        e.u = u; e.v = v; e.r = r; e.minl = minl; e.maxl = maxl; e.ur = ur; e.angle = angle; e.valid = occupies ? 1 : 3;
        q.push_back(e);
        const size_t o = desc.size();
        desc.resize(o + 32);
        std::memcpy(desc.data() + o, d.ptr(0), 32);
        tag.push_back(t);
    }
    int size() const { return (int)q.size(); }
};

struct Target {                                    // the searched frame / keyframe
    const std::vector<cv::KeyPoint>* keys;
    const cv::Mat* desc;
    int row0 = 0;                                  // first descriptor row of `keys` (NLeft for the right camera of a two-camera keyframe)
    const float* uright;                           // may be NULL
    float bounds[4];
    float qorigin[2];
};

Target target_of(const Frame& F, bool with_uright)
{
    Target t;
    t.keys = &F.mvKeysUn; t.desc = &F.mDescriptors;
    t.uright = (with_uright && (int)F.mvuRight.size() == F.N && F.N > 0) ? F.mvuRight.data() : NULL;
    t.bounds[0] = Frame::mnMinX; t.bounds[1] = Frame::mnMaxX; t.bounds[2] = Frame::mnMinY; t.bounds[3] = Frame::mnMaxY;
    t.qorigin[0] = Frame::mnMinX; t.qorigin[1] = Frame::mnMinY;
    return t;
}

// KeyFrame::GetFeaturesInArea (R/src/KeyFrame.cc:889-934) runs on the grid that Frame::AssignFeaturesToGrid built on the float
// bounds, but offsets the query by KeyFrame::mnMinX / mnMinY, which are ints (R/include/KeyFrame.h:501-504)
Target target_of(const KeyFrame* pKF, bool with_uright)
{
    Target t;
    t.keys = &pKF->mvKeysUn; t.desc = &pKF->mDescriptors;
    t.uright = (with_uright && (int)pKF->mvuRight.size() == pKF->N && pKF->N > 0) ? pKF->mvuRight.data() : NULL;
    t.bounds[0] = Frame::mnMinX; t.bounds[1] = Frame::mnMaxX; t.bounds[2] = Frame::mnMinY; t.bounds[3] = Frame::mnMaxY;
    t.qorigin[0] = (float)pKF->mnMinX; t.qorigin[1] = (float)pKF->mnMinY;
    return t;
}

struct SearchSpec {
    int mode;                    // 0 best only + optional rotation histogram, 1 best / second with the level rule, 3 independent best
    float nnratio; bool check_ori; int max_dist;
    const std::vector<float>* inv_sigma2; double chi2_mono, chi2_stereo;
    SearchSpec(int m, float r, bool o, int d) : mode(m), nnratio(r), check_ori(o), max_dist(d), inv_sigma2(NULL), chi2_mono(0), chi2_stereo(0) {}
};

// One device search.  assigned (modes 0 / 1): in >= 0 = occupied keypoint; out = index of the query that owns it.
// Queries beyond the context's capacity are searched in consecutive chunks with the occupancy carried over (exact for every mode
// without a global rotation histogram; with one, the query count of a frame never reaches the capacity of 24000).
int run_search(const SearchSpec& spec, const Queries& Q, const Target& T, std::vector<int32_t>& assigned,
               std::vector<int32_t>* best_idx, std::vector<int32_t>* best_dist)
{
    const int nq = Q.size(), n2 = (int)T.keys->size();
    if (best_idx) { best_idx->assign(nq, -1); best_dist->assign(nq, 256); }
    if (nq == 0 || n2 == 0) return 0;
    if (n2 > 24000) fail("more than 24000 keypoints in one frame");
    std::vector<uint8_t> tmp;
    const uint8_t* d2 = rows32(*T.desc, tmp) + (size_t)T.row0 * 32;
    orbx_proj_options o;
    std::memset(&o, 0, sizeof(o));
    for (int i = 0; i < 4; i++) o.bounds[i] = T.bounds[i];
    o.query_origin[0] = T.qorigin[0]; o.query_origin[1] = T.qorigin[1];
    o.nnratio = spec.nnratio; o.check_ori = spec.check_ori ? 1 : 0; o.max_dist = spec.max_dist;
    if (spec.chi2_mono > 0) {
        o.nlevels = (int)spec.inv_sigma2->size() < ORBX_MAX_LEVELS ? (int)spec.inv_sigma2->size() : ORBX_MAX_LEVELS;
        for (int l = 0; l < o.nlevels; l++) o.inv_level_sigma2[l] = (*spec.inv_sigma2)[l];
        o.chi2_mono = spec.chi2_mono; o.chi2_stereo = spec.chi2_stereo;
    }
    const int cap = 24000;
    if (nq > cap && spec.mode == 0 && spec.check_ori) fail("more than 24000 queries in a search with a rotation histogram");
    int total = 0;
    // occupancy carried between chunks: a claim by a non-occupying query must stay invisible to the next chunk
    std::vector<int32_t> owner;
    if (nq > cap && spec.mode != 3) owner = assigned;
    for (int q0 = 0; q0 < nq; q0 += cap) {
        const int cnt = nq - q0 < cap ? nq - q0 : cap;
        int min_pool = 0;
        for (int attempt = 0;; attempt++) {
            Context& c = context(cnt > n2 ? cnt : n2, min_pool);
            std::vector<int32_t> a = assigned;
            int nm = 0;
            const int rc = orbx_search_by_projection_opts(c.m, spec.mode, Q.q.data() + q0, Q.desc.data() + (size_t)q0 * 32, cnt, kp_ptr(*T.keys), d2,
                                                          T.uright, n2, &o, spec.mode == 3 ? NULL : a.data(),
                                                          best_idx ? best_idx->data() + q0 : NULL, best_dist ? best_dist->data() + q0 : NULL, &nm);
            if (rc == ORBX_E_CAPACITY && attempt < 6) { min_pool = c.pool * 4; continue; }       // wide windows: a larger candidate pool
            if (rc != ORBX_OK) fail("orbx_search_by_projection_opts");
            total += nm;
            if (spec.mode != 3) {
                if (nq <= cap) assigned.swap(a);
                else
                    for (int i = 0; i < n2; i++)
                        if (a[i] != assigned[i]) {                 // claimed in this chunk by query a[i] (chunk-local index)
                            owner[i] = q0 + a[i];
                            if (!(Q.q[q0 + a[i]].valid & 2)) assigned[i] = q0 + a[i];
                        }
            }
            break;
        }
    }
    if (nq > cap && spec.mode != 3) assigned.swap(owner);
    return total;
}

// ---- two-camera frames (KannalaBrandt8 rig, Frame::Nleft != -1) ----
// The frame's keypoints are mvKeys (left, [0, Nleft)) followed by mvKeysRight ([Nleft, N)); descriptor rows follow the same order
// (Frame.cc:1079-1082).  Every map point has a left and a right query; the device call interleaves them over one occupancy table
// as the reference's loop does (orbx_search_by_projection_rig).
struct RigQueries {
    std::vector<orbx_proj_query> ql, qr;
    std::vector<uint8_t> desc;
    std::vector<int> tag;
    int add(const cv::Mat& d, bool occupies, int t)
    {
        orbx_proj_query e; std::memset(&e, 0, sizeof(e));
        e.valid = occupies ? 0 : 2;                               // bit 0 is set by left() / right()
        ql.push_back(e); qr.push_back(e);
        const size_t o = desc.size();
        desc.resize(o + 32);
        std::memcpy(desc.data() + o, d.ptr(0), 32);
        tag.push_back(t);
        return (int)ql.size() - 1;
    }
    static void fill(orbx_proj_query& e, float u, float v, float r, int minl, int maxl, float angle)
    {
        e.u = u; e.v = v; e.r = r; e.minl = minl; e.maxl = maxl; e.ur = 0.f; e.angle = angle; e.valid |= 1;
    }
    void left(int i, float u, float v, float r, int minl, int maxl, float angle) { fill(ql[i], u, v, r, minl, maxl, angle); }
    void right(int i, float u, float v, float r, int minl, int maxl, float angle) { fill(qr[i], u, v, r, minl, maxl, angle); }
    int size() const { return (int)ql.size(); }
};

int run_rig_search(int mode, float nnratio, bool check_ori, int max_dist, const RigQueries& Q, const Frame& F, bool partners,
                   std::vector<int32_t>& assigned)
{
    const int nq = Q.size(), n2 = F.N;
    if (nq == 0 || n2 == 0) return 0;
    if (n2 > 24000 || nq > 24000) fail("more than 24000 keypoints / map points in a two-camera search");
    std::vector<cv::KeyPoint> keys(F.mvKeys.begin(), F.mvKeys.begin() + F.Nleft);
    keys.insert(keys.end(), F.mvKeysRight.begin(), F.mvKeysRight.begin() + F.Nright);
    std::vector<uint8_t> tmp;
    const uint8_t* d2 = rows32(F.mDescriptors, tmp);
    orbx_proj_options o;
    std::memset(&o, 0, sizeof(o));
    o.bounds[0] = Frame::mnMinX; o.bounds[1] = Frame::mnMaxX; o.bounds[2] = Frame::mnMinY; o.bounds[3] = Frame::mnMaxY;
    o.query_origin[0] = Frame::mnMinX; o.query_origin[1] = Frame::mnMinY;
    o.nnratio = nnratio; o.check_ori = check_ori ? 1 : 0; o.max_dist = max_dist;
    std::vector<int32_t> l2r, r2l;
    if (partners) { l2r.assign(F.mvLeftToRightMatch.begin(), F.mvLeftToRightMatch.end()); r2l.assign(F.mvRightToLeftMatch.begin(), F.mvRightToLeftMatch.end()); }
    int min_pool = 0, nm = 0;
    for (int attempt = 0;; attempt++) {
        Context& c = context(nq > n2 ? nq : n2, min_pool);
        std::vector<int32_t> a = assigned;
        const int rc = orbx_search_by_projection_rig(c.m, mode, Q.ql.data(), Q.qr.data(), Q.desc.data(), nq, kp_ptr(keys), d2, F.Nleft, F.Nright,
                                                     partners ? l2r.data() : NULL, partners ? r2l.data() : NULL, &o, a.data(), &nm);
        if (rc == ORBX_E_CAPACITY && attempt < 6) { min_pool = c.pool * 4; continue; }
        if (rc != ORBX_OK) fail("orbx_search_by_projection_rig");
        assigned.swap(a);
        break;
    }
    return nm;
}

// pose split of a similarity transformation as the Sim3 overloads do it (R/src/ORBmatcher.cc:484-488, :598-602, :1616-1620)
struct Sim3Pose {
    cv::Mat Rcw, tcw, Ow;
    explicit Sim3Pose(const cv::Mat& Scw)
    {
        cv::Mat sRcw = Scw.rowRange(0,3).colRange(0,3);
        const float scw = sqrt(sRcw.row(0).dot(sRcw.row(0)));
        Rcw = sRcw/scw;
        tcw = Scw.rowRange(0,3).col(3)/scw;
        Ow = -Rcw.t()*tcw;
    }
};

// rotation-histogram bin of a pair of keypoint angles (:772-778 and its copies)
inline int rotation_bin(float a1, float a2)
{
    const float factor = 1.0f/ORBmatcher::HISTO_LENGTH;
    float rot = a1-a2;
    if(rot<0.0)
        rot+=360.0f;
    int bin = round(rot*factor);
    if(bin==ORBmatcher::HISTO_LENGTH)
        bin=0;
    return bin;
}

// limit >= 0: features with an index >= limit are left out (the right camera's features of a two-camera keyframe, which
// SearchByBoW(KeyFrame*, KeyFrame*) skips, R/src/ORBmatcher.cc:854-856, :874-876)
void feature_vector_csr(const DBoW2::FeatureVector& fv, std::vector<int32_t>& nodes, std::vector<int32_t>& start, std::vector<int32_t>& feat, int limit = -1)
{
    nodes.clear(); start.assign(1, 0); feat.clear();
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it) {
        nodes.push_back((int32_t)it->first);
        for (size_t j = 0; j < it->second.size(); j++)
            if (limit < 0 || (int)it->second[j] < limit) feat.push_back((int32_t)it->second[j]);
        start.push_back((int32_t)feat.size());
    }
}

}  // namespace

void ORBmatcher::SetDevice(int device) { tl_device = device; }

ORBmatcher::ORBmatcher(float nnratio, bool checkOri): mfNNratio(nnratio), mbCheckOrientation(checkOri)
{
}

// R/src/ORBmatcher.cc:2358-2374.  A single pair is a scalar accessor that host code outside the front-end calls in its own loops
// (Frame.cc:860, MapPoint.cc:496); every search below computes its distances on the device.
int ORBmatcher::DescriptorDistance(const cv::Mat &a, const cv::Mat &b)
{
    const uint32_t *pa = a.ptr<uint32_t>(), *pb = b.ptr<uint32_t>();
    int dist = 0;
    for (int i = 0; i < 8; i++) dist += __builtin_popcount(pa[i] ^ pb[i]);
    return dist;
}

float ORBmatcher::RadiusByViewingCos(const float &viewCos)              // :216-222
{
    return viewCos>0.998 ? 2.5 : 4.0;
}

// :225-245 / :247-267: distance of kp2 to the epipolar line of kp1 under F12, against the chi-square bound of the keypoint's level
static inline bool epipolar_distance(const cv::KeyPoint &kp1, const cv::KeyPoint &kp2, const cv::Mat &F12, float& dsqr)
{
    const float a = kp1.pt.x*F12.at<float>(0,0)+kp1.pt.y*F12.at<float>(1,0)+F12.at<float>(2,0);
    const float b = kp1.pt.x*F12.at<float>(0,1)+kp1.pt.y*F12.at<float>(1,1)+F12.at<float>(2,1);
    const float c = kp1.pt.x*F12.at<float>(0,2)+kp1.pt.y*F12.at<float>(1,2)+F12.at<float>(2,2);
    const float num = a*kp2.pt.x+b*kp2.pt.y+c;
    const float den = a*a+b*b;
    if(den==0)
        return false;
    dsqr = num*num/den;
    return true;
}

bool ORBmatcher::CheckDistEpipolarLine(const cv::KeyPoint &kp1,const cv::KeyPoint &kp2,const cv::Mat &F12,const KeyFrame* pKF2, const bool b1)
{
    float dsqr;
    if (!epipolar_distance(kp1, kp2, F12, dsqr)) return false;
    return b1 ? dsqr<6.63*pKF2->mvLevelSigma2[kp2.octave] : dsqr<3.84*pKF2->mvLevelSigma2[kp2.octave];
}

bool ORBmatcher::CheckDistEpipolarLine2(const cv::KeyPoint &kp1, const cv::KeyPoint &kp2, const cv::Mat &F12, const KeyFrame *pKF2, const float unc)
{
    float dsqr;
    if (!epipolar_distance(kp1, kp2, F12, dsqr)) return false;
    return unc==1.f ? dsqr<3.84*pKF2->mvLevelSigma2[kp2.octave] : dsqr<3.84*pKF2->mvLevelSigma2[kp2.octave]*unc;
}

// :2312-2353: the three most populated bins; the second / third are dropped below 10 % of the first; ties go to the lower bin
void ORBmatcher::ComputeThreeMaxima(vector<int>* histo, const int L, int &ind1, int &ind2, int &ind3)
{
    int best[3] = {0, 0, 0};
    int* ind[3] = {&ind1, &ind2, &ind3};
    for (int i = 0; i < L; i++) {
        const int s = histo[i].size();
        for (int r = 0; r < 3; r++)
            if (s > best[r]) {
                for (int t = 2; t > r; t--) { best[t] = best[t-1]; *ind[t] = *ind[t-1]; }
                best[r] = s; *ind[r] = i;
                break;
            }
    }
    if (best[1] < 0.1f*(float)best[0]) { ind2 = -1; ind3 = -1; }
    else if (best[2] < 0.1f*(float)best[0]) ind3 = -1;
}

// ---------------------------------------------------------------------------------------------------------------------------
// R/src/ORBmatcher.cc:44-214  Tracking::SearchLocalPoints: local-map points (already tested by Frame::isInFrustum) against F
int ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, const float th, const bool bFarPoints, const float thFarPoints)
{
    const bool bFactor = th!=1.0;
    if (F.Nleft != -1) {
        // :144-212: every point searches the left image (mbTrackInView) and then the right one (mbTrackInViewR); a match also takes
        // the keypoint's stereo partner (mvLeftToRightMatch / mvRightToLeftMatch)
        RigQueries Q;
        for (size_t iMP = 0; iMP < vpMapPoints.size(); iMP++) {
            MapPoint* pMP = vpMapPoints[iMP];
            if (!pMP->mbTrackInView && !pMP->mbTrackInViewR) continue;
            if (bFarPoints && pMP->mTrackDepth>thFarPoints) continue;
            if (pMP->isBad()) continue;
            const int qi = Q.add(pMP->GetDescriptor(), pMP->Observations()>0, (int)iMP);
            if (pMP->mbTrackInView) {
                const int &nPredictedLevel = pMP->mnTrackScaleLevel;
                float r = RadiusByViewingCos(pMP->mTrackViewCos);
                if(bFactor)
                    r*=th;
                Q.left(qi, pMP->mTrackProjX, pMP->mTrackProjY, r*F.mvScaleFactors[nPredictedLevel], nPredictedLevel-1, nPredictedLevel, 0.f);
            }
            if (pMP->mbTrackInViewR) {
                const int &nPredictedLevel = pMP->mnTrackScaleLevelR;
                if (nPredictedLevel != -1) {
                    float r = RadiusByViewingCos(pMP->mTrackViewCosR);                 // no th factor on this side (:148)
                    Q.right(qi, pMP->mTrackProjXR, pMP->mTrackProjYR, r*F.mvScaleFactors[nPredictedLevel], nPredictedLevel-1, nPredictedLevel, 0.f);
                }
            }
        }
        std::vector<int32_t> assigned(F.N, -1);
        const int occupied = Q.size();
        for (int i = 0; i < F.N; i++)
            if (F.mvpMapPoints[i] && F.mvpMapPoints[i]->Observations()>0) assigned[i] = occupied;
        const int nmatches = run_rig_search(1, mfNNratio, false, TH_HIGH, Q, F, true, assigned);
        for (int i = 0; i < F.N; i++)
            if (assigned[i] >= 0 && assigned[i] < occupied) F.mvpMapPoints[i] = vpMapPoints[Q.tag[assigned[i]]];
        return nmatches;
    }
    Queries Q;
    for (size_t iMP = 0; iMP < vpMapPoints.size(); iMP++) {
        MapPoint* pMP = vpMapPoints[iMP];
        if (!pMP->mbTrackInView && !pMP->mbTrackInViewR) continue;
        if (bFarPoints && pMP->mTrackDepth>thFarPoints) continue;
        if (pMP->isBad()) continue;
        if (!pMP->mbTrackInView) continue;
        const int &nPredictedLevel = pMP->mnTrackScaleLevel;
        float r = RadiusByViewingCos(pMP->mTrackViewCos);       // the window depends on the viewing direction (:66-70)
        if(bFactor)
            r*=th;
        Q.add(pMP->mTrackProjX, pMP->mTrackProjY, r*F.mvScaleFactors[nPredictedLevel], nPredictedLevel-1, nPredictedLevel,
              pMP->mTrackProjXR, 0.f, pMP->Observations()>0, pMP->GetDescriptor(), (int)iMP);
    }
    // a keypoint is skipped only when it holds a MapPoint WITH observations (:89-91)
    std::vector<int32_t> assigned(F.N, -1);
    const int occupied = Q.size();
    for (int i = 0; i < F.N; i++)
        if (F.mvpMapPoints[i] && F.mvpMapPoints[i]->Observations()>0) assigned[i] = occupied;
    SearchSpec spec(1, mfNNratio, false, TH_HIGH);
    const int nmatches = run_search(spec, Q, target_of(F, true), assigned, NULL, NULL);
    for (int i = 0; i < F.N; i++)
        if (assigned[i] >= 0 && assigned[i] < occupied) F.mvpMapPoints[i] = vpMapPoints[Q.tag[assigned[i]]];
    return nmatches;
}

// Writes the outcome of a mode-0 search into the frame's MapPoint slots: the owner where a query kept its claim, NULL where a
// claim was cleared by the rotation check (the reference NULLs the slot whatever it held before, :2176-2180 / :2298-2302).
template <class GetPoint>
static void write_claims(std::vector<MapPoint*>& slots, const std::vector<int32_t>& assigned, int nqueries, GetPoint point_of_query)
{
    for (size_t i = 0; i < slots.size(); i++) {
        if (assigned[i] >= 0 && assigned[i] < nqueries) slots[i] = point_of_query(assigned[i]);
        else if (assigned[i] == -2) slots[i] = static_cast<MapPoint*>(NULL);
    }
}

// R/src/ORBmatcher.cc:1970-2186  Tracking::TrackWithMotionModel: the last frame's points projected with the predicted pose
int ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono)
{
    if (CurrentFrame.Nleft == -1 && LastFrame.Nleft != -1) unsupported("SearchByProjection(Frame&, const Frame&) from a two-camera frame into a one-camera frame");
    const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0,3).colRange(0,3);
    const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0,3).col(3);
    const cv::Mat twc = -Rcw.t()*tcw;
    const cv::Mat Rlw = LastFrame.mTcw.rowRange(0,3).colRange(0,3);
    const cv::Mat tlw = LastFrame.mTcw.rowRange(0,3).col(3);
    const cv::Mat tlc = Rlw*twc+tlw;
    // moving forward / backward along the optical axis shifts the octave range that is searched (:1990-1991, :2026-2031)
    const bool bForward = tlc.at<float>(2)>CurrentFrame.mb && !bMono;
    const bool bBackward = -tlc.at<float>(2)>CurrentFrame.mb && !bMono;

    if (CurrentFrame.Nleft != -1) {
        // :2093-2160: after the left image, the point is projected into the right camera (mTrl) and the best free keypoint of
        // mvKeysRight within the same radius / octave range takes it; both claims enter the one rotation histogram
        RigQueries Q;
        for (int i = 0; i < LastFrame.N; i++) {
            MapPoint* pMP = LastFrame.mvpMapPoints[i];
            if (!pMP || LastFrame.mvbOutlier[i]) continue;
            cv::Mat x3Dw = pMP->GetWorldPos();
            cv::Mat x3Dc = Rcw*x3Dw+tcw;
            const float invzc = 1.0/x3Dc.at<float>(2);
            if(invzc<0)
                continue;
            cv::Point2f uv = CurrentFrame.mpCamera->project(x3Dc);
            if(uv.x<CurrentFrame.mnMinX || uv.x>CurrentFrame.mnMaxX)
                continue;
            if(uv.y<CurrentFrame.mnMinY || uv.y>CurrentFrame.mnMaxY)
                continue;
            const cv::KeyPoint& kpLF = (LastFrame.Nleft == -1) ? LastFrame.mvKeysUn[i]
                                                                : (i < LastFrame.Nleft) ? LastFrame.mvKeys[i] : LastFrame.mvKeysRight[i - LastFrame.Nleft];
            const int nLastOctave = (LastFrame.Nleft == -1 || i < LastFrame.Nleft) ? LastFrame.mvKeys[i].octave
                                                                                   : LastFrame.mvKeysRight[i - LastFrame.Nleft].octave;
            const float radius = th*CurrentFrame.mvScaleFactors[nLastOctave];
            int minl, maxl;
            if (bForward) { minl = nLastOctave; maxl = -1; }
            else if (bBackward) { minl = 0; maxl = nLastOctave; }
            else { minl = nLastOctave-1; maxl = nLastOctave+1; }
            const int qi = Q.add(pMP->GetDescriptor(), pMP->Observations()>0, i);
            Q.left(qi, uv.x, uv.y, radius, minl, maxl, kpLF.angle);
            cv::Mat x3Dr = CurrentFrame.mTrl.colRange(0,3).rowRange(0,3) * x3Dc + CurrentFrame.mTrl.col(3);
            cv::Point2f uvr = CurrentFrame.mpCamera->project(x3Dr);
            Q.right(qi, uvr.x, uvr.y, radius, minl, maxl, kpLF.angle);
        }
        std::vector<int32_t> assigned(CurrentFrame.N, -1);
        const int occupied = Q.size();
        for (int i = 0; i < CurrentFrame.N; i++)
            if (CurrentFrame.mvpMapPoints[i] && CurrentFrame.mvpMapPoints[i]->Observations()>0) assigned[i] = occupied;
        const int nmatches = run_rig_search(0, mfNNratio, mbCheckOrientation, TH_HIGH, Q, CurrentFrame, false, assigned);
        write_claims(CurrentFrame.mvpMapPoints, assigned, occupied, [&](int q) { return LastFrame.mvpMapPoints[Q.tag[q]]; });
        return nmatches;
    }
    Queries Q;
    for (int i = 0; i < LastFrame.N; i++) {
        MapPoint* pMP = LastFrame.mvpMapPoints[i];
        if (!pMP || LastFrame.mvbOutlier[i]) continue;
        cv::Mat x3Dw = pMP->GetWorldPos();
        cv::Mat x3Dc = Rcw*x3Dw+tcw;
        const float invzc = 1.0/x3Dc.at<float>(2);
        if(invzc<0)
            continue;
        cv::Point2f uv = CurrentFrame.mpCamera->project(x3Dc);
        if(uv.x<CurrentFrame.mnMinX || uv.x>CurrentFrame.mnMaxX)
            continue;
        if(uv.y<CurrentFrame.mnMinY || uv.y>CurrentFrame.mnMaxY)
            continue;
        const int nLastOctave = LastFrame.mvKeys[i].octave;
        const float radius = th*CurrentFrame.mvScaleFactors[nLastOctave];
        int minl, maxl;
        if (bForward) { minl = nLastOctave; maxl = -1; }
        else if (bBackward) { minl = 0; maxl = nLastOctave; }
        else { minl = nLastOctave-1; maxl = nLastOctave+1; }
        const float ur = uv.x - CurrentFrame.mbf*invzc;          // predicted right coordinate for the stereo gate (:2051-2055)
        Q.add(uv.x, uv.y, radius, minl, maxl, ur, LastFrame.mvKeysUn[i].angle, pMP->Observations()>0, pMP->GetDescriptor(), i);
    }
    std::vector<int32_t> assigned(CurrentFrame.N, -1);
    const int occupied = Q.size();
    for (int i = 0; i < CurrentFrame.N; i++)
        if (CurrentFrame.mvpMapPoints[i] && CurrentFrame.mvpMapPoints[i]->Observations()>0) assigned[i] = occupied;     // :2045-2047
    SearchSpec spec(0, mfNNratio, mbCheckOrientation, TH_HIGH);
    const int nmatches = run_search(spec, Q, target_of(CurrentFrame, true), assigned, NULL, NULL);
    write_claims(CurrentFrame.mvpMapPoints, assigned, occupied, [&](int q) { return LastFrame.mvpMapPoints[Q.tag[q]]; });
    return nmatches;
}

// R/src/ORBmatcher.cc:2188-2310  Tracking::Relocalization: the points of a candidate keyframe projected into the frame
int ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const set<MapPoint*> &sAlreadyFound, const float th , const int ORBdist)
{
    if (CurrentFrame.Nleft != -1) unsupported("SearchByProjection(Frame&, KeyFrame*, ...) on a two-camera frame");
    const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0,3).colRange(0,3);
    const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0,3).col(3);
    const cv::Mat Ow = -Rcw.t()*tcw;
    const vector<MapPoint*> vpMPs = pKF->GetMapPointMatches();

    Queries Q;
    for (size_t i = 0; i < vpMPs.size(); i++) {
        MapPoint* pMP = vpMPs[i];
        if (!pMP || pMP->isBad() || sAlreadyFound.count(pMP)) continue;
        cv::Mat x3Dw = pMP->GetWorldPos();
        cv::Mat x3Dc = Rcw*x3Dw+tcw;
        const cv::Point2f uv = CurrentFrame.mpCamera->project(x3Dc);
        if(uv.x<CurrentFrame.mnMinX || uv.x>CurrentFrame.mnMaxX)
            continue;
        if(uv.y<CurrentFrame.mnMinY || uv.y>CurrentFrame.mnMaxY)
            continue;
        // the depth must lie inside the scale-invariance range of the point (:2222-2232)
        cv::Mat PO = x3Dw-Ow;
        float dist3D = cv::norm(PO);
        const float maxDistance = pMP->GetMaxDistanceInvariance();
        const float minDistance = pMP->GetMinDistanceInvariance();
        if(dist3D<minDistance || dist3D>maxDistance)
            continue;
        int nPredictedLevel = pMP->PredictScale(dist3D,&CurrentFrame);
        const float radius = th*CurrentFrame.mvScaleFactors[nPredictedLevel];
        Q.add(uv.x, uv.y, radius, nPredictedLevel-1, nPredictedLevel+1, 0.f, pKF->mvKeysUn[i].angle, true, pMP->GetDescriptor(), (int)i);
    }
    std::vector<int32_t> assigned(CurrentFrame.N, -1);
    const int occupied = Q.size();
    for (int i = 0; i < CurrentFrame.N; i++)
        if (CurrentFrame.mvpMapPoints[i]) assigned[i] = occupied;                      // any MapPoint blocks the slot here (:2253-2254)
    SearchSpec spec(0, mfNNratio, mbCheckOrientation, ORBdist);
    const int nmatches = run_search(spec, Q, target_of(CurrentFrame, false), assigned, NULL, NULL);
    write_claims(CurrentFrame.mvpMapPoints, assigned, occupied, [&](int q) { return vpMPs[Q.tag[q]]; });
    return nmatches;
}

// The candidate geometry shared by the searches that project map points into a KEYFRAME with a given pose: depth, image,
// scale-invariance range, viewing angle (< 60 degrees), predicted level (:506-545, :620-657, :1451-1499, :1642-1686).
// How (u, v) is formed differs between the overloads and is left to the caller.
namespace {
struct KeyFrameProjection {
    float radius; int level;
};
inline bool distance_and_normal_ok(MapPoint* pMP, const cv::Mat& p3Dw, const cv::Mat& Ow, float& dist)
{
    const float maxDistance = pMP->GetMaxDistanceInvariance();
    const float minDistance = pMP->GetMinDistanceInvariance();
    cv::Mat PO = p3Dw-Ow;
    dist = cv::norm(PO);
    if(dist<minDistance || dist>maxDistance)
        return false;
    cv::Mat Pn = pMP->GetNormal();
    if(PO.dot(Pn)<0.5*dist)
        return false;
    return true;
}
}

// R/src/ORBmatcher.cc:473-586  LoopClosing: candidate points under a similarity transformation against the keyframe's free features
int ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*> &vpPoints,
                                   vector<MapPoint*> &vpMatched, int th, float ratioHamming)
{
    const Sim3Pose P(Scw);
    set<MapPoint*> spAlreadyFound(vpMatched.begin(), vpMatched.end());
    spAlreadyFound.erase(static_cast<MapPoint*>(NULL));

    Queries Q;
    for (int iMP = 0, iendMP = vpPoints.size(); iMP < iendMP; iMP++) {
        MapPoint* pMP = vpPoints[iMP];
        if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
        cv::Mat p3Dw = pMP->GetWorldPos();
        cv::Mat p3Dc = P.Rcw*p3Dw+P.tcw;
        if(p3Dc.at<float>(2)<0.0)
            continue;
        const float x = p3Dc.at<float>(0);
        const float y = p3Dc.at<float>(1);
        const float z = p3Dc.at<float>(2);
        const cv::Point2f uv = pKF->mpCamera->project(cv::Point3f(x,y,z));
        if(!pKF->IsInImage(uv.x,uv.y))
            continue;
        float dist;
        if (!distance_and_normal_ok(pMP, p3Dw, P.Ow, dist)) continue;
        int nPredictedLevel = pMP->PredictScale(dist,pKF);
        const float radius = th*pKF->mvScaleFactors[nPredictedLevel];
        Q.add(uv.x, uv.y, radius, nPredictedLevel-1, nPredictedLevel, 0.f, 0.f, true, pMP->GetDescriptor(), iMP);
    }
    std::vector<int32_t> assigned(pKF->N, -1);
    const int occupied = Q.size();
    for (int i = 0; i < pKF->N && i < (int)vpMatched.size(); i++)
        if (vpMatched[i]) assigned[i] = occupied;                                     // :560-561
    // "bestDist <= TH_LOW * ratioHamming" with an int on the left (:577)
    SearchSpec spec(0, mfNNratio, false, (int)std::floor(TH_LOW*ratioHamming));
    const int nmatches = run_search(spec, Q, target_of(pKF, false), assigned, NULL, NULL);
    for (int i = 0; i < pKF->N; i++)
        if (assigned[i] >= 0 && assigned[i] < occupied) vpMatched[i] = vpPoints[Q.tag[assigned[i]]];
    return nmatches;
}

// R/src/ORBmatcher.cc:588-700  place recognition / merging: the same, remembering which keyframe every matched point came from
int ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*> &vpPoints, const std::vector<KeyFrame*> &vpPointsKFs,
                                   std::vector<MapPoint*> &vpMatched, std::vector<KeyFrame*> &vpMatchedKF, int th, float ratioHamming)
{
    const float &fx = pKF->fx;
    const float &fy = pKF->fy;
    const float &cx = pKF->cx;
    const float &cy = pKF->cy;
    const Sim3Pose P(Scw);
    set<MapPoint*> spAlreadyFound(vpMatched.begin(), vpMatched.end());
    spAlreadyFound.erase(static_cast<MapPoint*>(NULL));

    Queries Q;
    for (int iMP = 0, iendMP = vpPoints.size(); iMP < iendMP; iMP++) {
        MapPoint* pMP = vpPoints[iMP];
        if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
        cv::Mat p3Dw = pMP->GetWorldPos();
        cv::Mat p3Dc = P.Rcw*p3Dw+P.tcw;
        if(p3Dc.at<float>(2)<0.0)
            continue;
        // this overload projects with the keyframe's intrinsics directly (:629-633)
        const float invz = 1/p3Dc.at<float>(2);
        const float x = p3Dc.at<float>(0)*invz;
        const float y = p3Dc.at<float>(1)*invz;
        const float u = fx*x+cx;
        const float v = fy*y+cy;
        if(!pKF->IsInImage(u,v))
            continue;
        float dist;
        if (!distance_and_normal_ok(pMP, p3Dw, P.Ow, dist)) continue;
        int nPredictedLevel = pMP->PredictScale(dist,pKF);
        const float radius = th*pKF->mvScaleFactors[nPredictedLevel];
        Q.add(u, v, radius, nPredictedLevel-1, nPredictedLevel, 0.f, 0.f, true, pMP->GetDescriptor(), iMP);
    }
    std::vector<int32_t> assigned(pKF->N, -1);
    const int occupied = Q.size();
    for (int i = 0; i < pKF->N && i < (int)vpMatched.size(); i++)
        if (vpMatched[i]) assigned[i] = occupied;
    SearchSpec spec(0, mfNNratio, false, (int)std::floor(TH_LOW*ratioHamming));
    const int nmatches = run_search(spec, Q, target_of(pKF, false), assigned, NULL, NULL);
    for (int i = 0; i < pKF->N; i++)
        if (assigned[i] >= 0 && assigned[i] < occupied) {
            vpMatched[i] = vpPoints[Q.tag[assigned[i]]];
            vpMatchedKF[i] = vpPointsKFs[Q.tag[assigned[i]]];
        }
    return nmatches;
}

// ---------------------------------------------------------------------------------------------------------------------------
// R/src/ORBmatcher.cc:269-471  relocalisation / loop detection: keyframe MapPoints against the features of a frame, node by node
int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame &F, vector<MapPoint*> &vpMapPointMatches)
{
    const vector<MapPoint*> vpMapPointsKF = pKF->GetMapPointMatches();
    vpMapPointMatches = vector<MapPoint*>(F.N,static_cast<MapPoint*>(NULL));
    const int n1 = pKF->N, n2 = F.N;
    if (n1 == 0 || n2 == 0) return 0;
    if (F.Nleft != -1) {
        // :344-431: a best / second pair per camera of the frame; the keyframe's keypoint is taken from the camera it belongs to
        // (:379-382), the frame's from mvKeys / mvKeysRight (:386-389, :416-419)
        if (pKF->NLeft == -1 && pKF->mpCamera2) unsupported("SearchByBoW: a one-camera keyframe with a second camera model");
        std::vector<cv::KeyPoint> k1v(n1), k2v(F.mvKeys.begin(), F.mvKeys.begin() + F.Nleft);
        for (int i = 0; i < n1; i++)
            k1v[i] = (!pKF->mpCamera2) ? pKF->mvKeysUn[i] : (i >= pKF->NLeft) ? pKF->mvKeysRight[i - pKF->NLeft] : pKF->mvKeys[i];
        k2v.insert(k2v.end(), F.mvKeysRight.begin(), F.mvKeysRight.begin() + F.Nright);
        std::vector<uint8_t> valid1(n1, 0);
        for (int i = 0; i < n1 && i < (int)vpMapPointsKF.size(); i++)
            valid1[i] = (vpMapPointsKF[i] && !vpMapPointsKF[i]->isBad()) ? 1 : 0;
        std::vector<int32_t> n1v, s1v, f1v, n2v, s2v, f2v, ml(n1, -1), mr(n1, -1);
        feature_vector_csr(pKF->mFeatVec, n1v, s1v, f1v);
        feature_vector_csr(F.mFeatVec, n2v, s2v, f2v);
        std::vector<uint8_t> t1, t2;
        int nmatches = 0;
        Context& c = context(n1 > n2 ? n1 : n2);
        if (orbx_search_by_bow_rig(c.m, kp_ptr(k1v), rows32(pKF->mDescriptors, t1), valid1.data(), n1, n1v.data(), s1v.data(), f1v.data(),
                                   (int)n1v.size(), kp_ptr(k2v), rows32(F.mDescriptors, t2), n2, F.Nleft, n2v.data(), s2v.data(), f2v.data(),
                                   (int)n2v.size(), mfNNratio, mbCheckOrientation ? 1 : 0, ml.data(), mr.data(), &nmatches) != ORBX_OK)
            fail("orbx_search_by_bow_rig");
        for (int i1 = 0; i1 < n1; i1++) {
            if (ml[i1] >= 0) vpMapPointMatches[ml[i1]] = vpMapPointsKF[i1];
            if (mr[i1] >= 0) vpMapPointMatches[mr[i1]] = vpMapPointsKF[i1];
        }
        return nmatches;
    }
    if (pKF->NLeft != -1) unsupported("SearchByBoW(KeyFrame*, Frame&) from a two-camera keyframe into a one-camera frame");
    std::vector<uint8_t> valid1(n1, 0);
    for (int i = 0; i < n1 && i < (int)vpMapPointsKF.size(); i++)
        valid1[i] = (vpMapPointsKF[i] && !vpMapPointsKF[i]->isBad()) ? 1 : 0;          // :301-307
    std::vector<int32_t> n1v, s1v, f1v, n2v, s2v, f2v, m12(n1, -1);
    feature_vector_csr(pKF->mFeatVec, n1v, s1v, f1v);
    feature_vector_csr(F.mFeatVec, n2v, s2v, f2v);
    std::vector<uint8_t> t1, t2;
    int nmatches = 0;
    Context& c = context(n1 > n2 ? n1 : n2);
    if (orbx_search_by_bow(c.m, 0, kp_ptr(pKF->mvKeysUn), rows32(pKF->mDescriptors, t1), valid1.data(), n1, n1v.data(), s1v.data(), f1v.data(),
                           (int)n1v.size(), kp_ptr(F.mvKeys), rows32(F.mDescriptors, t2), NULL, n2, n2v.data(), s2v.data(), f2v.data(),
                           (int)n2v.size(), mfNNratio, mbCheckOrientation ? 1 : 0, m12.data(), &nmatches) != ORBX_OK)
        fail("orbx_search_by_bow");
    for (int i1 = 0; i1 < n1; i1++)
        if (m12[i1] >= 0) vpMapPointMatches[m12[i1]] = vpMapPointsKF[i1];              // stored per FRAME feature (:368)
    return nmatches;
}

// R/src/ORBmatcher.cc:819-959  loop closing: the MapPoints of two keyframes, node by node
int ORBmatcher::SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, vector<MapPoint *> &vpMatches12)
{
    const vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches();
    const vector<MapPoint*> vpMapPoints2 = pKF2->GetMapPointMatches();
    vpMatches12 = vector<MapPoint*>(vpMapPoints1.size(),static_cast<MapPoint*>(NULL));
    const int n1 = (int)pKF1->mvKeysUn.size(), n2 = (int)pKF2->mvKeysUn.size();
    if (n1 == 0 || n2 == 0) return 0;
    std::vector<uint8_t> valid1(n1, 0), valid2(n2, 0);
    for (int i = 0; i < n1 && i < (int)vpMapPoints1.size(); i++) valid1[i] = (vpMapPoints1[i] && !vpMapPoints1[i]->isBad()) ? 1 : 0;
    for (int i = 0; i < n2 && i < (int)vpMapPoints2.size(); i++) valid2[i] = (vpMapPoints2[i] && !vpMapPoints2[i]->isBad()) ? 1 : 0;
    std::vector<int32_t> n1v, s1v, f1v, n2v, s2v, f2v, m12(n1, -1);
    feature_vector_csr(pKF1->mFeatVec, n1v, s1v, f1v, pKF1->NLeft != -1 ? n1 : -1);
    feature_vector_csr(pKF2->mFeatVec, n2v, s2v, f2v, pKF2->NLeft != -1 ? n2 : -1);
    std::vector<uint8_t> t1, t2;
    int nmatches = 0;
    Context& c = context(n1 > n2 ? n1 : n2);
    if (orbx_search_by_bow(c.m, 1, kp_ptr(pKF1->mvKeysUn), rows32(pKF1->mDescriptors, t1), valid1.data(), n1, n1v.data(), s1v.data(), f1v.data(),
                           (int)n1v.size(), kp_ptr(pKF2->mvKeysUn), rows32(pKF2->mDescriptors, t2), valid2.data(), n2, n2v.data(), s2v.data(),
                           f2v.data(), (int)n2v.size(), mfNNratio, mbCheckOrientation ? 1 : 0, m12.data(), &nmatches) != ORBX_OK)
        fail("orbx_search_by_bow");
    for (int i1 = 0; i1 < n1 && i1 < (int)vpMatches12.size(); i1++)
        if (m12[i1] >= 0) vpMatches12[i1] = vpMapPoints2[m12[i1]];
    return nmatches;
}

// R/src/ORBmatcher.cc:702-817  monocular map initialisation
int ORBmatcher::SearchForInitialization(Frame &F1, Frame &F2, vector<cv::Point2f> &vbPrevMatched, vector<int> &vnMatches12, int windowSize)
{
    const int n1 = (int)F1.mvKeysUn.size(), n2 = (int)F2.mvKeysUn.size();
    vnMatches12 = vector<int>(n1,-1);
    if (n1 == 0) return 0;
    const float bounds[4] = {Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY};
    std::vector<uint8_t> t1, t2;
    int nmatches = 0;
    int min_pool = 0;
    for (int attempt = 0;; attempt++) {
        Context& c = context(n1 > n2 ? n1 : n2, min_pool);
        const int rc = orbx_search_for_initialization(c.m, kp_ptr(F1.mvKeysUn), rows32(F1.mDescriptors, t1), n1, kp_ptr(F2.mvKeysUn),
                                                      rows32(F2.mDescriptors, t2), n2, bounds, reinterpret_cast<float*>(vbPrevMatched.data()),
                                                      vnMatches12.data(), windowSize, mfNNratio, mbCheckOrientation ? 1 : 0, &nmatches);
        if (rc == ORBX_E_CAPACITY && attempt < 6) { min_pool = c.pool * 4; continue; }
        if (rc != ORBX_OK) fail("orbx_search_for_initialization");
        break;
    }
    return nmatches;
}

// ---------------------------------------------------------------------------------------------------------------------------
// SearchForTriangulation, both overloads (R/src/ORBmatcher.cc:961-1202, :1204-1394).  Per vocabulary node the two keyframes
// share, every feature of keyframe 1 without a MapPoint looks for the feature of keyframe 2 (without a MapPoint) with the least
// Hamming distance <= TH_LOW that also passes a geometric predicate; a later candidate with the same distance replaces an
// earlier one.  This fork never sets vbMatched2, so queries do not interact.  The Hamming distances of ALL candidate pairs come
// from one device call (orbx_hamming_pairs); the predicate goes through the camera's virtual interface, so any camera model and
// the two-camera branches work.
namespace {
struct PairList { std::vector<int32_t> i1, i2; std::vector<int32_t> first; std::vector<int32_t> query; };

template <class Accept1, class Accept2>
void gather_node_pairs(KeyFrame* pKF1, KeyFrame* pKF2, Accept1 accept1, Accept2 accept2, PairList& L)
{
    const DBoW2::FeatureVector &vFeatVec1 = pKF1->mFeatVec;
    const DBoW2::FeatureVector &vFeatVec2 = pKF2->mFeatVec;
    DBoW2::FeatureVector::const_iterator f1it = vFeatVec1.begin(), f2it = vFeatVec2.begin();
    const DBoW2::FeatureVector::const_iterator f1end = vFeatVec1.end(), f2end = vFeatVec2.end();
    while (f1it != f1end && f2it != f2end) {
        if (f1it->first == f2it->first) {
            for (size_t a = 0; a < f1it->second.size(); a++) {
                const int idx1 = f1it->second[a];
                if (!accept1(idx1)) continue;
                L.query.push_back(idx1);
                L.first.push_back((int32_t)L.i1.size());
                for (size_t b = 0; b < f2it->second.size(); b++) {
                    const int idx2 = f2it->second[b];
                    if (!accept2(idx2)) continue;
                    L.i1.push_back(idx1); L.i2.push_back(idx2);
                }
            }
            f1it++; f2it++;
        }
        else if (f1it->first < f2it->first) f1it = vFeatVec1.lower_bound(f2it->first);
        else f2it = vFeatVec2.lower_bound(f1it->first);
    }
    L.first.push_back((int32_t)L.i1.size());
}

void pair_distances(KeyFrame* pKF1, KeyFrame* pKF2, const PairList& L, std::vector<int32_t>& dist)
{
    const size_t n = L.i1.size();
    dist.assign(n, 256);
    if (n == 0) return;
    std::vector<uint8_t> a(n * 32), b(n * 32);
    for (size_t k = 0; k < n; k++) {
        std::memcpy(a.data() + k * 32, pKF1->mDescriptors.ptr(L.i1[k]), 32);
        std::memcpy(b.data() + k * 32, pKF2->mDescriptors.ptr(L.i2[k]), 32);
    }
    Context& c = context(1);
    if (orbx_hamming_pairs(c.m, a.data(), b.data(), (int)n, dist.data()) != ORBX_OK) fail("orbx_hamming_pairs");
}

inline const cv::KeyPoint& keypoint_of(const KeyFrame* pKF, int idx)
{
    return (pKF->NLeft == -1) ? pKF->mvKeysUn[idx] : (idx < pKF->NLeft) ? pKF->mvKeys[idx] : pKF->mvKeysRight[idx - pKF->NLeft];
}
}

int ORBmatcher::SearchForTriangulation(KeyFrame *pKF1, KeyFrame *pKF2, cv::Mat F12,
                                       vector<pair<size_t, size_t> > &vMatchedPairs, const bool bOnlyStereo, const bool bCoarse)
{
    (void)F12;          // unused by the reference as well: the epipolar test recomputes it from the relative pose (:1140)
    // epipole of camera 1 in image 2 (:968-973)
    cv::Mat Cw = pKF1->GetCameraCenter();
    cv::Mat R2w = pKF2->GetRotation();
    cv::Mat t2w = pKF2->GetTranslation();
    cv::Mat C2 = R2w*Cw+t2w;
    cv::Point2f ep = pKF2->mpCamera->project(C2);
    cv::Mat R1w = pKF1->GetRotation();
    cv::Mat t1w = pKF1->GetTranslation();
    cv::Mat R12, t12;
    cv::Mat Rll,Rlr,Rrl,Rrr;
    cv::Mat tll,tlr,trl,trr;
    GeometricCamera* pCamera1 = pKF1->mpCamera, *pCamera2 = pKF2->mpCamera;
    const bool two_cameras = pKF1->mpCamera2 && pKF2->mpCamera2;
    if (!pKF1->mpCamera2 && !pKF2->mpCamera2) {
        R12 = R1w*R2w.t();
        t12 = -R1w*R2w.t()*t2w+t1w;
    } else {
        Rll = pKF1->GetRotation() * pKF2->GetRotation().t();
        Rlr = pKF1->GetRotation() * pKF2->GetRightRotation().t();
        Rrl = pKF1->GetRightRotation() * pKF2->GetRotation().t();
        Rrr = pKF1->GetRightRotation() * pKF2->GetRightRotation().t();
        tll = pKF1->GetRotation() * (-pKF2->GetRotation().t() * pKF2->GetTranslation()) + pKF1->GetTranslation();
        tlr = pKF1->GetRotation() * (-pKF2->GetRightRotation().t() * pKF2->GetRightTranslation()) + pKF1->GetTranslation();
        trl = pKF1->GetRightRotation() * (-pKF2->GetRotation().t() * pKF2->GetTranslation()) + pKF1->GetRightTranslation();
        trr = pKF1->GetRightRotation() * (-pKF2->GetRightRotation().t() * pKF2->GetRightTranslation()) + pKF1->GetRightTranslation();
    }

    auto stereo1 = [&](int i) { return !pKF1->mpCamera2 && pKF1->mvuRight[i]>=0; };
    auto stereo2 = [&](int i) { return !pKF2->mpCamera2 && pKF2->mvuRight[i]>=0; };
    PairList L;
    gather_node_pairs(pKF1, pKF2,
                      [&](int i1) { return !pKF1->GetMapPoint(i1) && (!bOnlyStereo || stereo1(i1)); },
                      [&](int i2) { return !pKF2->GetMapPoint(i2) && (!bOnlyStereo || stereo2(i2)); }, L);
    std::vector<int32_t> dist;
    pair_distances(pKF1, pKF2, L, dist);

    int nmatches=0;
    vector<int> vMatches12(pKF1->N,-1);
    vector<int> rotHist[HISTO_LENGTH];
    for (size_t qi = 0; qi < L.query.size(); qi++) {
        const int idx1 = L.query[qi];
        const cv::KeyPoint &kp1 = keypoint_of(pKF1, idx1);
        const bool bStereo1 = stereo1(idx1);
        const bool bRight1 = !(pKF1->NLeft == -1 || idx1 < pKF1->NLeft);
        int bestDist = TH_LOW;
        int bestIdx2 = -1;
        for (int k = L.first[qi]; k < L.first[qi + 1]; k++) {
            const int idx2 = L.i2[k];
            if(dist[k]>TH_LOW || dist[k]>bestDist)
                continue;
            const cv::KeyPoint &kp2 = keypoint_of(pKF2, idx2);
            const bool bStereo2 = stereo2(idx2);
            const bool bRight2 = !(pKF2->NLeft == -1 || idx2 < pKF2->NLeft);
            if(!bStereo1 && !bStereo2 && !pKF1->mpCamera2)
            {
                // monocular pairs too close to the epipole are skipped (:1101-1109)
                const float distex = ep.x-kp2.pt.x;
                const float distey = ep.y-kp2.pt.y;
                if(distex*distex+distey*distey<100*pKF2->mvScaleFactors[kp2.octave])
                    continue;
            }
            if (two_cameras) {
                if (bRight1 && bRight2) { R12 = Rrr; t12 = trr; pCamera1 = pKF1->mpCamera2; pCamera2 = pKF2->mpCamera2; }
                else if (bRight1 && !bRight2) { R12 = Rrl; t12 = trl; pCamera1 = pKF1->mpCamera2; pCamera2 = pKF2->mpCamera; }
                else if (!bRight1 && bRight2) { R12 = Rlr; t12 = tlr; pCamera1 = pKF1->mpCamera; pCamera2 = pKF2->mpCamera2; }
                else { R12 = Rll; t12 = tll; pCamera1 = pKF1->mpCamera; pCamera2 = pKF2->mpCamera; }
            }
            if(pCamera1->epipolarConstrain(pCamera2,kp1,kp2,R12,t12,pKF1->mvLevelSigma2[kp1.octave],pKF2->mvLevelSigma2[kp2.octave])||bCoarse)
            {
                bestIdx2 = idx2;
                bestDist = dist[k];
            }
        }
        if (bestIdx2 >= 0) {
            vMatches12[idx1]=bestIdx2;
            nmatches++;
            if (mbCheckOrientation) rotHist[rotation_bin(kp1.angle, keypoint_of(pKF2, bestIdx2).angle)].push_back(idx1);
        }
    }
    if (mbCheckOrientation) {
        int ind1=-1, ind2=-1, ind3=-1;
        ComputeThreeMaxima(rotHist,HISTO_LENGTH,ind1,ind2,ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0; j < rotHist[i].size(); j++) { vMatches12[rotHist[i][j]]=-1; nmatches--; }
        }
    }
    vMatchedPairs.clear();
    vMatchedPairs.reserve(nmatches);
    for (size_t i = 0; i < vMatches12.size(); i++)
        if (vMatches12[i] >= 0) vMatchedPairs.push_back(make_pair(i,vMatches12[i]));
    return nmatches;
}

int ORBmatcher::SearchForTriangulation(KeyFrame *pKF1, KeyFrame *pKF2, cv::Mat F12,
                                       vector<pair<size_t, size_t> > &vMatchedPairs, const bool bOnlyStereo, vector<cv::Mat> &vMatchedPoints)
{
    (void)F12; (void)bOnlyStereo;       // neither is read by the reference body (:1204-1394)
    GeometricCamera* pCamera1 = pKF1->mpCamera, *pCamera2 = pKF2->mpCamera;
    cv::Mat Tcw1,Tcw2;
    PairList L;
    gather_node_pairs(pKF1, pKF2, [&](int i1) { return !pKF1->GetMapPoint(i1); }, [&](int i2) { return !pKF2->GetMapPoint(i2); }, L);
    std::vector<int32_t> dist;
    pair_distances(pKF1, pKF2, L, dist);

    int nmatches=0;
    vector<int> vMatches12(pKF1->N,-1);
    vector<cv::Mat> vMatchesPoints12(pKF1 -> N);
    vector<int> rotHist[HISTO_LENGTH];
    for (size_t qi = 0; qi < L.query.size(); qi++) {
        const int idx1 = L.query[qi];
        const cv::KeyPoint &kp1 = keypoint_of(pKF1, idx1);
        const bool bRight1 = !(pKF1->NLeft == -1 || idx1 < pKF1->NLeft);
        int bestDist = TH_LOW;
        int bestIdx2 = -1;
        cv::Mat bestPoint;
        for (int k = L.first[qi]; k < L.first[qi + 1]; k++) {
            const int idx2 = L.i2[k];
            if(dist[k]>TH_LOW || dist[k]>bestDist)
                continue;
            const cv::KeyPoint &kp2 = keypoint_of(pKF2, idx2);
            const bool bRight2 = !(pKF2->NLeft == -1 || idx2 < pKF2->NLeft);
            if (bRight1) { Tcw1 = pKF1->GetRightPose(); pCamera1 = pKF1->mpCamera2; } else { Tcw1 = pKF1->GetPose(); pCamera1 = pKF1->mpCamera; }
            if (bRight2) { Tcw2 = pKF2->GetRightPose(); pCamera2 = pKF2->mpCamera2; } else { Tcw2 = pKF2->GetPose(); pCamera2 = pKF2->mpCamera; }
            cv::Mat x3D;
            if(pCamera1->matchAndtriangulate(kp1,kp2,pCamera2,Tcw1,Tcw2,pKF1->mvLevelSigma2[kp1.octave],pKF2->mvLevelSigma2[kp2.octave],x3D)){
                bestIdx2 = idx2;
                bestDist = dist[k];
                bestPoint = x3D;
            }
        }
        if (bestIdx2 >= 0) {
            vMatches12[idx1]=bestIdx2;
            vMatchesPoints12[idx1] = bestPoint;
            nmatches++;
            if (mbCheckOrientation) rotHist[rotation_bin(kp1.angle, keypoint_of(pKF2, bestIdx2).angle)].push_back(idx1);
        }
    }
    if (mbCheckOrientation) {
        int ind1=-1, ind2=-1, ind3=-1;
        ComputeThreeMaxima(rotHist,HISTO_LENGTH,ind1,ind2,ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0; j < rotHist[i].size(); j++) { vMatches12[rotHist[i][j]]=-1; nmatches--; }
        }
    }
    vMatchedPairs.clear();
    vMatchedPairs.reserve(nmatches);
    for (size_t i = 0; i < vMatches12.size(); i++) {
        if (vMatches12[i] < 0) continue;
        vMatchedPairs.push_back(make_pair(i,vMatches12[i]));
        vMatchedPoints.push_back(vMatchesPoints12[i]);
    }
    return nmatches;
}

// ---------------------------------------------------------------------------------------------------------------------------
// R/src/ORBmatcher.cc:1395-1605  LocalMapping::SearchInNeighbors: project points into a keyframe, fuse duplicates.
// The best feature of every point is independent of the others (occupied features are candidates too), so one device search
// returns all of them; the map update that follows is sequential and stays here, re-checking what earlier fusions changed.
int ORBmatcher::Fuse(KeyFrame *pKF, const vector<MapPoint *> &vpMapPoints, const float th, const bool bRight)
{
    // :1400-1413: the right camera of a two-camera keyframe has its own pose, model, keypoints (mvKeysRight) and grid (mGridRight)
    cv::Mat Rcw = bRight ? pKF->GetRightRotation() : pKF->GetRotation();
    cv::Mat tcw = bRight ? pKF->GetRightTranslation() : pKF->GetTranslation();
    cv::Mat Ow = bRight ? pKF->GetRightCameraCenter() : pKF->GetCameraCenter();
    GeometricCamera* pCamera = bRight ? pKF->mpCamera2 : pKF->mpCamera;
    const float &bf = pKF->mbf;
    Target target = target_of(pKF, true);
    int first = 0;                                                  // keyframe index of the searched camera's first keypoint (:1555)
    if (pKF->NLeft != -1) {
        // :1520-1522; mvuRight of a two-camera frame is all -1 (Frame.cc:1123): every candidate takes the 2-dof gate
        target.keys = bRight ? &pKF->mvKeysRight : &pKF->mvKeys;
        target.uright = NULL;
        if (bRight) { target.row0 = pKF->NLeft; first = pKF->NLeft; }
    } else if (bRight) unsupported("Fuse(bRight) on a one-camera keyframe");

    Queries Q;
    const int nMPs = vpMapPoints.size();
    for (int i = 0; i < nMPs; i++) {
        MapPoint* pMP = vpMapPoints[i];
        if (!pMP || pMP->isBad()) continue;
        // IsInKeyFrame is re-evaluated below, in sequence; a point that is in the keyframe NOW stays there, so it can be dropped here
        if (pMP->IsInKeyFrame(pKF)) continue;
        cv::Mat p3Dw = pMP->GetWorldPos();
        cv::Mat p3Dc = Rcw*p3Dw + tcw;
        if(p3Dc.at<float>(2)<0.0f)
            continue;
        const float invz = 1/p3Dc.at<float>(2);
        const float x = p3Dc.at<float>(0);
        const float y = p3Dc.at<float>(1);
        const float z = p3Dc.at<float>(2);
        const cv::Point2f uv = pCamera->project(cv::Point3f(x,y,z));
        if(!pKF->IsInImage(uv.x,uv.y))
            continue;
        const float ur = uv.x-bf*invz;
        float dist3D;
        if (!distance_and_normal_ok(pMP, p3Dw, Ow, dist3D)) continue;
        int nPredictedLevel = pMP->PredictScale(dist3D,pKF);
        const float radius = th*pKF->mvScaleFactors[nPredictedLevel];
        Q.add(uv.x, uv.y, radius, nPredictedLevel-1, nPredictedLevel, ur, 0.f, true, pMP->GetDescriptor(), i);
    }
    // reprojection gates per candidate: 3 dof (7.8) for features with a right coordinate, 2 dof (5.99) otherwise (:1525-1552)
    SearchSpec spec(3, mfNNratio, false, 256);
    spec.inv_sigma2 = &pKF->mvInvLevelSigma2; spec.chi2_mono = 5.99; spec.chi2_stereo = 7.8;
    std::vector<int32_t> none, bestIdx, bestDist;
    run_search(spec, Q, target, none, &bestIdx, &bestDist);

    int nFused=0;
    for (int q = 0; q < Q.size(); q++) {
        MapPoint* pMP = vpMapPoints[Q.tag[q]];
        if (pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;          // an earlier fusion of this call may have changed either
        if (bestIdx[q] < 0 || bestDist[q] > TH_LOW) continue;
        bestIdx[q] += first;
        MapPoint* pMPinKF = pKF->GetMapPoint(bestIdx[q]);
        if (pMPinKF) {
            if (!pMPinKF->isBad()) {
                if(pMPinKF->Observations()>pMP->Observations())
                    pMP->Replace(pMPinKF);
                else
                    pMPinKF->Replace(pMP);
            }
        } else {
            pMP->AddObservation(pKF,bestIdx[q]);
            pKF->AddMapPoint(pMP,bestIdx[q]);
        }
        nFused++;
    }
    return nFused;
}

// R/src/ORBmatcher.cc:1607-1742  LoopClosing::SearchAndFuse: the same under a similarity transformation; duplicates are reported
int ORBmatcher::Fuse(KeyFrame *pKF, cv::Mat Scw, const vector<MapPoint *> &vpPoints, float th, vector<MapPoint *> &vpReplacePoint)
{
    const Sim3Pose P(Scw);
    const set<MapPoint*> spAlreadyFound = pKF->GetMapPoints();

    Queries Q;
    const int nPoints = vpPoints.size();
    for (int iMP = 0; iMP < nPoints; iMP++) {
        MapPoint* pMP = vpPoints[iMP];
        if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
        cv::Mat p3Dw = pMP->GetWorldPos();
        cv::Mat p3Dc = P.Rcw*p3Dw+P.tcw;
        if(p3Dc.at<float>(2)<0.0f)
            continue;
        const float x = p3Dc.at<float>(0);
        const float y = p3Dc.at<float>(1);
        const float z = p3Dc.at<float>(2);
        const cv::Point2f uv = pKF->mpCamera->project(cv::Point3f(x,y,z));
        if(!pKF->IsInImage(uv.x,uv.y))
            continue;
        float dist3D;
        if (!distance_and_normal_ok(pMP, p3Dw, P.Ow, dist3D)) continue;
        const int nPredictedLevel = pMP->PredictScale(dist3D,pKF);
        const float radius = th*pKF->mvScaleFactors[nPredictedLevel];
        Q.add(uv.x, uv.y, radius, nPredictedLevel-1, nPredictedLevel, 0.f, 0.f, true, pMP->GetDescriptor(), iMP);
    }
    SearchSpec spec(3, mfNNratio, false, 256);
    std::vector<int32_t> none, bestIdx, bestDist;
    run_search(spec, Q, target_of(pKF, false), none, &bestIdx, &bestDist);

    int nFused=0;
    for (int q = 0; q < Q.size(); q++) {
        const int iMP = Q.tag[q];
        MapPoint* pMP = vpPoints[iMP];
        if (bestIdx[q] < 0 || bestDist[q] > 100) continue;               // this fork accepts up to 100 here (:1721-1722)
        MapPoint* pMPinKF = pKF->GetMapPoint(bestIdx[q]);
        if (pMPinKF) {
            if(!pMPinKF->isBad())
                vpReplacePoint[iMP] = pMPinKF;
        } else {
            pMP->AddObservation(pKF,bestIdx[q]);
            pKF->AddMapPoint(pMP,bestIdx[q]);
        }
        nFused++;
    }
    return nFused;
}

// R/src/ORBmatcher.cc:1744-1968  loop closing: mutual consistency of the two projections under the Sim3 [s12 R12 | t12]
int ORBmatcher::SearchBySim3(KeyFrame *pKF1, KeyFrame *pKF2, vector<MapPoint*> &vpMatches12,
                             const float &s12, const cv::Mat &R12, const cv::Mat &t12, const float th)
{
    cv::Mat R1w = pKF1->GetRotation();
    cv::Mat t1w = pKF1->GetTranslation();
    cv::Mat R2w = pKF2->GetRotation();
    cv::Mat t2w = pKF2->GetTranslation();
    cv::Mat sR12 = s12*R12;
    cv::Mat sR21 = (1.0/s12)*R12.t();
    cv::Mat t21 = -sR21*t12;

    const vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches();
    const int N1 = vpMapPoints1.size();
    const vector<MapPoint*> vpMapPoints2 = pKF2->GetMapPointMatches();
    const int N2 = vpMapPoints2.size();
    vector<bool> vbAlreadyMatched1(N1,false);
    vector<bool> vbAlreadyMatched2(N2,false);
    for (int i = 0; i < N1; i++) {
        MapPoint* pMP = vpMatches12[i];
        if (!pMP) continue;
        vbAlreadyMatched1[i]=true;
        int idx2 = get<0>(pMP->GetIndexInKeyFrame(pKF2));
        if(idx2>=0 && idx2<N2)
            vbAlreadyMatched2[idx2]=true;
    }

    // one direction: the points of `from` (world -> own camera -> other camera) into the image of `to`
    auto project_all = [&](const vector<MapPoint*>& points, const vector<bool>& done, const cv::Mat& Rw, const cv::Mat& tw,
                           const cv::Mat& sR, const cv::Mat& t, KeyFrame* to, vector<int>& match) {
        const float &fx = to->fx;
        const float &fy = to->fy;
        const float &cx = to->cx;
        const float &cy = to->cy;
        Queries Q;
        for (int i = 0; i < (int)points.size(); i++) {
            MapPoint* pMP = points[i];
            if (!pMP || done[i] || pMP->isBad()) continue;
            cv::Mat p3Dw = pMP->GetWorldPos();
            cv::Mat p3Dca = Rw*p3Dw + tw;
            cv::Mat p3Dcb = sR*p3Dca + t;
            if(p3Dcb.at<float>(2)<0.0)
                continue;
            const float invz = 1.0/p3Dcb.at<float>(2);
            const float x = p3Dcb.at<float>(0)*invz;
            const float y = p3Dcb.at<float>(1)*invz;
            const float u = fx*x+cx;
            const float v = fy*y+cy;
            if(!to->IsInImage(u,v))
                continue;
            const float maxDistance = pMP->GetMaxDistanceInvariance();
            const float minDistance = pMP->GetMinDistanceInvariance();
            const float dist3D = cv::norm(p3Dcb);
            if(dist3D<minDistance || dist3D>maxDistance )
                continue;
            const int nPredictedLevel = pMP->PredictScale(dist3D,to);
            const float radius = th*to->mvScaleFactors[nPredictedLevel];
            Q.add(u, v, radius, nPredictedLevel-1, nPredictedLevel, 0.f, 0.f, true, pMP->GetDescriptor(), i);
        }
        SearchSpec spec(3, mfNNratio, false, 256);
        std::vector<int32_t> none, bestIdx, bestDist;
        run_search(spec, Q, target_of(to, false), none, &bestIdx, &bestDist);
        for (int q = 0; q < Q.size(); q++)
            if (bestIdx[q] >= 0 && bestDist[q] <= TH_HIGH) match[Q.tag[q]] = bestIdx[q];
    };
    vector<int> vnMatch1(N1,-1);
    vector<int> vnMatch2(N2,-1);
    project_all(vpMapPoints1, vbAlreadyMatched1, R1w, t1w, sR21, t21, pKF2, vnMatch1);
    project_all(vpMapPoints2, vbAlreadyMatched2, R2w, t2w, sR12, t12, pKF1, vnMatch2);

    // a match is kept when both directions agree (:1947-1963)
    int nFound = 0;
    for (int i1 = 0; i1 < N1; i1++) {
        const int idx2 = vnMatch1[i1];
        if (idx2 >= 0 && vnMatch2[idx2] == i1) {
            vpMatches12[i1] = vpMapPoints2[idx2];
            nFound++;
        }
    }
    return nFound;
}

// ---------------------------------------------------------------------------------------------------------------------------
// Addition: Frame::ComputeStereoMatches (R/src/Frame.cc:785-962) on the device-resident results and pyramids of the two extractors
void ORBmatcher::ComputeStereoMatches(ORBextractor* pLeft, ORBextractor* pRight, float mb, float mbf,
                                      std::vector<float> &mvuRight, std::vector<float> &mvDepth)
{
    if (!pLeft || !pRight || !pLeft->handle() || !pRight->handle()) throw std::runtime_error("ORBmatcher (B200): ComputeStereoMatches before the extractors ran");
    if (pLeft->device() != pRight->device()) throw std::runtime_error("ORBmatcher (B200): the two extractors of a stereo frame must share a device");
    const int cl = orbx_extractor_max_keypoints(pLeft->handle()), cr = orbx_extractor_max_keypoints(pRight->handle());
    const int cap = cl > cr ? cl : cr;
    const int saved = tl_device;
    tl_device = pLeft->device();                       // the matcher context lives where the extractors' buffers are
    Context& c = context(cap);
    tl_device = saved;
    mvuRight.assign(cap, -1.0f); mvDepth.assign(cap, -1.0f);
    int n = 0;
    if (orbx_stereo_matches(c.m, pLeft->handle(), pRight->handle(), 0, 0, 0, 0, mb, mbf, mvuRight.data(), mvDepth.data(), NULL, cap, &n) != ORBX_OK)
        fail("orbx_stereo_matches");
    mvuRight.resize(n); mvDepth.resize(n);
}

// ---------------------------------------------------------------------------------------------------------------------------
// Addition: the cv::BFMatcher member of Frame (R/include/Frame.h:288, R/src/Frame.cc:26) as a class over orbx_bf_knn2; it is
// what Frame::ComputeStereoFishEyeMatches calls (R/src/Frame.cc:1130).  See dropin/BFMatcher.h.
BFMatcherB200::BFMatcherB200(int normType, bool crossCheck) : norm_(normType)
{
    if (crossCheck) throw std::runtime_error("BFMatcherB200: crossCheck is not what the reference uses and is not implemented");
}

void BFMatcherB200::knnMatch(cv::InputArray queryDescriptors, cv::InputArray trainDescriptors, std::vector<std::vector<cv::DMatch> >& matches, int k) const
{
    const cv::Mat q = queryDescriptors.getMat(), t = trainDescriptors.getMat();
    if (norm_ != cv::NORM_HAMMING || k != 2 || (q.rows > 0 && q.cols != 32) || (t.rows > 0 && t.cols != 32))
        throw std::runtime_error("BFMatcherB200: only NORM_HAMMING, k = 2 and 32-byte descriptor rows (what Frame.cc:1130 asks for)");
    matches.assign(q.rows, std::vector<cv::DMatch>());
    if (q.rows == 0 || t.rows == 0) return;
    std::vector<uint8_t> tq, tt;
    const uint8_t* pq = rows32(q, tq);
    const uint8_t* pt = rows32(t, tt);
    std::vector<int32_t> idx((size_t)q.rows * 2), dist((size_t)q.rows * 2);
    Context& c = context(1);
    if (orbx_bf_knn2(c.m, pq, q.rows, pt, t.rows, idx.data(), dist.data()) != ORBX_OK) fail("orbx_bf_knn2");
    for (int i = 0; i < q.rows; i++)
        for (int j = 0; j < 2; j++)
            if (idx[2 * i + j] >= 0) {
                cv::DMatch dm(i, idx[2 * i + j], (float)dist[2 * i + j]);
                dm.imgIdx = 0;                       // one train image, as cv::BFMatcher reports it
                matches[i].push_back(dm);
            }
}

} //namespace ORB_SLAM
