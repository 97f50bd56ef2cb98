// scenario.cc - TEST INFRASTRUCTURE.  One program, built three ways (tests/scenario/Makefile), that drives EVERY public method of
// ORB_SLAM3::ORBextractor / ORBmatcher / ORBVocabulary through the reference's own Frame / KeyFrame / MapPoint code on a small
// synthetic world, and dumps every result bit for bit:
//   scenario_ref         the reference's ORBextractor.cc, ORBmatcher.cc, DBoW2 (oracle/_ref objects)              -> the truth
//   scenario_dropin_cpu  dropin/*.cc over the oracle-backed C-ABI stand-in (orbx_on_oracle.cc)                   -> host logic, no GPU
//   scenario_dropin_gpu  dropin/*.cc over multi_orbslam3_b200/liborbx_b200.so                                    -> the product path
// tests/test_dropin_scenario.py demands identical dumps.  The source uses nothing but the reference's public interface, so the
// three builds differ only in which ORBextractor / ORBmatcher / ORBVocabulary they link.
//
//   scenario <frames.raw> <W> <H> <nframes> <vocabulary.txt> <out.txt>
// frames: nframes mono stream frames (frame t+1 = frame t shifted by (3, 2) px), then one stereo pair (left, right).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include "orbslam_world.h"
#ifdef SCENARIO_DROPIN
#include "BFMatcher.h"
#endif

using namespace ORB_SLAM3;
using std::vector;

void scenario_before_extract();      // hooks_ref.cc / hooks_dropin.cc
void scenario_after_extract();

namespace {

FILE* g_out = NULL;
unsigned bits(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
void dump_ints(const char* name, const vector<int>& v)
{
    fprintf(g_out, "%s %zu:", name, v.size());
    for (size_t i = 0; i < v.size(); i++) fprintf(g_out, " %d", v[i]);
    fprintf(g_out, "\n");
}
void dump_floats(const char* name, const vector<float>& v)
{
    fprintf(g_out, "%s %zu:", name, v.size());
    for (size_t i = 0; i < v.size(); i++) fprintf(g_out, " %08x", bits(v[i]));
    fprintf(g_out, "\n");
}
vector<int> ids_of(const vector<MapPoint*>& v)
{
    vector<int> r(v.size());
    for (size_t i = 0; i < v.size(); i++) r[i] = v[i] ? (int)v[i]->mnId : -1;
    return r;
}
void dump_points(const char* name, const vector<MapPoint*>& v) { dump_ints(name, ids_of(v)); }

struct World {
    int W, H;
    float fx, fy, cx, cy, bf;
    Pinhole* camera;
    cv::Mat K, dist;
    ORBVocabulary voc;
    Map map;
    vector<std::unique_ptr<MapPoint> > points;
};

cv::Mat pose(float angle_y, float tx, float ty, float tz)
{
    cv::Mat T = cv::Mat::eye(4, 4, CV_32F);
    const float c = std::cos(angle_y), s = std::sin(angle_y);
    T.at<float>(0, 0) = c; T.at<float>(0, 2) = s; T.at<float>(2, 0) = -s; T.at<float>(2, 2) = c;
    T.at<float>(0, 3) = tx; T.at<float>(1, 3) = ty; T.at<float>(2, 3) = tz;
    return T;
}

// what the Frame constructors do around the extractor call (R/src/Frame.cc:240-330), minus IMU and image-bound estimation
void finish_frame(World& w, Frame& F, ORBextractor* ex, bool synth_depth)
{
    F.N = F.mvKeys.size();
    F.mnScaleLevels = ex->GetLevels();
    F.mfScaleFactor = ex->GetScaleFactor();
    F.mfLogScaleFactor = log(F.mfScaleFactor);
    F.mvScaleFactors = ex->GetScaleFactors();
    F.mvInvScaleFactors = ex->GetInverseScaleFactors();
    F.mvLevelSigma2 = ex->GetScaleSigmaSquares();
    F.mvInvLevelSigma2 = ex->GetInverseScaleSigmaSquares();
    F.mK = w.K.clone(); F.mDistCoef = w.dist.clone();
    F.mbf = w.bf; F.mb = w.bf / w.fx; F.mThDepth = 35.f * F.mb;
    F.mpCamera = w.camera; F.mpCamera2 = NULL;
    F.UndistortKeyPoints();
    F.mvuRight.assign(F.N, -1.f); F.mvDepth.assign(F.N, -1.f);
    if (synth_depth)                 // an RGB-D like frame: two thirds of the features have a depth / right coordinate
        for (int i = 0; i < F.N; i++)
            if (i % 3 != 0) {
                const float z = 3.6f + 0.8f * ((i * 37) % 100) / 100.f;
                F.mvDepth[i] = z; F.mvuRight[i] = F.mvKeysUn[i].pt.x - F.mbf / z;
            }
    F.mvpMapPoints.assign(F.N, static_cast<MapPoint*>(NULL));
    F.mvbOutlier.assign(F.N, false);
    F.AssignFeaturesToGrid();
    F.mpORBvocabulary = &w.voc;
    F.ComputeBoW();
}

Frame* make_frame(World& w, const uint8_t* img, ORBextractor* ex, int lap1)
{
    Frame* F = new Frame();
    F->mpORBextractorLeft = ex; F->mpORBextractorRight = NULL;
    cv::Mat im(w.H, w.W, CV_8UC1, (void*)img, (size_t)w.W);
    scenario_before_extract();
    F->ExtractORB(0, im, 0, lap1);
    scenario_after_extract();
    finish_frame(w, *F, ex, true);
    return F;
}

// a map point seen from a frame's feature, as MapPoint::MapPoint(Pos, pMap, pFrame, idxF) sets it up (R/src/MapPoint.cc:84-122)
MapPoint* point_from_feature(World& w, Frame& F, KeyFrame* pKF, int idx, float z)
{
    const cv::KeyPoint& kp = F.mvKeysUn[idx];
    cv::Mat Xc = (cv::Mat_<float>(3, 1) << (kp.pt.x - w.cx) / w.fx * z, (kp.pt.y - w.cy) / w.fy * z, z);
    cv::Mat Rwc = F.mTcw.rowRange(0, 3).colRange(0, 3).t();
    cv::Mat Ow = -Rwc * F.mTcw.rowRange(0, 3).col(3);
    cv::Mat Xw = Rwc * Xc + Ow;
    w.points.emplace_back(new MapPoint(Xw, pKF, &w.map));
    MapPoint* p = w.points.back().get();
    p->mDescriptor = F.mDescriptors.row(idx).clone();
    cv::Mat PC = Xw - Ow;
    const float dist = cv::norm(PC);
    p->mNormalVector = PC / dist;
    const float levelScale = F.mvScaleFactors[kp.octave];
    p->mfMaxDistance = dist * levelScale;
    p->mfMinDistance = p->mfMaxDistance / F.mvScaleFactors[F.mnScaleLevels - 1];
    return p;
}

// A two-camera frame as the rig constructor leaves it (R/src/Frame.cc:1040-1095), minus IMU and ComputeStereoFishEyeMatches
// (KannalaBrandt8 triangulation): left and right extraction, the stacked descriptor matrix, both grids, the rig calibration;
// the stereo partner tables hold mutual nearest descriptors (any consistent pairing serves the matcher).
Frame* make_rig_frame(World& w, const uint8_t* imgL, const uint8_t* imgR, ORBextractor* exL, ORBextractor* exR, const cv::Mat& Tlr)
{
    Frame* F = new Frame();
    F->mpORBextractorLeft = exL; F->mpORBextractorRight = exR;
    cv::Mat imL(w.H, w.W, CV_8UC1, (void*)imgL, (size_t)w.W), imR(w.H, w.W, CV_8UC1, (void*)imgR, (size_t)w.W);
    scenario_before_extract();
    F->ExtractORB(0, imL, 0, 0);
    scenario_before_extract();
    F->ExtractORB(1, imR, 0, 0);
    scenario_after_extract();
    F->Nleft = F->mvKeys.size(); F->Nright = F->mvKeysRight.size(); F->N = F->Nleft + F->Nright;
    F->mnScaleLevels = exL->GetLevels();
    F->mfScaleFactor = exL->GetScaleFactor();
    F->mfLogScaleFactor = log(F->mfScaleFactor);
    F->mvScaleFactors = exL->GetScaleFactors();
    F->mvInvScaleFactors = exL->GetInverseScaleFactors();
    F->mvLevelSigma2 = exL->GetScaleSigmaSquares();
    F->mvInvLevelSigma2 = exL->GetInverseScaleSigmaSquares();
    F->mK = w.K.clone(); F->mDistCoef = w.dist.clone();
    F->mbf = w.bf; F->mb = w.bf / w.fx; F->mThDepth = 35.f * F->mb;
    F->mpCamera = w.camera; F->mpCamera2 = w.camera;
    F->mTlr = Tlr.clone();
    F->mRlr = F->mTlr.rowRange(0, 3).colRange(0, 3);
    F->mtlr = F->mTlr.col(3);
    cv::Mat Rrl = F->mTlr.rowRange(0, 3).colRange(0, 3).t();
    cv::Mat trl = Rrl * (-1 * F->mTlr.col(3));
    cv::hconcat(Rrl, trl, F->mTrl);
    // stereo partners: mutual nearest neighbours under 40 bits
    F->mvLeftToRightMatch.assign(F->Nleft, -1); F->mvRightToLeftMatch.assign(F->Nright, -1);
    vector<int> bestR(F->Nleft, -1), bestL(F->Nright, -1), distL(F->Nright, 256);
    for (int i = 0; i < F->Nleft; i++) {
        int bd = 256;
        for (int j = 0; j < F->Nright; j++) {
            const int d = ORBmatcher::DescriptorDistance(F->mDescriptors.row(i), F->mDescriptorsRight.row(j));
            if (d < bd) { bd = d; bestR[i] = j; }
            if (d < distL[j]) { distL[j] = d; bestL[j] = i; }
        }
        if (bd >= 40) bestR[i] = -1;
    }
    for (int i = 0; i < F->Nleft; i++)
        if (bestR[i] >= 0 && bestL[bestR[i]] == i) { F->mvLeftToRightMatch[i] = bestR[i]; F->mvRightToLeftMatch[bestR[i]] = i; }
    // all descriptors in one matrix (cv::vconcat, :1081)
    cv::Mat all(F->N, 32, CV_8UC1);
    for (int i = 0; i < F->Nleft; i++) memcpy(all.ptr(i), F->mDescriptors.ptr(i), 32);
    for (int i = 0; i < F->Nright; i++) memcpy(all.ptr(F->Nleft + i), F->mDescriptorsRight.ptr(i), 32);
    F->mDescriptors = all;
    F->mvuRight.assign(F->Nleft, -1.f); F->mvDepth.assign(F->Nleft, -1.f);      // ComputeStereoFishEyeMatches, :1123-1124
    F->mvpMapPoints.assign(F->N, static_cast<MapPoint*>(NULL));
    F->mvbOutlier.assign(F->N, false);
    F->AssignFeaturesToGrid();
    F->UndistortKeyPoints();
    return F;
}

void hash_extraction(const char* name, const Frame& F)
{
    unsigned long long h = 1469598103934665603ull;
    auto mix = [&](const void* p, size_t n) { const unsigned char* b = (const unsigned char*)p; for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; } };
    if (F.N) mix(F.mvKeys.data(), sizeof(cv::KeyPoint) * F.mvKeys.size());
    for (int i = 0; i < F.mDescriptors.rows; i++) mix(F.mDescriptors.ptr(i), 32);
    fprintf(g_out, "%s N=%d mono=%d hash=%016llx bow=%zu fv=%zu\n", name, F.N, F.monoLeft, h, F.mBowVec.size(), F.mFeatVec.size());
}

}  // namespace

int main(int argc, char** argv)
{
    if (argc != 7) { fprintf(stderr, "usage: scenario frames.raw W H nframes vocabulary.txt out.txt\n"); return 2; }
    World w;
    w.W = atoi(argv[2]); w.H = atoi(argv[3]);
    const int nframes = atoi(argv[4]);
    if (nframes < 4) { fprintf(stderr, "need at least 4 stream frames\n"); return 2; }
    const size_t fsz = (size_t)w.W * w.H;
    vector<uint8_t> raw(fsz * (nframes + 2));
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(raw.data(), 1, raw.size(), f) != raw.size()) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    fclose(f);
    g_out = fopen(argv[6], "w");
    if (!g_out) return 2;
    try {
        w.fx = 435.2f; w.fy = 435.2f; w.cx = w.W * 0.5f - 0.3f; w.cy = w.H * 0.5f + 0.2f; w.bf = 40.f;
        w.camera = new Pinhole(vector<float>{w.fx, w.fy, w.cx, w.cy});
        w.K = w.camera->toK();
        w.dist = cv::Mat::zeros(4, 1, CV_32F);
        Frame::fx = w.fx; Frame::fy = w.fy; Frame::cx = w.cx; Frame::cy = w.cy; Frame::invfx = 1.0f / w.fx; Frame::invfy = 1.0f / w.fy;
        Frame::mnMinX = 0.f; Frame::mnMaxX = (float)w.W; Frame::mnMinY = 0.f; Frame::mnMaxY = (float)w.H;
        Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(Frame::mnMaxX - Frame::mnMinX);
        Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(Frame::mnMaxY - Frame::mnMinY);
        if (!w.voc.loadFromTextFile(argv[5])) throw std::runtime_error("vocabulary not loaded");

        ORBextractor ex(1000, 1.2f, 8, 20, 7), exIni(5000, 1.2f, 8, 20, 7), exRight(1000, 1.2f, 8, 20, 7);

        // ---- frames of the stream, with poses that follow the (3, 2) px / frame image motion at ~4 m depth ----
        const float z0 = 4.0f;
        vector<Frame*> F;
        for (int t = 0; t < nframes; t++) {
            F.push_back(make_frame(w, raw.data() + fsz * t, &ex, 0));
            F[t]->SetPose(pose(0.0015f * t, 3.0f * t * z0 / w.fx + 0.004f * t, 2.0f * t * z0 / w.fy, 0.03f * t));
            char nm[32]; snprintf(nm, sizeof(nm), "extract[%d]", t);
            hash_extraction(nm, *F[t]);
        }

        // ---- SearchForInitialization on two 5000-feature frames (Tracking::MonocularInitialization) ----
        {
            std::unique_ptr<Frame> A(make_frame(w, raw.data(), &exIni, 1000)), B(make_frame(w, raw.data() + fsz, &exIni, 1000));
            hash_extraction("extract_ini[0]", *A); hash_extraction("extract_ini[1]", *B);
            vector<cv::Point2f> prev(A->mvKeysUn.size());
            for (size_t i = 0; i < prev.size(); i++) prev[i] = A->mvKeysUn[i].pt;
            vector<int> m12;
            ORBmatcher matcher(0.9, true);
            int n = matcher.SearchForInitialization(*A, *B, prev, m12, 100);
            fprintf(g_out, "SearchForInitialization n=%d\n", n);
            dump_ints("  matches12", m12);
            vector<float> p; for (size_t i = 0; i < prev.size(); i++) { p.push_back(prev[i].x); p.push_back(prev[i].y); }
            dump_floats("  prevMatched", p);
            n = matcher.SearchForInitialization(*A, *B, prev, m12, 30);           // second call with the updated vbPrevMatched
            fprintf(g_out, "SearchForInitialization(2) n=%d\n", n);
            dump_ints("  matches12", m12);
        }

        // ---- the map: points from the features of frame 0 and frame 1; keyframes 0 and 1 ----
        vector<MapPoint*> pts0(F[0]->N, static_cast<MapPoint*>(NULL));
        for (int i = 0; i < F[0]->N; i++)
            if (i % 5 != 4) { pts0[i] = point_from_feature(w, *F[0], NULL, i, z0 * (0.9f + 0.2f * ((i * 61) % 100) / 100.f)); F[0]->mvpMapPoints[i] = pts0[i]; }
        KeyFrame KF0(*F[0]);
        for (int i = 0; i < F[0]->N; i++)
            if (pts0[i]) {
                pts0[i]->mpRefKF = &KF0;
                if (i % 7 != 0) pts0[i]->AddObservation(&KF0, i);       // the others stay without observations (temporal points)
            }
        for (int i = 0; i < F[0]->N; i++)
            if (pts0[i] && pts0[i]->Observations() > 0) pts0[i]->UpdateNormalAndDepth();
        vector<MapPoint*> pts1(F[1]->N, static_cast<MapPoint*>(NULL));
        for (int i = 0; i < F[1]->N; i++)
            if (i % 4 == 0) { pts1[i] = point_from_feature(w, *F[1], NULL, i, z0 * (0.92f + 0.16f * ((i * 29) % 100) / 100.f)); F[1]->mvpMapPoints[i] = pts1[i]; }
        KeyFrame KF1(*F[1]);
        for (int i = 0; i < F[1]->N; i++)
            if (pts1[i]) { pts1[i]->mpRefKF = &KF1; pts1[i]->AddObservation(&KF1, i); pts1[i]->UpdateNormalAndDepth(); }
        fprintf(g_out, "map points0=%zu points1=%zu\n", (size_t)std::count_if(pts0.begin(), pts0.end(), [](MapPoint* p) { return p != NULL; }),
                (size_t)std::count_if(pts1.begin(), pts1.end(), [](MapPoint* p) { return p != NULL; }));

        // ---- SearchByProjection(CurrentFrame, LastFrame): frame 0 (with its points, some without observations) -> frame 2 ----
        for (int variant = 0; variant < 3; variant++) {
            Frame& Cur = *F[2];
            Cur.mvpMapPoints.assign(Cur.N, static_cast<MapPoint*>(NULL));
            if (variant == 2)               // some slots already hold points: with observations (blocking) and without (not blocking)
                for (int i = 0; i < Cur.N; i += 9) Cur.mvpMapPoints[i] = pts0[(i * 3) % F[0]->N];
            ORBmatcher matcher(0.9, variant != 1);
            const int n = matcher.SearchByProjection(Cur, *F[0], variant == 1 ? 7.f : 15.f, variant == 1);
            fprintf(g_out, "SearchByProjection(Cur,Last)[%d] n=%d\n", variant, n);
            dump_points("  mvpMapPoints", Cur.mvpMapPoints);
        }

        // ---- SearchByProjection(F, vpMapPoints): the local map against frame 3 (Tracking::SearchLocalPoints) ----
        {
            Frame& Cur = *F[3];
            Cur.mvpMapPoints.assign(Cur.N, static_cast<MapPoint*>(NULL));
            for (int i = 0; i < Cur.N; i += 11) Cur.mvpMapPoints[i] = pts1[(i / 11 * 4) % F[1]->N];     // tracked so far
            vector<MapPoint*> local;
            for (size_t i = 0; i < pts0.size(); i++) if (pts0[i]) local.push_back(pts0[i]);
            for (size_t i = 0; i < pts1.size(); i++) if (pts1[i]) local.push_back(pts1[i]);
            int nvis = 0;
            for (size_t i = 0; i < local.size(); i++) nvis += Cur.isInFrustum(local[i], 0.5) ? 1 : 0;
            for (int th = 1; th <= 5; th += 2) {
                vector<MapPoint*> keep = Cur.mvpMapPoints;
                ORBmatcher matcher(0.8, true);
                const int n = matcher.SearchByProjection(Cur, local, (float)th, th == 5, 4.1f);
                fprintf(g_out, "SearchByProjection(F,MapPoints) th=%d visible=%d n=%d\n", th, nvis, n);
                dump_points("  mvpMapPoints", Cur.mvpMapPoints);
                Cur.mvpMapPoints = keep;
            }
        }

        // ---- SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist): relocalisation ----
        {
            Frame& Cur = *F[2];
            Cur.mvpMapPoints.assign(Cur.N, static_cast<MapPoint*>(NULL));
            std::set<MapPoint*> found;
            for (int i = 0; i < F[0]->N; i += 6) if (pts0[i]) { found.insert(pts0[i]); Cur.mvpMapPoints[(i * 5) % Cur.N] = pts0[i]; }
            ORBmatcher matcher(0.9, true);
            const int n = matcher.SearchByProjection(Cur, &KF0, found, 10, 100);
            fprintf(g_out, "SearchByProjection(reloc) n=%d\n", n);
            dump_points("  mvpMapPoints", Cur.mvpMapPoints);
            const int n2 = matcher.SearchByProjection(Cur, &KF0, found, 3, 64);
            fprintf(g_out, "SearchByProjection(reloc,narrow) n=%d\n", n2);
            dump_points("  mvpMapPoints", Cur.mvpMapPoints);
        }

        // ---- SearchByBoW, both overloads ----
        {
            ORBmatcher matcher(0.75, true);
            vector<MapPoint*> vp;
            int n = matcher.SearchByBoW(&KF0, *F[2], vp);
            fprintf(g_out, "SearchByBoW(KF,F) n=%d\n", n);
            dump_points("  matches", vp);
            ORBmatcher matcher2(0.8, true);
            n = matcher2.SearchByBoW(&KF0, &KF1, vp);
            fprintf(g_out, "SearchByBoW(KF,KF) n=%d\n", n);
            dump_points("  matches", vp);
        }

        // ---- SearchByProjection with a Sim3, both overloads (LoopClosing) ----
        {
            const cv::Mat T = KF1.GetPose();
            const float s = 1.02f;
            cv::Mat Scw = cv::Mat::eye(4, 4, CV_32F);
            for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) Scw.at<float>(r, c) = T.at<float>(r, c) * s;
            vector<MapPoint*> candidates;
            vector<KeyFrame*> candidateKFs;
            for (size_t i = 0; i < pts0.size(); i++) if (pts0[i] && pts0[i]->Observations() > 0) { candidates.push_back(pts0[i]); candidateKFs.push_back(&KF0); }
            vector<MapPoint*> matched(KF1.N, static_cast<MapPoint*>(NULL));
            for (int i = 0; i < KF1.N; i += 13) matched[i] = candidates[(i / 13) % candidates.size()];
            vector<MapPoint*> m1 = matched;
            ORBmatcher matcher(0.75, true);
            int n = matcher.SearchByProjection(&KF1, Scw, candidates, m1, 10, 1.0);
            fprintf(g_out, "SearchByProjection(Sim3) n=%d\n", n);
            dump_points("  vpMatched", m1);
            vector<MapPoint*> m2 = matched;
            vector<KeyFrame*> mkf(KF1.N, static_cast<KeyFrame*>(NULL));
            n = matcher.SearchByProjection(&KF1, Scw, candidates, candidateKFs, m2, mkf, 6, 1.5);
            fprintf(g_out, "SearchByProjection(Sim3,KFs) n=%d\n", n);
            dump_points("  vpMatched", m2);
            vector<int> kfids(mkf.size());
            for (size_t i = 0; i < mkf.size(); i++) kfids[i] = mkf[i] ? (int)mkf[i]->mnId : -1;
            dump_ints("  vpMatchedKF", kfids);
        }

        // ---- SearchForTriangulation, both overloads (LocalMapping::CreateNewMapPoints) ----
        {
            cv::Mat F12 = cv::Mat::eye(3, 3, CV_32F);          // ignored by the reference (the epipolar test rebuilds it from the poses)
            for (int variant = 0; variant < 3; variant++) {
                vector<std::pair<size_t, size_t> > pairs;
                ORBmatcher matcher(0.6, variant != 2);
                const int n = matcher.SearchForTriangulation(&KF0, &KF1, F12, pairs, variant == 1, variant == 2);
                fprintf(g_out, "SearchForTriangulation[%d] n=%d\n", variant, n);
                vector<int> flat;
                for (size_t i = 0; i < pairs.size(); i++) { flat.push_back((int)pairs[i].first); flat.push_back((int)pairs[i].second); }
                dump_ints("  pairs", flat);
            }
            vector<std::pair<size_t, size_t> > pairs;
            vector<cv::Mat> pts;
            ORBmatcher matcher(0.6, true);
            const int n = matcher.SearchForTriangulation(&KF0, &KF1, F12, pairs, false, pts);
            fprintf(g_out, "SearchForTriangulation(points) n=%d pairs=%zu\n", n, pairs.size());
        }

        // ---- SearchBySim3 (loop closing) ----
        {
            cv::Mat R1w = KF0.GetRotation(), t1w = KF0.GetTranslation(), R2w = KF1.GetRotation(), t2w = KF1.GetTranslation();
            cv::Mat R12 = R1w * R2w.t();
            cv::Mat t12 = -R12 * t2w + t1w;
            vector<MapPoint*> m12(KF0.N, static_cast<MapPoint*>(NULL));
            ORBmatcher matcher(0.75, true);
            const float s12 = 1.0f;
            int n = matcher.SearchBySim3(&KF0, &KF1, m12, s12, R12, t12, 7.5);
            fprintf(g_out, "SearchBySim3 n=%d\n", n);
            dump_points("  matches12", m12);
            n = matcher.SearchBySim3(&KF0, &KF1, m12, s12, R12, t12, 15.f);       // second round on the remaining points
            fprintf(g_out, "SearchBySim3(2) n=%d\n", n);
            dump_points("  matches12", m12);
        }

        // ---- Fuse with a Sim3 (LoopClosing::SearchAndFuse), then Fuse (LocalMapping::SearchInNeighbors); both change the map ----
        {
            cv::Mat Scw = KF1.GetPose();
            vector<MapPoint*> candidates;
            for (size_t i = 0; i < pts0.size(); i++) if (pts0[i] && pts0[i]->Observations() > 0 && i % 2 == 0) candidates.push_back(pts0[i]);
            vector<MapPoint*> replace(candidates.size(), static_cast<MapPoint*>(NULL));
            ORBmatcher matcher(0.8, true);
            int n = matcher.Fuse(&KF1, Scw, candidates, 4.f, replace);
            fprintf(g_out, "Fuse(Sim3) n=%d\n", n);
            dump_points("  vpReplacePoint", replace);
            dump_points("  KF1 points", KF1.GetMapPointMatches());
            vector<MapPoint*> all;
            for (size_t i = 0; i < pts0.size(); i++) all.push_back(pts0[i]);       // NULL entries included, as GetMapPointMatches gives them
            n = matcher.Fuse(&KF1, all, 3.f, false);
            fprintf(g_out, "Fuse n=%d\n", n);
            dump_points("  KF1 points", KF1.GetMapPointMatches());
            vector<int> bad, obs;
            for (size_t i = 0; i < w.points.size(); i++) { bad.push_back(w.points[i]->isBad() ? 1 : 0); obs.push_back(w.points[i]->Observations()); }
            dump_ints("  bad", bad); dump_ints("  observations", obs);
        }

        // ---- a stereo frame: both extractors + Frame::ComputeStereoMatches (reference body, reads mvImagePyramid of both) ----
        {
            Frame S;
            S.mpORBextractorLeft = &ex; S.mpORBextractorRight = &exRight;
            cv::Mat imL(w.H, w.W, CV_8UC1, raw.data() + fsz * nframes, (size_t)w.W), imR(w.H, w.W, CV_8UC1, raw.data() + fsz * (nframes + 1), (size_t)w.W);
            scenario_before_extract();
            S.ExtractORB(0, imL, 0, 0);
            scenario_before_extract();
            S.ExtractORB(1, imR, 0, 0);
            scenario_after_extract();
            finish_frame(w, S, &ex, false);
            S.ComputeStereoMatches();
            fprintf(g_out, "ComputeStereoMatches N=%d Nr=%zu\n", S.N, S.mvKeysRight.size());
            dump_floats("  mvuRight", S.mvuRight); dump_floats("  mvDepth", S.mvDepth);
#ifdef SCENARIO_DROPIN
            // the drop-in's device-side replacement must give the same two vectors (only where a device is behind the ABI)
            try {
                vector<float> u, d;
                ORBmatcher::ComputeStereoMatches(&ex, &exRight, S.mb, S.mbf, u, d);
                const bool same = u.size() == S.mvuRight.size() && (u.empty() || memcmp(u.data(), S.mvuRight.data(), 4 * u.size()) == 0) &&
                                  d.size() == S.mvDepth.size() && (d.empty() || memcmp(d.data(), S.mvDepth.data(), 4 * d.size()) == 0);
                fprintf(stderr, "device ComputeStereoMatches %s\n", same ? "identical" : "DIFFERS");
                if (!same) throw std::logic_error("ORBmatcher::ComputeStereoMatches differs from Frame::ComputeStereoMatches");
            } catch (const std::runtime_error& e) {
                fprintf(stderr, "device ComputeStereoMatches unavailable: %s\n", e.what());
            }
#endif
        }

        // ---- two-camera frames (Frame::Nleft != -1): the rig branches of the two tracking searches (R/src/ORBmatcher.cc:144-213,
        //      :2093-2160).  Left camera = stream frame t, right camera = stream frame t+1 (the scene shifted by (3, 2) px), so the
        //      rig calibration is the pose step between two stream frames. ----
        {
            cv::Mat Tlr = cv::Mat::eye(3, 4, CV_32F);
            Tlr.at<float>(0, 3) = -(3.0f * z0 / w.fx + 0.004f); Tlr.at<float>(1, 3) = -(2.0f * z0 / w.fy); Tlr.at<float>(2, 3) = -0.03f;
            std::unique_ptr<Frame> LastRig(make_rig_frame(w, raw.data(), raw.data() + fsz, &ex, &exRight, Tlr));
            std::unique_ptr<Frame> CurRig(make_rig_frame(w, raw.data() + fsz * 2, raw.data() + fsz * 3, &ex, &exRight, Tlr));
            LastRig->SetPose(F[0]->mTcw); CurRig->SetPose(F[2]->mTcw);
            fprintf(g_out, "rig frames: last %d + %d, current %d + %d, partners %zu\n", LastRig->Nleft, LastRig->Nright, CurRig->Nleft, CurRig->Nright,
                    (size_t)std::count_if(CurRig->mvLeftToRightMatch.begin(), CurRig->mvLeftToRightMatch.end(), [](int v) { return v >= 0; }));
            // the last rig frame sees the points of frame 0 with its left camera and those of frame 1 with its right camera
            // (same images and extractor parameters: the keypoints coincide index by index)
            if (LastRig->Nleft == F[0]->N && LastRig->Nright == F[1]->N) {
                for (int i = 0; i < F[0]->N; i++) LastRig->mvpMapPoints[i] = pts0[i];
                for (int i = 0; i < F[1]->N; i++) LastRig->mvpMapPoints[LastRig->Nleft + i] = pts1[i];
                for (int i = 0; i < LastRig->N; i += 17) LastRig->mvbOutlier[i] = true;
            }
            for (int variant = 0; variant < 3; variant++) {
                CurRig->mvpMapPoints.assign(CurRig->N, static_cast<MapPoint*>(NULL));
                if (variant == 2)
                    for (int i = 0; i < CurRig->N; i += 9) CurRig->mvpMapPoints[i] = pts0[(i * 3) % F[0]->N];
                ORBmatcher matcher(0.9, variant != 1);
                const Frame& Last = variant == 1 ? *F[0] : *LastRig;          // a one-camera last frame is legal too (:2018-2019)
                const int n = matcher.SearchByProjection(*CurRig, Last, variant == 1 ? 7.f : 15.f, variant == 1);
                fprintf(g_out, "SearchByProjection(CurRig,Last)[%d] n=%d\n", variant, n);
                dump_points("  mvpMapPoints", CurRig->mvpMapPoints);
            }
            {
                CurRig->mvpMapPoints.assign(CurRig->N, static_cast<MapPoint*>(NULL));
                for (int i = 0; i < CurRig->N; i += 11) CurRig->mvpMapPoints[i] = pts1[(i / 11 * 4) % F[1]->N];
                vector<MapPoint*> local;
                for (size_t i = 0; i < pts0.size(); i++) if (pts0[i]) local.push_back(pts0[i]);
                for (size_t i = 0; i < pts1.size(); i++) if (pts1[i]) local.push_back(pts1[i]);
                int nl = 0, nr = 0;
                for (size_t i = 0; i < local.size(); i++) { CurRig->isInFrustum(local[i], 0.5); nl += local[i]->mbTrackInView; nr += local[i]->mbTrackInViewR; }
                for (int th = 1; th <= 5; th += 2) {
                    vector<MapPoint*> keep = CurRig->mvpMapPoints;
                    ORBmatcher matcher(0.8, true);
                    const int n = matcher.SearchByProjection(*CurRig, local, (float)th, th == 5, 4.1f);
                    fprintf(g_out, "SearchByProjection(Rig,MapPoints) th=%d visible=%d/%d n=%d\n", th, nl, nr, n);
                    dump_points("  mvpMapPoints", CurRig->mvpMapPoints);
                    CurRig->mvpMapPoints = keep;
                }
            }

            // ---- SearchByBoW(KeyFrame*, Frame&) into a two-camera frame, from a one-camera and from a two-camera keyframe (:344-431) ----
            {
                CurRig->mpORBvocabulary = &w.voc; LastRig->mpORBvocabulary = &w.voc;
                CurRig->ComputeBoW(); LastRig->ComputeBoW();
                ORBmatcher matcher(0.75, true);
                vector<MapPoint*> vp;
                int n = matcher.SearchByBoW(&KF0, *CurRig, vp);
                fprintf(g_out, "SearchByBoW(KF,Rig) n=%d\n", n);
                dump_points("  matches", vp);
                KeyFrame KFL(*LastRig);
                ORBmatcher matcher2(0.9, false);
                n = matcher2.SearchByBoW(&KFL, *CurRig, vp);
                fprintf(g_out, "SearchByBoW(RigKF,Rig) n=%d\n", n);
                dump_points("  matches", vp);
            }

            // ---- Fuse on a two-camera keyframe, left camera then right camera (LocalMapping::SearchInNeighbors, :1395-1560) ----
            {
                // (the reference reads mvuRight[idx] with the RIGHT camera's index although the vector has Nleft entries, :1527: the
                // keyframe used here has Nright <= Nleft, so that the read stays inside the vector)
                const bool last_ok = LastRig->Nright <= LastRig->Nleft, cur_ok = CurRig->Nright <= CurRig->Nleft;
                Frame& Rig = (last_ok || !cur_ok) ? *LastRig : *CurRig;
                Rig.SetPose(&Rig == LastRig.get() ? F[0]->mTcw : F[2]->mTcw);
                Rig.mvpMapPoints.assign(Rig.N, static_cast<MapPoint*>(NULL));
                for (int i = 0; i < Rig.N; i += 7) Rig.mvpMapPoints[i] = pts1[(i / 7 * 4) % F[1]->N];
                KeyFrame KFR(Rig);
                vector<MapPoint*> all0, all1;
                for (size_t i = 0; i < pts0.size(); i++) all0.push_back(pts0[i]);
                for (size_t i = 0; i < pts1.size(); i++) all1.push_back(pts1[i]);
                ORBmatcher matcher(0.8, true);
                int n = matcher.Fuse(&KFR, all0, 3.f, false);
                fprintf(g_out, "Fuse(RigKF,left) n=%d\n", n);
                dump_points("  KFR points", KFR.GetMapPointMatches());
                if (last_ok || cur_ok) {
                    n = matcher.Fuse(&KFR, all1, 3.f, true);
                    fprintf(g_out, "Fuse(RigKF,right) n=%d\n", n);
                    dump_points("  KFR points", KFR.GetMapPointMatches());
                }
                vector<int> bad, obs;
                for (size_t i = 0; i < w.points.size(); i++) { bad.push_back(w.points[i]->isBad() ? 1 : 0); obs.push_back(w.points[i]->Observations()); }
                dump_ints("  bad", bad); dump_ints("  observations", obs);
            }
        }

        // ---- the brute-force matcher of Frame::ComputeStereoFishEyeMatches (R/src/Frame.cc:1112-1130): the descriptors of the lapping
        //      areas of two frames, knnMatch(k = 2); cv::BFMatcher in the reference build, BFMatcherB200 in the drop-in builds.  Also a
        //      train set of one row (one match per query) and an empty one ----
        {
#ifdef SCENARIO_DROPIN
            const BFMatcherB200 bf(cv::NORM_HAMMING);
#else
            const cv::BFMatcher bf(cv::NORM_HAMMING);
#endif
            const int monoLeft = F[0]->N / 3, monoRight = F[1]->N / 4;
            const cv::Mat dl = F[0]->mDescriptors.rowRange(monoLeft, F[0]->mDescriptors.rows);
            const cv::Mat dr = F[1]->mDescriptors.rowRange(monoRight, F[1]->mDescriptors.rows);
            const cv::Mat trains[3] = {dr, dr.rowRange(0, 1), dr.rowRange(0, 0)};
            for (int c = 0; c < 3; c++) {
                vector<vector<cv::DMatch> > mm;
                bf.knnMatch(dl, trains[c], mm, 2);
                vector<int> flat; int good = 0;
                for (size_t i = 0; i < mm.size(); i++) {
                    flat.push_back((int)mm[i].size());
                    for (size_t j = 0; j < mm[i].size(); j++) { flat.push_back(mm[i][j].queryIdx); flat.push_back(mm[i][j].trainIdx); flat.push_back((int)mm[i][j].distance); }
                    if (mm[i].size() >= 2 && mm[i][0].distance < mm[i][1].distance * 0.7) good++;      // Lowe's ratio of :1136
                }
                fprintf(g_out, "BFmatcher.knnMatch queries=%d train=%d ratio-accepted=%d\n", dl.rows, trains[c].rows, good);
                dump_ints("  matches", flat);
            }
        }

        // ---- DescriptorDistance ----
        {
            vector<int> d;
            for (int i = 0; i + 1 < F[0]->N && i < 400; i += 2) d.push_back(ORBmatcher::DescriptorDistance(F[0]->mDescriptors.row(i), F[1]->mDescriptors.row(i + 1)));
            dump_ints("DescriptorDistance", d);
        }
        for (size_t i = 0; i < F.size(); i++) delete F[i];
    } catch (const std::exception& e) {
        fprintf(stderr, "scenario failed: %s\n", e.what());
        fclose(g_out);
        return 1;
    }
    fclose(g_out);
    return 0;
}
