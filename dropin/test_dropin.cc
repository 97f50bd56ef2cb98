// Exercises the drop-in classes the way Frame / Tracking do (Frame.cc:393-400, Tracking.cc:2216-2217):
//   test_dropin <w> <h> <frame0.raw> <frame1.raw> <out.bin>
// writes: n0, mono0, keypoints0, descriptors0, n1, keypoints1, descriptors1, nmatches, vnMatches12, pyramid level 3,
//         one descriptor distance, then the two SearchByProjection results (count + per-keypoint map-point index)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ORBextractor.h"
#include "ORBmatcher.h"
#include "ORBVocabulary.h"
using namespace ORB_SLAM3;

static std::vector<unsigned char> slurp(const char* p, size_t n)
{
    std::vector<unsigned char> v(n);
    FILE* f = fopen(p, "rb");
    if (!f || fread(v.data(), 1, n, f) != n) { fprintf(stderr, "cannot read %s\n", p); exit(2); }
    fclose(f);
    return v;
}

int main(int argc, char** argv)
{
    if (argc < 6) return 1;
    const int w = atoi(argv[1]), h = atoi(argv[2]);
    std::vector<unsigned char> i0 = slurp(argv[3], (size_t)w * h), i1 = slurp(argv[4], (size_t)w * h);
    ORBextractor* ext = new ORBextractor(1000, 1.2f, 8, 20, 7);        // Tracking.cc:145
    Frame F[2];
    std::vector<int> lap = {0, 0};
    int mono[2];
    for (int k = 0; k < 2; k++) {
        cv::Mat im(h, w, CV_8UC1, k ? i1.data() : i0.data());
        mono[k] = (*ext)(im, cv::Mat(), F[k].mvKeys, F[k].mDescriptors, lap);
        F[k].mvKeysUn = F[k].mvKeys; F[k].N = (int)F[k].mvKeys.size();
    }
    Frame::mnMinX = 0; Frame::mnMaxX = (float)w; Frame::mnMinY = 0; Frame::mnMaxY = (float)h;
    std::vector<cv::Point2f> prev(F[0].mvKeysUn.size());
    for (size_t i = 0; i < prev.size(); i++) prev[i] = F[0].mvKeysUn[i].pt;
    std::vector<int> m12;
    ORBmatcher matcher(0.9f, true);
    const int nm = matcher.SearchForInitialization(F[0], F[1], prev, m12, 100);
    ext->SyncPyramidToHost();
    FILE* o = fopen(argv[5], "wb");
    for (int k = 0; k < 2; k++) {
        int n = F[k].N;
        fwrite(&n, 4, 1, o); fwrite(&mono[k], 4, 1, o);
        fwrite(F[k].mvKeys.data(), sizeof(cv::KeyPoint), n, o);
        for (int i = 0; i < n; i++) fwrite(F[k].mDescriptors.ptr(i), 1, 32, o);
    }
    fwrite(&nm, 4, 1, o);
    fwrite(m12.data(), 4, m12.size(), o);
    int lw = ext->mvImagePyramid[3].cols, lh = ext->mvImagePyramid[3].rows;
    fwrite(&lw, 4, 1, o); fwrite(&lh, 4, 1, o);
    for (int y = 0; y < lh; y++) fwrite(ext->mvImagePyramid[3].ptr(y), 1, lw, o);
    const int d = ORBmatcher::DescriptorDistance(F[0].mDescriptors.row(0), F[1].mDescriptors.row(0));
    fwrite(&d, 4, 1, o);

    // ---- SearchByProjection(Frame&, vector<MapPoint*>&, th): local-map tracking (Tracking::SearchLocalPoints) ----
    // one map point per keypoint of frame 0, predicted in frame 1 at its own position + the stream's (3, 2) motion
    const int n0 = F[0].N;
    std::vector<MapPoint> mps(n0);
    std::vector<MapPoint*> vmp(n0);
    for (int i = 0; i < n0; i++) {
        MapPoint& mp = mps[i];
        mp.mTrackProjX = F[0].mvKeysUn[i].pt.x + 3.f; mp.mTrackProjY = F[0].mvKeysUn[i].pt.y + 2.f; mp.mTrackProjXR = 0.f;
        mp.mnTrackScaleLevel = F[0].mvKeysUn[i].octave; mp.mTrackViewCos = (i % 3) ? 0.9995f : 0.9f;
        mp.mbTrackInView = (i % 7) != 0; mp.mDescriptor = F[0].mDescriptors.row(i);
        mp.mWorldPos[0] = mp.mTrackProjX; mp.mWorldPos[1] = mp.mTrackProjY; mp.mWorldPos[2] = 1.f;
        vmp[i] = &mp;
    }
    Frame A = F[1];
    A.mvScaleFactors = ext->GetScaleFactors();
    A.mvpMapPoints.assign(A.N, nullptr); A.mvbOutlier.assign(A.N, false);
    ORBmatcher mA(0.8f, true);
    const int nA = mA.SearchByProjection(A, vmp, 3.f);
    fwrite(&nA, 4, 1, o);
    for (int i = 0; i < A.N; i++) { int v = A.mvpMapPoints[i] ? (int)(A.mvpMapPoints[i] - mps.data()) : -1; fwrite(&v, 4, 1, o); }

    // ---- SearchByProjection(Frame& Current, const Frame& Last, th, bMono): TrackWithMotionModel ----
    GeometricCamera cam;                                            // identity pose and intrinsics: project(X) = (X, Y) / Z
    Frame Last = F[0];
    Last.mvScaleFactors = ext->GetScaleFactors();
    Last.mvpMapPoints = vmp; Last.mvbOutlier.assign(n0, false);
    for (int i = 0; i < n0; i += 11) Last.mvbOutlier[i] = true;
    Frame Cur = F[1];
    Cur.mvScaleFactors = ext->GetScaleFactors(); Cur.mpCamera = &cam;
    Cur.mvpMapPoints.assign(Cur.N, nullptr); Cur.mvbOutlier.assign(Cur.N, false);
    ORBmatcher mB(0.9f, true);
    const int nB = mB.SearchByProjection(Cur, Last, 7.f, true);
    fwrite(&nB, 4, 1, o);
    for (int i = 0; i < Cur.N; i++) { int v = Cur.mvpMapPoints[i] ? (int)(Cur.mvpMapPoints[i] - mps.data()) : -1; fwrite(&v, 4, 1, o); }

    // ---- Frame::ComputeStereoMatches: one extractor per camera (Frame.cc:92-95); frame 1 is the left view, frame 0 the
    //      right view (the stream moves by +3 px in x: disparity 3, 2 px of vertical offset inside the row band) ----
    ORBextractor* extR = new ORBextractor(1000, 1.2f, 8, 20, 7);
    Frame SL, SR;
    {
        cv::Mat iml(h, w, CV_8UC1, i1.data()), imr(h, w, CV_8UC1, i0.data());
        (*ext)(iml, cv::Mat(), SL.mvKeys, SL.mDescriptors, lap);
        (*extR)(imr, cv::Mat(), SR.mvKeys, SR.mDescriptors, lap);
    }
    std::vector<float> uR, depth;
    ORBmatcher::ComputeStereoMatches(ext, extR, 0.11f, 47.9f, uR, depth);
    int ns = (int)uR.size();
    fwrite(&ns, 4, 1, o);
    fwrite(uR.data(), 4, ns, o); fwrite(depth.data(), 4, ns, o);

    // ---- Frame::ComputeBoW: ORBVocabulary::loadFromTextFile + transform(features, BowVector, FeatureVector, 4) ----
    int nbow = -1;
    if (argc > 6) {
        ORBVocabulary voc;
        if (!voc.loadFromTextFile(argv[6])) { fprintf(stderr, "cannot load vocabulary %s\n", argv[6]); return 3; }
        std::vector<cv::Mat> vDesc;                                   // Converter::toDescriptorVector
        for (int i = 0; i < F[0].N; i++) vDesc.push_back(F[0].mDescriptors.row(i));
        DBoW2::BowVector bow; DBoW2::FeatureVector fv;
        voc.transform(vDesc, bow, fv, 4);
        nbow = (int)bow.size();
        fwrite(&nbow, 4, 1, o);
        for (DBoW2::BowVector::const_iterator it = bow.begin(); it != bow.end(); ++it) { unsigned w = it->first; double x = it->second; fwrite(&w, 4, 1, o); fwrite(&x, 8, 1, o); }
        int nfv = (int)fv.size();
        fwrite(&nfv, 4, 1, o);
        for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it) {
            unsigned nid = it->first; int c = (int)it->second.size();
            fwrite(&nid, 4, 1, o); fwrite(&c, 4, 1, o); fwrite(it->second.data(), 4, c, o);
        }
        unsigned nw = voc.size(); fwrite(&nw, 4, 1, o);

        // ---- ORBmatcher::SearchByBoW(KeyFrame*, Frame&) (relocalisation) and (KeyFrame*, KeyFrame*) (loop detection) ----
        KeyFrame KF1, KF2;
        KF1.N = F[0].N; KF1.mvKeys = F[0].mvKeys; KF1.mvKeysUn = F[0].mvKeysUn; KF1.mDescriptors = F[0].mDescriptors;
        KF2.N = F[1].N; KF2.mvKeys = F[1].mvKeys; KF2.mvKeysUn = F[1].mvKeysUn; KF2.mDescriptors = F[1].mDescriptors;
        voc.transform(KF1.mDescriptors, KF1.mBowVec, KF1.mFeatVec, 4);
        voc.transform(KF2.mDescriptors, KF2.mBowVec, KF2.mFeatVec, 4);
        std::vector<MapPoint> mps2(F[1].N);
        KF1.mvpMapPoints.assign(KF1.N, nullptr); KF2.mvpMapPoints.assign(KF2.N, nullptr);
        for (int i = 0; i < KF1.N; i++) { if (i % 6) KF1.mvpMapPoints[i] = &mps[i]; mps[i].bad = (i % 10) == 3; }
        for (int i = 0; i < KF2.N; i++) { if (i % 5) KF2.mvpMapPoints[i] = &mps2[i]; mps2[i].bad = (i % 13) == 4; }
        Frame FB = F[1];
        FB.mFeatVec = KF2.mFeatVec;
        ORBmatcher mbow(0.7f, true);
        std::vector<MapPoint*> vpm;
        const int nR = mbow.SearchByBoW(&KF1, FB, vpm);
        fwrite(&nR, 4, 1, o);
        for (int i = 0; i < FB.N; i++) { int v = vpm[i] ? (int)(vpm[i] - mps.data()) : -1; fwrite(&v, 4, 1, o); }
        ORBmatcher mloop(0.8f, true);
        std::vector<MapPoint*> vpm12;
        const int nL = mloop.SearchByBoW(&KF1, &KF2, vpm12);
        fwrite(&nL, 4, 1, o);
        for (int i = 0; i < KF1.N; i++) { int v = vpm12[i] ? (int)(vpm12[i] - mps2.data()) : -1; fwrite(&v, 4, 1, o); }
    }
    fclose(o);
    delete extR;
    printf("dropin ok: %d / %d keypoints, %d init matches, %d / %d projection matches\n", F[0].N, F[1].N, nm, nA, nB);
    return 0;
}
