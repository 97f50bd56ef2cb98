// Exercises the drop-in classes the way Frame / Tracking do (Frame.cc:393-400, Tracking.cc:2216-2217):
//   test_dropin <w> <h> <frame0.raw> <frame1.raw> <out.bin>
// writes: n0, mono0, keypoints0, descriptors0, n1, keypoints1, descriptors1, nmatches, vnMatches12
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ORBextractor.h"
#include "ORBmatcher.h"
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
    fclose(o);
    printf("dropin ok: %d / %d keypoints, %d matches\n", F[0].N, F[1].N, nm);
    return 0;
}
