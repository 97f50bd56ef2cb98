// cvmin_impl.cc - definitions of the image primitives declared in dropin/cvmin/cvmin.h, for the oracle builds only.
// Every routine forwards to the C restatement in oracle/orb_oracle.c that tests/test_oracle.py pins against cv2 4.13.0
// (resize at 125 size pairs, blur, FAST on 160 sub-images x 2 thresholds, fastAtan2 lattice, BFMatcher, undistortPoints).
// TEST INFRASTRUCTURE: only oracle/_ref links this file.
#include <opencv2/core/core.hpp>
#include "../orb_oracle.h"

namespace cv {

void resize(InputArray _src, OutputArray _dst, Size dsize, double fx, double fy, int interpolation)
{
    Mat src = _src.getMat();
    if (interpolation != INTER_LINEAR || src.type() != CV_8U || fx != 0 || fy != 0) cvmin_fail("resize: only u8 INTER_LINEAR with an explicit size");
    _dst.create(dsize.height, dsize.width, CV_8U);
    Mat dst = _dst.getMat();
    orc_resize_linear_u8(src.data, src.cols, src.rows, (int)src.step, dst.data, dst.cols, dst.rows, (int)dst.step);
}

static inline int reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) { if (p < 0) p = -p; else p = 2 * (n - 1) - p; }
    return p;
}

void copyMakeBorder(InputArray _src, OutputArray _dst, int top, int bottom, int left, int right, int borderType, const Scalar&)
{
    Mat src = _src.getMat();
    if ((borderType & ~BORDER_ISOLATED) != BORDER_REFLECT_101 || src.type() != CV_8U) cvmin_fail("copyMakeBorder: only u8 BORDER_REFLECT_101");
    _dst.create(src.rows + top + bottom, src.cols + left + right, CV_8U);
    Mat dst = _dst.getMat();
    // src may be the interior view of dst (ORBextractor.cc:1167): interior pixels map to themselves, border pixels only read
    // interior pixels, so the copy is safe in place
    for (int y = 0; y < dst.rows; y++) {
        const uchar* s = src.ptr(reflect101(y - top, src.rows));
        uchar* d = dst.ptr(y);
        for (int x = 0; x < dst.cols; x++) {
            d[x] = s[reflect101(x - left, src.cols)];
        }
    }
}

void GaussianBlur(InputArray _src, OutputArray _dst, Size ksize, double sigmaX, double sigmaY, int borderType)
{
    Mat src = _src.getMat();
    if (ksize.width != 7 || ksize.height != 7 || sigmaX != 2 || (sigmaY != 2 && sigmaY != 0) || borderType != BORDER_REFLECT_101 || src.type() != CV_8U)
        cvmin_fail("GaussianBlur: only u8 7x7 sigma 2 BORDER_REFLECT_101");
    Mat tmp(src.rows, src.cols, CV_8U);
    orc_gaussian_blur7(src.data, src.cols, src.rows, (int)src.step, tmp.data, (int)tmp.step);
    _dst.create(src.rows, src.cols, CV_8U);
    Mat dst = _dst.getMat();
    tmp.copyTo(_OutputArray(dst));
}

void FAST(InputArray _image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression)
{
    Mat img = _image.getMat();
    if (img.type() != CV_8U) cvmin_fail("FAST: u8 only");
    keypoints.clear();
    if (img.cols < 7 || img.rows < 7) return;
    static thread_local std::vector<int32_t> buf;
    const int cap = ((img.cols + 1) / 2) * ((img.rows + 1) / 2) * (nonmaxSuppression ? 1 : 4) + 16;
    if ((int)buf.size() < cap * 3) buf.resize((size_t)cap * 3);
    const int n = orc_fast9_16(img.data, img.cols, img.rows, (int)img.step, threshold, nonmaxSuppression ? 1 : 0, buf.data(), cap);
    if (n > cap) cvmin_fail("FAST: corner buffer too small");
    keypoints.reserve(n);
    for (int i = 0; i < n; i++)
        keypoints.push_back(KeyPoint((float)buf[3 * i], (float)buf[3 * i + 1], 7.f, -1.f, (float)buf[3 * i + 2]));
}

float fastAtan2(float y, float x) { return orc_fast_atan2(y, x); }

void KeyPointsFilter::retainBest(std::vector<KeyPoint>&, int) { cvmin_fail("KeyPointsFilter::retainBest (only reachable from the reference's dead ComputeKeyPointsOld)"); }

void undistortPoints(InputArray _src, OutputArray _dst, InputArray _K, InputArray _dist, InputArray _R, InputArray _P)
{
    // Frame::UndistortKeyPoints form (R/src/Frame.cc:721-754): src N x 2 CV_32F (reshaped to 2 channels by the caller and back),
    // K 3x3 CV_32F, dist 4 or 5 x 1 CV_32F, R empty, P = K
    Mat src = _src.getMat(), K = _K.getMat(), dist = _dist.getMat(), P = _P.getMat();
    if (!_R.empty() && _R.getMat().total() != 0) cvmin_fail("undistortPoints: R must be empty");
    if (src.type() != CV_32F || K.type() != CV_32F || dist.type() != CV_32F || P.type() != CV_32F) cvmin_fail("undistortPoints: CV_32F only");
    const int n = (int)(src.total() / 2);
    std::vector<OrcKeyPoint> in(n), out(n);
    for (int i = 0; i < n; i++) { in[i].x = src.at<float>(2 * i); in[i].y = src.at<float>(2 * i + 1); }
    float Kf[9], Pf[9], df[8] = {0};
    for (int i = 0; i < 9; i++) { Kf[i] = K.at<float>(i / 3, i % 3); Pf[i] = P.at<float>(i / 3, i % 3); }
    const int nd = (int)dist.total() >= 5 ? 5 : 4;
    for (int i = 0; i < nd; i++) df[i] = dist.at<float>(i);
    // the oracle routine short-cuts dist[0] == 0 as the caller does; cv::undistortPoints itself does not, so nudge nothing here:
    // Frame::UndistortKeyPoints never calls this with dist[0] == 0 (:723-727)
    orc_undistort_keypoints(in.data(), n, Kf, df, nd, Pf, out.data());
    _dst.create(src.rows, src.cols, CV_32F);
    Mat dst = _dst.getMat();
    for (int i = 0; i < n; i++) { dst.at<float>(2 * i) = out[i].x; dst.at<float>(2 * i + 1) = out[i].y; }
}

void BFMatcher::knnMatch(InputArray _q, InputArray _t, std::vector<std::vector<DMatch> >& matches, int k) const
{
    Mat q = _q.getMat(), t = _t.getMat();
    if (norm_ != NORM_HAMMING || k != 2 || q.cols != 32 || t.cols != 32 || !q.isContinuous() || !t.isContinuous()) cvmin_fail("BFMatcher: only NORM_HAMMING, k = 2, 32-byte rows");
    std::vector<int32_t> idx((size_t)q.rows * 2), dist((size_t)q.rows * 2);
    orc_bf_knn2(q.data, q.rows, t.data, t.rows, idx.data(), dist.data());
    matches.assign(q.rows, std::vector<DMatch>());
    for (int i = 0; i < q.rows; i++)
        for (int j = 0; j < 2; j++)
            if (idx[2 * i + j] >= 0) matches[i].push_back(DMatch(i, idx[2 * i + j], (float)dist[2 * i + j]));
}

}  // namespace cv
