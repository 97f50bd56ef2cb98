// orbx_search.cu - searches over explicit candidate sets: the generic CSR candidate matcher (inner loops of
// SearchForTriangulation / Fuse / SearchBySim3), ORBmatcher::SearchByBoW (R/src/ORBmatcher.cc:269-471, :819-959) and
// MapPoint::ComputeDistinctiveDescriptors (R/src/MapPoint.cc:448-524).
#include "orbx_match_internal.h"

namespace {

// generic CSR candidate matching: one warp per query, candidates in list order, top-2 by (distance, list position)
__global__ void __launch_bounds__(256) k_match_candidates(const uint8_t* q, int nq, const uint8_t* t, const int32_t* offsets,
                                                        const int32_t* indices, int32_t* out_idx, int32_t* out_dist)
{
    const int lane = threadIdx.x & 31;
    const int qi = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (qi >= nq) return;
    const int o0 = offsets[qi], o1 = offsets[qi + 1];
    const uint4 a0 = reinterpret_cast<const uint4*>(q)[2 * qi], a1 = reinterpret_cast<const uint4*>(q)[2 * qi + 1];
    int d0 = 0x7fffffff, k0 = 0x7fffffff, d1 = 0x7fffffff, k1 = 0x7fffffff;
    uint32_t e0 = 0, e1 = 0;
    for (int k = o0 + lane; k < o1; k += 32) {
        const int j = indices[k];
        const int d = hamming256(a0, a1, reinterpret_cast<const uint4*>(t)[2 * j], reinterpret_cast<const uint4*>(t)[2 * j + 1]);
        const int r = k - o0;
        if (d < d0) { d1 = d0; k1 = k0; e1 = e0; d0 = d; k0 = r; e0 = (uint32_t)j; }
        else if (d < d1) { d1 = d; k1 = r; e1 = (uint32_t)j; }
    }
    if (o1 - o0 > 65535) {     // ranks beyond 16 bits: fall back to the shuffle tree on (dist, rank) pairs
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const int od0 = __shfl_xor_sync(0xffffffffu, d0, o), ok0 = __shfl_xor_sync(0xffffffffu, k0, o);
            const uint32_t oe0 = __shfl_xor_sync(0xffffffffu, e0, o);
            const int od1 = __shfl_xor_sync(0xffffffffu, d1, o), ok1 = __shfl_xor_sync(0xffffffffu, k1, o);
            const uint32_t oe1 = __shfl_xor_sync(0xffffffffu, e1, o);
            int ld, lk; uint32_t le;
            if (od0 < d0 || (od0 == d0 && ok0 < k0)) { ld = d0; lk = k0; le = e0; d0 = od0; k0 = ok0; e0 = oe0; }
            else { ld = od0; lk = ok0; le = oe0; }
            if (od1 < d1 || (od1 == d1 && ok1 < k1)) { d1 = od1; k1 = ok1; e1 = oe1; }
            if (ld < d1 || (ld == d1 && lk < k1)) { d1 = ld; k1 = lk; e1 = le; }
        }
    } else {
        warp_top2(d0, k0, e0, d1, e1, k1);
    }
    if (lane == 0) {
        out_idx[2 * qi] = d0 == 0x7fffffff ? -1 : (int)e0; out_dist[2 * qi] = d0 == 0x7fffffff ? -1 : d0;
        out_idx[2 * qi + 1] = d1 == 0x7fffffff ? -1 : (int)e1; out_dist[2 * qi + 1] = d1 == 0x7fffffff ? -1 : d1;
    }
}

// ---- ORBmatcher::SearchByBoW (R/src/ORBmatcher.cc:269-471, :819-959) ----
// The FeatureVectors of both sides are CSR tables sorted by node id.  Features of different nodes never interact (a
// feature belongs to one node), so one warp owns one common node and walks its set-1 features in list order exactly as
// the reference does: lanes score the node's set-2 features that are still free, warp top-2 by (distance, list rank),
// threshold + ratio test, claim.  The rotation histogram is global: bins are counted with atomics and applied by
// k_bow_finish.
struct BowArgs {
    int mode;
    const orbx_keypoint* k1; const uint8_t* d1; const uint8_t* valid1; int n1;
    const int32_t* fv1_nodes; const int32_t* fv1_start; const int32_t* fv1_feat; int nfv1;
    const orbx_keypoint* k2; const uint8_t* d2; const uint8_t* valid2; int n2;
    const int32_t* fv2_nodes; const int32_t* fv2_start; const int32_t* fv2_feat; int nfv2;
    float nnratio; int check_ori;
    int32_t* matches12;      // [n1], preset to -1
    uint8_t* claimed2;       // [n2], preset to 0
    uint8_t* bin_of;         // [n1]
    int32_t* hist;           // [HISTO_LENGTH + 1]: bins, then the match count; preset to 0
};

__global__ void __launch_bounds__(256) k_bow_match(BowArgs A)
{
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (w >= A.nfv1) return;
    const int node = A.fv1_nodes[w];
    int lo = 0, hi = A.nfv2;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (A.fv2_nodes[mid] < node) lo = mid + 1; else hi = mid; }
    if (lo >= A.nfv2 || A.fv2_nodes[lo] != node) return;
    const int a0 = A.fv1_start[w], a1 = A.fv1_start[w + 1], b0 = A.fv2_start[lo], b1 = A.fv2_start[lo + 1];
    for (int ia = a0; ia < a1; ia++) {
        const int i1 = A.fv1_feat[ia];
        if (!A.valid1[i1]) continue;
        const uint4 q0 = reinterpret_cast<const uint4*>(A.d1)[2 * i1], q1 = reinterpret_cast<const uint4*>(A.d1)[2 * i1 + 1];
        int d0 = 0x7fffffff, r0 = 0x7fffffff, d1 = 0x7fffffff, r1 = 0x7fffffff;
        uint32_t e0 = 0, e1 = 0;
        for (int ib = b0 + lane; ib < b1; ib += 32) {
            const int i2 = A.fv2_feat[ib];
            if (A.claimed2[i2]) continue;
            if (A.mode == 1 && A.valid2 && !A.valid2[i2]) continue;
            const int d = hamming256(q0, q1, reinterpret_cast<const uint4*>(A.d2)[2 * i2], reinterpret_cast<const uint4*>(A.d2)[2 * i2 + 1]);
            const int r = ib - b0;
            if (d < d0) { d1 = d0; r1 = r0; e1 = e0; d0 = d; r0 = r; e0 = (uint32_t)i2; }
            else if (d < d1) { d1 = d; r1 = r; e1 = (uint32_t)i2; }
        }
        warp_top2(d0, r0, e0, d1, e1, r1);
        const int best1 = d0 == 0x7fffffff ? 256 : d0, best2 = d1 == 0x7fffffff ? 256 : d1;
        const bool pass = A.mode == 0 ? best1 <= ORBX_TH_LOW : best1 < ORBX_TH_LOW;
        if (pass && (float)best1 < __fmul_rn(A.nnratio, (float)best2)) {
            if (lane == 0) {
                A.matches12[i1] = (int32_t)e0; A.claimed2[e0] = 1;
                if (A.check_ori) { const int bin = rot_bin(A.k1[i1].angle, A.k2[e0].angle); A.bin_of[i1] = (uint8_t)bin; atomicAdd(&A.hist[bin], 1); }
                atomicAdd(&A.hist[ORBX_HISTO_LENGTH], 1);
            }
        }
        __syncwarp();          // the claim is visible to every lane before the next set-1 feature is scored
    }
}

// rotation consistency (:437-460): keep the three dominant bins
__global__ void __launch_bounds__(256) k_bow_finish(BowArgs A, int32_t* nmatches)
{
    __shared__ int removed;
    if (threadIdx.x == 0) removed = 0;
    __syncthreads();
    if (A.check_ori) {
        int ind1, ind2, ind3;
        three_maxima(A.hist, ind1, ind2, ind3);
        int local = 0;
        for (int i = threadIdx.x; i < A.n1; i += blockDim.x)
            if (A.matches12[i] >= 0) {
                const int bin = A.bin_of[i];
                if (bin != ind1 && bin != ind2 && bin != ind3) { A.matches12[i] = -1; local++; }
            }
        if (local) atomicAdd(&removed, local);
    }
    __syncthreads();
    if (threadIdx.x == 0) *nmatches = A.hist[ORBX_HISTO_LENGTH] - removed;
}

// ---- SearchByBoW(KeyFrame*, Frame&) on a two-camera frame (Frame::Nleft != -1, R/src/ORBmatcher.cc:344-431) ----
// The frame's features [0, n2_left) are the left camera's, the rest the right camera's.  Per keyframe feature the reference keeps a
// (best, second) pair for each camera; the left match needs best <= TH_LOW and the ratio test, the right match needs
// bestLeft <= TH_LOW (it sits inside that branch) and bestRight <= TH_LOW (its ratio test is disabled by `|| true`, :402).
// Both claim their feature and enter the one rotation histogram.
__global__ void __launch_bounds__(256) k_bow_match_rig(BowArgs A, int n2_left, int32_t* matches12r, uint8_t* bin_of_r)
{
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (w >= A.nfv1) return;
    const int node = A.fv1_nodes[w];
    int lo = 0, hi = A.nfv2;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (A.fv2_nodes[mid] < node) lo = mid + 1; else hi = mid; }
    if (lo >= A.nfv2 || A.fv2_nodes[lo] != node) return;
    const int a0 = A.fv1_start[w], a1 = A.fv1_start[w + 1], b0 = A.fv2_start[lo], b1 = A.fv2_start[lo + 1];
    for (int ia = a0; ia < a1; ia++) {
        const int i1 = A.fv1_feat[ia];
        if (!A.valid1[i1]) continue;
        const uint4 q0 = reinterpret_cast<const uint4*>(A.d1)[2 * i1], q1 = reinterpret_cast<const uint4*>(A.d1)[2 * i1 + 1];
        int d0 = 0x7fffffff, r0 = 0x7fffffff, d1 = 0x7fffffff, r1 = 0x7fffffff;          // left camera
        int f0 = 0x7fffffff, s0 = 0x7fffffff, f1 = 0x7fffffff, s1 = 0x7fffffff;          // right camera
        uint32_t e0 = 0, e1 = 0, g0 = 0, g1 = 0;
        for (int ib = b0 + lane; ib < b1; ib += 32) {
            const int i2 = A.fv2_feat[ib];
            if (A.claimed2[i2]) continue;
            const int d = hamming256(q0, q1, reinterpret_cast<const uint4*>(A.d2)[2 * i2], reinterpret_cast<const uint4*>(A.d2)[2 * i2 + 1]);
            const int r = ib - b0;
            if (i2 < n2_left) {
                if (d < d0) { d1 = d0; r1 = r0; e1 = e0; d0 = d; r0 = r; e0 = (uint32_t)i2; }
                else if (d < d1) { d1 = d; r1 = r; e1 = (uint32_t)i2; }
            } else {
                if (d < f0) { f1 = f0; s1 = s0; g1 = g0; f0 = d; s0 = r; g0 = (uint32_t)i2; }
                else if (d < f1) { f1 = d; s1 = r; g1 = (uint32_t)i2; }
            }
        }
        warp_top2(d0, r0, e0, d1, e1, r1);
        warp_top2(f0, s0, g0, f1, g1, s1);
        const int best1 = d0 == 0x7fffffff ? 256 : d0, best2 = d1 == 0x7fffffff ? 256 : d1, best1R = f0 == 0x7fffffff ? 256 : f0;
        if (best1 <= ORBX_TH_LOW && lane == 0) {
            if ((float)best1 < __fmul_rn(A.nnratio, (float)best2)) {
                A.matches12[i1] = (int32_t)e0; A.claimed2[e0] = 1;
                if (A.check_ori) { const int bin = rot_bin(A.k1[i1].angle, A.k2[e0].angle); A.bin_of[i1] = (uint8_t)bin; atomicAdd(&A.hist[bin], 1); }
                atomicAdd(&A.hist[ORBX_HISTO_LENGTH], 1);
            }
            if (best1R <= ORBX_TH_LOW) {
                matches12r[i1] = (int32_t)g0; A.claimed2[g0] = 1;
                if (A.check_ori) { const int bin = rot_bin(A.k1[i1].angle, A.k2[g0].angle); bin_of_r[i1] = (uint8_t)bin; atomicAdd(&A.hist[bin], 1); }
                atomicAdd(&A.hist[ORBX_HISTO_LENGTH], 1);
            }
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(256) k_bow_finish_rig(BowArgs A, int32_t* matches12r, const uint8_t* bin_of_r, int32_t* nmatches)
{
    __shared__ int removed;
    if (threadIdx.x == 0) removed = 0;
    __syncthreads();
    if (A.check_ori) {
        int ind1, ind2, ind3;
        three_maxima(A.hist, ind1, ind2, ind3);
        int local = 0;
        for (int i = threadIdx.x; i < A.n1; i += blockDim.x) {
            if (A.matches12[i] >= 0) { const int bin = A.bin_of[i]; if (bin != ind1 && bin != ind2 && bin != ind3) { A.matches12[i] = -1; local++; } }
            if (matches12r[i] >= 0) { const int bin = bin_of_r[i]; if (bin != ind1 && bin != ind2 && bin != ind3) { matches12r[i] = -1; local++; } }
        }
        if (local) atomicAdd(&removed, local);
    }
    __syncthreads();
    if (threadIdx.x == 0) *nmatches = A.hist[ORBX_HISTO_LENGTH] - removed;
}

// ---- ORBmatcher::SearchForTriangulation (R/src/ORBmatcher.cc:961-1202), pinhole, no second camera ----
// Same node-by-node structure as SearchByBoW, but the queries of this fork do not interact (vbMatched2 is never set), so a
// warp handles one (node, keyframe-1 feature) at a time without claims: lanes score the node's free keyframe-2 features
// (distance <= TH_LOW, epipole distance, epipolar line), the best is the smallest distance and among equals the LAST in
// list order (the reference's `dist > bestDist` test lets a later equal candidate replace an earlier one).
struct TriArgs {
    BowArgs B;                         // feature sets, FeatureVectors, outputs (valid1 / valid2 = "has no MapPoint")
    const uint8_t* stereo1; const uint8_t* stereo2;
    float F[9]; float ep_x, ep_y;
    float scale2[ORBX_MAX_LEVELS], sigma2_2[ORBX_MAX_LEVELS];
    int only_stereo, coarse;
};

__global__ void __launch_bounds__(256) k_triangulation_match(TriArgs T)
{
    const BowArgs& A = T.B;
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (w >= A.nfv1) return;
    const int node = A.fv1_nodes[w];
    int lo = 0, hi = A.nfv2;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (A.fv2_nodes[mid] < node) lo = mid + 1; else hi = mid; }
    if (lo >= A.nfv2 || A.fv2_nodes[lo] != node) return;
    const int a0 = A.fv1_start[w], a1 = A.fv1_start[w + 1], b0 = A.fv2_start[lo], b1 = A.fv2_start[lo + 1];
    for (int ia = a0; ia < a1; ia++) {
        const int i1 = A.fv1_feat[ia];
        if (!A.valid1[i1]) continue;
        const bool st1 = T.stereo1 ? T.stereo1[i1] != 0 : false;
        if (T.only_stereo && !st1) continue;
        const orbx_keypoint kp1 = A.k1[i1];
        // epipolar line of kp1 in image 2: l = x1' F12 (Pinhole.cpp:129-131), float operations in the reference's order
        const float la = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, T.F[0]), __fmul_rn(kp1.y, T.F[3])), T.F[6]);
        const float lb = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, T.F[1]), __fmul_rn(kp1.y, T.F[4])), T.F[7]);
        const float lc = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, T.F[2]), __fmul_rn(kp1.y, T.F[5])), T.F[8]);
        const float den = __fadd_rn(__fmul_rn(la, la), __fmul_rn(lb, lb));
        const uint4 q0 = reinterpret_cast<const uint4*>(A.d1)[2 * i1], q1 = reinterpret_cast<const uint4*>(A.d1)[2 * i1 + 1];
        unsigned best = 0xFFFFFFFFu;                   // dist << 16 | (0xFFFF - rank): smallest distance, then the latest candidate
        for (int ib = b0 + lane; ib < b1; ib += 32) {
            const int i2 = A.fv2_feat[ib];
            if (!A.valid2[i2]) continue;
            const bool st2 = T.stereo2 ? T.stereo2[i2] != 0 : false;
            if (T.only_stereo && !st2) continue;
            const int d = hamming256(q0, q1, reinterpret_cast<const uint4*>(A.d2)[2 * i2], reinterpret_cast<const uint4*>(A.d2)[2 * i2 + 1]);
            if (d > ORBX_TH_LOW) continue;
            const orbx_keypoint kp2 = A.k2[i2];
            const int oct = min(max(kp2.octave, 0), ORBX_MAX_LEVELS - 1);
            if (!st1 && !st2) {
                const float ex = __fsub_rn(T.ep_x, kp2.x), ey = __fsub_rn(T.ep_y, kp2.y);
                if (__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)) < __fmul_rn(100.0f, T.scale2[oct])) continue;
            }
            if (!T.coarse) {
                if (den == 0.0f) continue;
                const float num = __fadd_rn(__fadd_rn(__fmul_rn(la, kp2.x), __fmul_rn(lb, kp2.y)), lc);
                const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
                if (!((double)dsqr < __dmul_rn(3.84, (double)T.sigma2_2[oct]))) continue;
            }
            const unsigned key = ((unsigned)d << 16) | (unsigned)(0xFFFF - (ib - b0));
            best = min(best, key);
        }
        best = __reduce_min_sync(0xffffffffu, best);
        if (best != 0xFFFFFFFFu && lane == 0) {
            const int i2 = A.fv2_feat[b0 + (0xFFFF - (int)(best & 0xFFFF))];
            A.matches12[i1] = i2;
            if (A.check_ori) { const int bin = rot_bin(kp1.angle, A.k2[i2].angle); A.bin_of[i1] = (uint8_t)bin; atomicAdd(&A.hist[bin], 1); }
            atomicAdd(&A.hist[ORBX_HISTO_LENGTH], 1);
        }
    }
}

// ---- MapPoint::ComputeDistinctiveDescriptors (R/src/MapPoint.cc:448-524), batched over map points ----
// One warp per map point.  For observation i the lanes compute its distances to all observations (kept in registers for
// up to 256 of them, recomputed beyond), and the median of the row (its (N-1)/2-th smallest value, the 0 of the diagonal
// included) is found by bisection on the value range [0, 256] with ballot counts instead of a sort.
constexpr int DD_R = 8;
__global__ void __launch_bounds__(256) k_distinctive(const uint8_t* desc, const int32_t* offsets, int npoints, int32_t* best)
{
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (p >= npoints) return;
    const int o = offsets[p], N = offsets[p + 1] - o;
    if (N <= 0) { if (lane == 0) best[p] = -1; return; }
    const uint4* D = reinterpret_cast<const uint4*>(desc) + 2 * (size_t)o;
    const int k = (N - 1) >> 1;                       // (int)(0.5 * (N - 1))
    const int steps = (N + 31) >> 5;
    int bestMedian = 0x7fffffff, bestIdx = 0;
    for (int i = 0; i < N; i++) {
        const uint4 q0 = D[2 * i], q1 = D[2 * i + 1];
        int cache[DD_R];
#pragma unroll
        for (int s = 0; s < DD_R; s++) {
            const int j = s * 32 + lane;
            cache[s] = (s < steps && j < N) ? hamming256(q0, q1, D[2 * j], D[2 * j + 1]) : 0x7fff;
        }
        int lo = 0, hi = 256;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            int cnt = 0;
#pragma unroll
            for (int s = 0; s < DD_R; s++) cnt += __popc(__ballot_sync(0xffffffffu, cache[s] <= mid));
            for (int s = DD_R; s < steps; s++) {
                const int j = s * 32 + lane;
                const bool le = j < N && hamming256(q0, q1, D[2 * j], D[2 * j + 1]) <= mid;
                cnt += __popc(__ballot_sync(0xffffffffu, le));
            }
            if (cnt >= k + 1) hi = mid; else lo = mid + 1;
        }
        if (lo < bestMedian) { bestMedian = lo; bestIdx = i; }
    }
    if (lane == 0) best[p] = bestIdx;
}

// ---- Frame::UndistortKeyPoints (R/src/Frame.cc:721-754) = cv::undistortPoints(pts, K, distCoef, I, K_new) ----
// Double arithmetic in OpenCV's operation order with explicitly rounded operations (no FMA contraction), exactly five
// iterations of the radial-tangential fixed point, result rounded to float: bit-identical to the CPU.
struct UndistortArgs { double k[12]; double fx, fy, cx, cy, ifx, ify; double RR[9]; };

__global__ void __launch_bounds__(256) k_undistort(const orbx_keypoint* in, const int32_t* n_slot, int n_fixed, int cap, UndistortArgs A,
                                                   orbx_keypoint* out)
{
    const int slot = blockIdx.y;
    const int n = n_slot ? n_slot[slot] : n_fixed;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    orbx_keypoint kp = in[(size_t)slot * cap + i];
    const double u = kp.x, v = kp.y;
    double x = __dmul_rn(__dsub_rn(u, A.cx), A.ifx), y = __dmul_rn(__dsub_rn(v, A.cy), A.ify);
    const double x0 = x, y0 = y;
    const double* k = A.k;
#pragma unroll 1
    for (int j = 0; j < 5; j++) {
        const double r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
        const double num = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(k[7], r2), k[6]), r2), k[5]), r2));
        const double den = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(k[4], r2), k[1]), r2), k[0]), r2));
        const double icdist = __ddiv_rn(num, den);
        if (icdist < 0) { x = __dmul_rn(__dsub_rn(u, A.cx), A.ifx); y = __dmul_rn(__dsub_rn(v, A.cy), A.ify); break; }
        // deltaX = 2*k2*x*y + k3*(r2 + 2*x*x) + k8*r2 + k9*r2*r2, left to right
        const double dX = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(2.0, k[2]), x), y),
                                                        __dmul_rn(k[3], __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, x), x)))),
                                              __dmul_rn(k[8], r2)), __dmul_rn(__dmul_rn(k[9], r2), r2));
        const double dY = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(k[2], __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, y), y))),
                                                        __dmul_rn(__dmul_rn(__dmul_rn(2.0, k[3]), x), y)),
                                              __dmul_rn(k[10], r2)), __dmul_rn(__dmul_rn(k[11], r2), r2));
        x = __dmul_rn(__dsub_rn(x0, dX), icdist);
        y = __dmul_rn(__dsub_rn(y0, dY), icdist);
    }
    const double xx = __dadd_rn(__dadd_rn(__dmul_rn(A.RR[0], x), __dmul_rn(A.RR[1], y)), A.RR[2]);
    const double yy = __dadd_rn(__dadd_rn(__dmul_rn(A.RR[3], x), __dmul_rn(A.RR[4], y)), A.RR[5]);
    const double ww = __ddiv_rn(1.0, __dadd_rn(__dadd_rn(__dmul_rn(A.RR[6], x), __dmul_rn(A.RR[7], y)), A.RR[8]));
    kp.x = (float)__dmul_rn(xx, ww); kp.y = (float)__dmul_rn(yy, ww);
    out[(size_t)slot * cap + i] = kp;
}

// ---- KF.msg wire format of the keypoints (msg/CvKeyPoint.msg, R/src/Converter.cc:218-244): 15 packed bytes ----
__global__ void k_kp_to_msg(const orbx_keypoint* kps, const int32_t* n_dev, int n_fixed, uint8_t* msg)
{
    const int n = n_dev ? *n_dev : n_fixed;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const orbx_keypoint k = kps[i];
    uint8_t* o = msg + (size_t)i * 15;
    const unsigned x = __float_as_uint(k.x), y = __float_as_uint(k.y), a = __float_as_uint(k.angle);
    o[0] = x; o[1] = x >> 8; o[2] = x >> 16; o[3] = x >> 24;
    o[4] = y; o[5] = y >> 8; o[6] = y >> 16; o[7] = y >> 24;
    o[8] = (uint8_t)(int)k.size;                          // (u_int8_t)kp.size, :228
    o[9] = a; o[10] = a >> 8; o[11] = a >> 16; o[12] = a >> 24;
    o[13] = (uint8_t)(int)k.response;                     // (u_int8_t)kp.response, :227
    o[14] = (uint8_t)(int8_t)k.octave;
}

__global__ void k_kp_from_msg(const uint8_t* msg, int n, orbx_keypoint* kps)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* o = msg + (size_t)i * 15;
    orbx_keypoint k;
    k.x = __uint_as_float((unsigned)o[0] | ((unsigned)o[1] << 8) | ((unsigned)o[2] << 16) | ((unsigned)o[3] << 24));
    k.y = __uint_as_float((unsigned)o[4] | ((unsigned)o[5] << 8) | ((unsigned)o[6] << 16) | ((unsigned)o[7] << 24));
    k.size = (float)o[8];
    k.angle = __uint_as_float((unsigned)o[9] | ((unsigned)o[10] << 8) | ((unsigned)o[11] << 16) | ((unsigned)o[12] << 24));
    k.response = (float)o[13]; k.octave = (int)(int8_t)o[14]; k.class_id = -1;
    kps[i] = k;
}

}  // namespace

// device-to-device unpack of n wire records (used by the keyframe DB, orbx_kfdb.cu)
void orbx_launch_kp_from_msg(const uint8_t* d_msg15, int n, orbx_keypoint* d_kps, cudaStream_t s)
{
    if (n <= 0) return;
    k_kp_from_msg<<<(n + 255) / 256, 256, 0, s>>>(d_msg15, n, d_kps); ORBX_COUNT_LAUNCH(1);
}

// Keypoints -> KF.msg records (15 bytes each) and back; host pointers, synchronous.
extern "C" int orbx_keypoints_to_msg(orbx_matcher* m, const orbx_keypoint* kps, int n, uint8_t* msg15)
{
    if (!m || n < 0 || (n > 0 && (!kps || !msg15))) return ORBX_E_INVALID;
    if (n == 0) return ORBX_OK;
    CKM(cudaSetDevice(m->p.device));
    int rc = orbx_m_gen_scratch(m, (sizeof(orbx_keypoint) + 16) * (size_t)n);
    if (rc) return rc;
    cudaStream_t s = m->stream;
    orbx_keypoint* dk = reinterpret_cast<orbx_keypoint*>(m->d_gen); uint8_t* dm = reinterpret_cast<uint8_t*>(dk + n);
    CKM(cudaMemcpyAsync(dk, kps, sizeof(orbx_keypoint) * n, cudaMemcpyHostToDevice, s));
    k_kp_to_msg<<<(n + 255) / 256, 256, 0, s>>>(dk, nullptr, n, dm); ORBX_COUNT_LAUNCH(1);
    CKM(cudaGetLastError());
    CKM(cudaMemcpyAsync(msg15, dm, (size_t)15 * n, cudaMemcpyDeviceToHost, s));
    CKM(cudaStreamSynchronize(s));
    return ORBX_OK;
}

extern "C" int orbx_keypoints_from_msg(orbx_matcher* m, const uint8_t* msg15, int n, orbx_keypoint* kps)
{
    if (!m || n < 0 || (n > 0 && (!kps || !msg15))) return ORBX_E_INVALID;
    if (n == 0) return ORBX_OK;
    CKM(cudaSetDevice(m->p.device));
    int rc = orbx_m_gen_scratch(m, (sizeof(orbx_keypoint) + 16) * (size_t)n);
    if (rc) return rc;
    cudaStream_t s = m->stream;
    orbx_keypoint* dk = reinterpret_cast<orbx_keypoint*>(m->d_gen); uint8_t* dm = reinterpret_cast<uint8_t*>(dk + n);
    CKM(cudaMemcpyAsync(dm, msg15, (size_t)15 * n, cudaMemcpyHostToDevice, s));
    k_kp_from_msg<<<(n + 255) / 256, 256, 0, s>>>(dm, n, dk); ORBX_COUNT_LAUNCH(1);
    CKM(cudaGetLastError());
    CKM(cudaMemcpyAsync(kps, dk, sizeof(orbx_keypoint) * n, cudaMemcpyDeviceToHost, s));
    CKM(cudaStreamSynchronize(s));
    return ORBX_OK;
}

// The keypoints of one result slot of an extractor as KF.msg records, written to a DEVICE buffer of
// orbx_extractor_max_keypoints(ex) * 15 bytes (the descriptors of the slot are already the 32-byte rows of Descriptor.msg);
// asynchronous on `stream`.
extern "C" int orbx_slot_keypoints_to_msg_device(orbx_extractor* ex, int slot, uint8_t* d_msg15, void* stream)
{
    if (!ex || !d_msg15) return ORBX_E_INVALID;
    orbx_keypoint* dk; uint8_t* dd; int32_t* dn; int cap, slots;
    int rc = orbx_extractor_results_device(ex, &dk, &dd, &dn, nullptr, &cap, &slots);
    if (rc) return rc;
    if (slot < 0 || slot >= slots) return ORBX_E_INVALID;
    CKM(cudaSetDevice(orbx_ex_device(ex)));
    cudaStream_t s = stream ? (cudaStream_t)stream : orbx_ex_stream(ex);
    k_kp_to_msg<<<(cap + 255) / 256, 256, 0, s>>>(dk + (size_t)slot * cap, dn + slot, 0, d_msg15); ORBX_COUNT_LAUNCH(1);
    CKM(cudaGetLastError());
    return ORBX_OK;
}

static int undistort_args(const float* K, const float* dist, int ndist, const float* P, UndistortArgs* A)
{
    if (!K || !P || !dist || ndist < 4 || ndist > 12) return ORBX_E_INVALID;
    for (int i = 0; i < 12; i++) A->k[i] = i < ndist ? (double)dist[i] : 0.0;
    A->fx = K[0]; A->fy = K[4]; A->cx = K[2]; A->cy = K[5];
    A->ifx = 1. / A->fx; A->ify = 1. / A->fy;
    for (int i = 0; i < 9; i++) A->RR[i] = (double)P[i];
    return ORBX_OK;
}

// Frame::UndistortKeyPoints on host arrays (see include/orbx.h).  Synchronous.
extern "C" int orbx_undistort_keypoints(orbx_matcher* m, const orbx_keypoint* kps, int n, const float* K, const float* dist, int ndist,
                                        const float* P, orbx_keypoint* kps_un)
{
    if (!m || n < 0 || (n > 0 && (!kps || !kps_un))) return ORBX_E_INVALID;
    if (n == 0) return ORBX_OK;
    UndistortArgs A;
    int rc = undistort_args(K, dist, ndist, P, &A);
    if (rc) return rc;
    if (dist[0] == 0.0f) { memcpy(kps_un, kps, sizeof(orbx_keypoint) * n); return ORBX_OK; }       // R/src/Frame.cc:723-727
    CKM(cudaSetDevice(m->p.device));
    if ((rc = orbx_m_gen_scratch(m, 2 * sizeof(orbx_keypoint) * (size_t)n))) return rc;
    cudaStream_t s = m->stream;
    orbx_keypoint* din = reinterpret_cast<orbx_keypoint*>(m->d_gen); orbx_keypoint* dout = din + n;
    CKM(cudaMemcpyAsync(din, kps, sizeof(orbx_keypoint) * n, cudaMemcpyHostToDevice, s));
    k_undistort<<<dim3((n + 255) / 256, 1), 256, 0, s>>>(din, nullptr, n, n, A, dout); ORBX_COUNT_LAUNCH(1);
    CKM(cudaGetLastError());
    CKM(cudaMemcpyAsync(kps_un, dout, sizeof(orbx_keypoint) * n, cudaMemcpyDeviceToHost, s));
    CKM(cudaStreamSynchronize(s));
    return ORBX_OK;
}

// The same on the keypoints an extractor holds in its result slots: d_kps_un is a DEVICE array
// [count][orbx_extractor_max_keypoints(ex)] (mvKeysUn of every frame of the batch); asynchronous on `stream`.
extern "C" int orbx_undistort_slots_device(orbx_extractor* ex, int first_slot, int count, const float* K, const float* dist, int ndist,
                                           const float* P, orbx_keypoint* d_kps_un, void* stream)
{
    if (!ex || !d_kps_un || count <= 0) return ORBX_E_INVALID;
    orbx_keypoint* dk; uint8_t* dd; int32_t* dn; int cap, slots;
    int rc = orbx_extractor_results_device(ex, &dk, &dd, &dn, nullptr, &cap, &slots);
    if (rc) return rc;
    if (first_slot < 0 || first_slot + count > slots) return ORBX_E_INVALID;
    UndistortArgs A;
    if ((rc = undistort_args(K, dist, ndist, P, &A))) return rc;
    CKM(cudaSetDevice(orbx_ex_device(ex)));
    cudaStream_t s = stream ? (cudaStream_t)stream : orbx_ex_stream(ex);
    if (dist[0] == 0.0f) {
        CKM(cudaMemcpyAsync(d_kps_un, dk + (size_t)first_slot * cap, sizeof(orbx_keypoint) * (size_t)cap * count, cudaMemcpyDeviceToDevice, s));
        return ORBX_OK;
    }
    k_undistort<<<dim3((cap + 255) / 256, count), 256, 0, s>>>(dk + (size_t)first_slot * cap, dn + first_slot, 0, cap, A, d_kps_un);
    ORBX_COUNT_LAUNCH(1);
    CKM(cudaGetLastError());
    return ORBX_OK;
}

// generic candidate matching for the host-side searches (SearchByBoW, SearchForTriangulation, Fuse, SearchBySim3)
extern "C" int orbx_match_candidates(orbx_matcher* m, const uint8_t* q, int nq, const uint8_t* t, int nt, const int32_t* offsets,
                                     const int32_t* indices, int32_t* idx, int32_t* dist)
{
    if (!m || nq < 0 || nt < 0 || (nq > 0 && (!q || !offsets || !idx || !dist))) return ORBX_E_INVALID;
    if (nq == 0) return ORBX_OK;
    const int ncand = offsets[nq];
    if (ncand < 0 || (ncand > 0 && (!indices || !t))) return ORBX_E_INVALID;
    for (int i = 0; i < nq; i++)
        if (offsets[i] < 0 || offsets[i] > offsets[i + 1]) { orbx_set_error("%s%s", "orbx_match_candidates: offsets must be non-decreasing", ""); return ORBX_E_INVALID; }
    for (int k = 0; k < ncand; k++)
        if ((unsigned)indices[k] >= (unsigned)nt) { orbx_set_error("%s%s", "orbx_match_candidates: candidate index out of range", ""); return ORBX_E_INVALID; }
    CKM(cudaSetDevice(m->p.device));
    cudaStream_t s = m->stream;
    uint8_t *dq, *dt; int32_t *doff, *dind, *dres;
    const size_t bytes = (size_t)nq * 32 + (size_t)(nt > 0 ? nt : 1) * 32 + sizeof(int32_t) * ((size_t)nq + 1 + (ncand > 0 ? ncand : 1) + 4 * (size_t)nq) + 256;
    { const int rcs = orbx_m_gen_scratch(m, bytes); if (rcs) return rcs; }
    dq = m->d_gen; dt = dq + (((size_t)nq * 32 + 63) & ~(size_t)63);
    doff = reinterpret_cast<int32_t*>(dt + (((size_t)(nt > 0 ? nt : 1) * 32 + 63) & ~(size_t)63));
    dind = doff + nq + 1; dres = dind + (ncand > 0 ? ncand : 1);
    CKM(cudaMemcpyAsync(dq, q, (size_t)nq * 32, cudaMemcpyHostToDevice, s));
    if (nt) CKM(cudaMemcpyAsync(dt, t, (size_t)nt * 32, cudaMemcpyHostToDevice, s));
    CKM(cudaMemcpyAsync(doff, offsets, sizeof(int32_t) * (nq + 1), cudaMemcpyHostToDevice, s));
    if (ncand) CKM(cudaMemcpyAsync(dind, indices, sizeof(int32_t) * ncand, cudaMemcpyHostToDevice, s));
    k_match_candidates<<<(nq + 7) / 8, 256, 0, s>>>(dq, nq, dt, doff, dind, dres, dres + 2 * nq); ORBX_COUNT_LAUNCH(1);
    CKM(cudaGetLastError());
    CKM(cudaMemcpyAsync(idx, dres, sizeof(int32_t) * 2 * nq, cudaMemcpyDeviceToHost, s));
    CKM(cudaMemcpyAsync(dist, dres + 2 * nq, sizeof(int32_t) * 2 * nq, cudaMemcpyDeviceToHost, s));
    CKM(cudaStreamSynchronize(s));
    return ORBX_OK;
}

// ORBmatcher::SearchByBoW on flat arrays (see include/orbx.h).  Host pointers, synchronous.
// validates the two feature sets + FeatureVectors, uploads them and fills BowArgs; `extra` more bytes are reserved behind
// the block (returned in *d_extra).  who = name for the error messages.
static int bow_stage(orbx_matcher* m, const char* who,
                     const orbx_keypoint* k1, const uint8_t* d1, const uint8_t* valid1, int n1,
                     const int32_t* fv1_nodes, const int32_t* fv1_start, const int32_t* fv1_feat, int nfv1,
                     const orbx_keypoint* k2, const uint8_t* d2, const uint8_t* valid2, int n2,
                     const int32_t* fv2_nodes, const int32_t* fv2_start, const int32_t* fv2_feat, int nfv2,
                     size_t extra, BowArgs* out, uint8_t** d_extra)
{
    if (!k1 || !d1 || !valid1 || !k2 || !d2 || !fv1_nodes || !fv1_start || !fv1_feat || !fv2_nodes || !fv2_start || !fv2_feat) return ORBX_E_INVALID;
    const int nf1 = fv1_start[nfv1], nf2 = fv2_start[nfv2];
    if (nf1 < 0 || nf2 < 0 || fv1_start[0] != 0 || fv2_start[0] != 0) return ORBX_E_INVALID;
    for (int i = 0; i < nfv1; i++) if (fv1_start[i] > fv1_start[i + 1] || (i && fv1_nodes[i - 1] >= fv1_nodes[i])) { orbx_set_error("%s%s", who, ": FeatureVector 1 must be sorted by node id"); return ORBX_E_INVALID; }
    for (int i = 0; i < nfv2; i++) if (fv2_start[i] > fv2_start[i + 1] || (i && fv2_nodes[i - 1] >= fv2_nodes[i])) { orbx_set_error("%s%s", who, ": FeatureVector 2 must be sorted by node id"); return ORBX_E_INVALID; }
    for (int i = 0; i < nf1; i++) if ((unsigned)fv1_feat[i] >= (unsigned)n1) { orbx_set_error("%s%s", who, ": feature index out of range"); return ORBX_E_INVALID; }
    for (int i = 0; i < nf2; i++) if ((unsigned)fv2_feat[i] >= (unsigned)n2) { orbx_set_error("%s%s", who, ": feature index out of range"); return ORBX_E_INVALID; }
    CKM(cudaSetDevice(m->p.device));
    cudaStream_t s = m->stream;
    // one scratch block: [k1][d1][valid1][k2][d2][valid2][fv tables][outputs][extra]
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_k1 = take(sizeof(orbx_keypoint) * n1), o_d1 = take((size_t)32 * n1), o_v1 = take(n1);
    const size_t o_k2 = take(sizeof(orbx_keypoint) * n2), o_d2 = take((size_t)32 * n2), o_v2 = take(n2);
    const size_t o_n1 = take(sizeof(int32_t) * nfv1), o_s1 = take(sizeof(int32_t) * (nfv1 + 1)), o_f1 = take(sizeof(int32_t) * (nf1 + 1));
    const size_t o_n2 = take(sizeof(int32_t) * nfv2), o_s2 = take(sizeof(int32_t) * (nfv2 + 1)), o_f2 = take(sizeof(int32_t) * (nf2 + 1));
    const size_t o_m = take(sizeof(int32_t) * n1), o_c = take(n2), o_b = take(n1), o_h = take(sizeof(int32_t) * (ORBX_HISTO_LENGTH + 2));
    const size_t o_x = take(extra);
    { const int rcs = orbx_m_gen_scratch(m, off); if (rcs) return rcs; }
    uint8_t* B = m->d_gen;
    // the caller's twelve arrays are pageable: they are gathered into a pinned mirror of the block (with the initial values of the
    // output tables) and travel with ONE copy instead of twelve copies and three memsets (~8 us each)
    if (o_x > m->h_gen_bytes) {
        if (m->h_gen) { cudaFreeHost(m->h_gen); m->h_gen = nullptr; m->h_gen_bytes = 0; }
        const size_t cap_bytes = o_x + o_x / 2;
        if (cudaMallocHost((void**)&m->h_gen, cap_bytes) != cudaSuccess) { orbx_set_error("%s%s", who, ": cudaMallocHost failed"); return ORBX_E_NOMEM; }
        m->h_gen_bytes = cap_bytes;
    }
    uint8_t* Hh = m->h_gen;
    memcpy(Hh + o_k1, k1, sizeof(orbx_keypoint) * n1);
    memcpy(Hh + o_d1, d1, (size_t)32 * n1);
    memcpy(Hh + o_v1, valid1, n1);
    memcpy(Hh + o_k2, k2, sizeof(orbx_keypoint) * n2);
    memcpy(Hh + o_d2, d2, (size_t)32 * n2);
    if (valid2) memcpy(Hh + o_v2, valid2, n2);
    memcpy(Hh + o_n1, fv1_nodes, sizeof(int32_t) * nfv1);
    memcpy(Hh + o_s1, fv1_start, sizeof(int32_t) * (nfv1 + 1));
    if (nf1) memcpy(Hh + o_f1, fv1_feat, sizeof(int32_t) * nf1);
    memcpy(Hh + o_n2, fv2_nodes, sizeof(int32_t) * nfv2);
    memcpy(Hh + o_s2, fv2_start, sizeof(int32_t) * (nfv2 + 1));
    if (nf2) memcpy(Hh + o_f2, fv2_feat, sizeof(int32_t) * nf2);
    memset(Hh + o_m, 0xFF, sizeof(int32_t) * n1);
    memset(Hh + o_c, 0, n2);
    memset(Hh + o_h, 0, sizeof(int32_t) * (ORBX_HISTO_LENGTH + 2));
    CKM(cudaMemcpyAsync(B, Hh, o_x, cudaMemcpyHostToDevice, s));
    BowArgs& A = *out;
    A.mode = 0;
    A.k1 = reinterpret_cast<const orbx_keypoint*>(B + o_k1); A.d1 = B + o_d1; A.valid1 = B + o_v1; A.n1 = n1;
    A.fv1_nodes = reinterpret_cast<const int32_t*>(B + o_n1); A.fv1_start = reinterpret_cast<const int32_t*>(B + o_s1);
    A.fv1_feat = reinterpret_cast<const int32_t*>(B + o_f1); A.nfv1 = nfv1;
    A.k2 = reinterpret_cast<const orbx_keypoint*>(B + o_k2); A.d2 = B + o_d2; A.valid2 = valid2 ? B + o_v2 : nullptr; A.n2 = n2;
    A.fv2_nodes = reinterpret_cast<const int32_t*>(B + o_n2); A.fv2_start = reinterpret_cast<const int32_t*>(B + o_s2);
    A.fv2_feat = reinterpret_cast<const int32_t*>(B + o_f2); A.nfv2 = nfv2;
    A.nnratio = 0.f; A.check_ori = 0;
    A.matches12 = reinterpret_cast<int32_t*>(B + o_m); A.claimed2 = B + o_c; A.bin_of = B + o_b; A.hist = reinterpret_cast<int32_t*>(B + o_h);
    if (d_extra) *d_extra = B + o_x;
    return ORBX_OK;
}

// rotation filter + result download shared by the node-based searches
static int bow_finish(orbx_matcher* m, const BowArgs& A, int32_t* matches12, int* nmatches)
{
    cudaStream_t s = m->stream;
    k_bow_finish<<<1, 256, 0, s>>>(A, A.hist + ORBX_HISTO_LENGTH + 1); ORBX_COUNT_LAUNCH(1);
    CKM(cudaGetLastError());
    // matches12 .. the histogram tail (the count) are one stretch of the block: one copy into the pinned mirror, then unpack
    const size_t o_m = reinterpret_cast<const uint8_t*>(A.matches12) - m->d_gen;
    const size_t o_e = reinterpret_cast<const uint8_t*>(A.hist + ORBX_HISTO_LENGTH + 2) - m->d_gen;
    CKM(cudaMemcpyAsync(m->h_gen + o_m, m->d_gen + o_m, o_e - o_m, cudaMemcpyDeviceToHost, s));
    CKM(cudaStreamSynchronize(s));
    memcpy(matches12, m->h_gen + o_m, sizeof(int32_t) * A.n1);
    const int nm = *reinterpret_cast<const int32_t*>(m->h_gen + (reinterpret_cast<const uint8_t*>(A.hist + ORBX_HISTO_LENGTH + 1) - m->d_gen));
    if (nmatches) *nmatches = nm;
    return ORBX_OK;
}

extern "C" int orbx_search_by_bow(orbx_matcher* m, int mode,
                                  const orbx_keypoint* k1, const uint8_t* d1, const uint8_t* valid1, int n1,
                                  const int32_t* fv1_nodes, const int32_t* fv1_start, const int32_t* fv1_feat, int nfv1,
                                  const orbx_keypoint* k2, const uint8_t* d2, const uint8_t* valid2, int n2,
                                  const int32_t* fv2_nodes, const int32_t* fv2_start, const int32_t* fv2_feat, int nfv2,
                                  float nnratio, int check_ori, int32_t* matches12, int* nmatches)
{
    if (!m || (mode != 0 && mode != 1) || n1 < 0 || n2 < 0 || n2 > 65535 || nfv1 < 0 || nfv2 < 0 || !matches12) return ORBX_E_INVALID;
    if (nmatches) *nmatches = 0;
    for (int i = 0; i < n1; i++) matches12[i] = -1;
    if (n1 == 0 || n2 == 0 || nfv1 == 0 || nfv2 == 0) return ORBX_OK;
    BowArgs A;
    int rc = bow_stage(m, "orbx_search_by_bow", k1, d1, valid1, n1, fv1_nodes, fv1_start, fv1_feat, nfv1,
                       k2, d2, valid2, n2, fv2_nodes, fv2_start, fv2_feat, nfv2, 0, &A, nullptr);
    if (rc) return rc;
    A.mode = mode; A.nnratio = nnratio; A.check_ori = check_ori;
    k_bow_match<<<(nfv1 + 7) / 8, 256, 0, m->stream>>>(A); ORBX_COUNT_LAUNCH(1);
    return bow_finish(m, A, matches12, nmatches);
}

// SearchByBoW(KeyFrame*, Frame&) on a two-camera frame (see include/orbx.h).  Host pointers, synchronous.
extern "C" int orbx_search_by_bow_rig(orbx_matcher* m,
                                      const orbx_keypoint* k1, const uint8_t* d1, const uint8_t* valid1, int n1,
                                      const int32_t* fv1_nodes, const int32_t* fv1_start, const int32_t* fv1_feat, int nfv1,
                                      const orbx_keypoint* k2, const uint8_t* d2, int n2, int n2_left,
                                      const int32_t* fv2_nodes, const int32_t* fv2_start, const int32_t* fv2_feat, int nfv2,
                                      float nnratio, int check_ori, int32_t* matches12_left, int32_t* matches12_right, int* nmatches)
{
    if (!m || n1 < 0 || n2 < 0 || n2 > 65535 || n2_left < 0 || n2_left > n2 || nfv1 < 0 || nfv2 < 0 || !matches12_left || !matches12_right) return ORBX_E_INVALID;
    if (nmatches) *nmatches = 0;
    for (int i = 0; i < n1; i++) { matches12_left[i] = -1; matches12_right[i] = -1; }
    if (n1 == 0 || n2 == 0 || nfv1 == 0 || nfv2 == 0) return ORBX_OK;
    BowArgs A;
    uint8_t* dx = nullptr;
    const size_t mbytes = (sizeof(int32_t) * (size_t)n1 + 255) & ~(size_t)255;
    int rc = bow_stage(m, "orbx_search_by_bow_rig", k1, d1, valid1, n1, fv1_nodes, fv1_start, fv1_feat, nfv1,
                       k2, d2, nullptr, n2, fv2_nodes, fv2_start, fv2_feat, nfv2, mbytes + n1, &A, &dx);
    if (rc) return rc;
    cudaStream_t s = m->stream;
    int32_t* d_m12r = reinterpret_cast<int32_t*>(dx); uint8_t* d_binr = dx + mbytes;
    CKM(cudaMemsetAsync(d_m12r, 0xFF, sizeof(int32_t) * n1, s));
    A.mode = 0; A.nnratio = nnratio; A.check_ori = check_ori;
    k_bow_match_rig<<<(nfv1 + 7) / 8, 256, 0, s>>>(A, n2_left, d_m12r, d_binr); ORBX_COUNT_LAUNCH(1);
    k_bow_finish_rig<<<1, 256, 0, s>>>(A, d_m12r, d_binr, A.hist + ORBX_HISTO_LENGTH + 1); ORBX_COUNT_LAUNCH(1);
    CKM(cudaGetLastError());
    int nm = 0;
    CKM(cudaMemcpyAsync(matches12_left, A.matches12, sizeof(int32_t) * n1, cudaMemcpyDeviceToHost, s));
    CKM(cudaMemcpyAsync(matches12_right, d_m12r, sizeof(int32_t) * n1, cudaMemcpyDeviceToHost, s));
    CKM(cudaMemcpyAsync(&nm, A.hist + ORBX_HISTO_LENGTH + 1, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CKM(cudaStreamSynchronize(s));
    if (nmatches) *nmatches = nm;
    return ORBX_OK;
}

// ORBmatcher::SearchForTriangulation on flat arrays (see include/orbx.h).  Host pointers, synchronous.
extern "C" int orbx_search_for_triangulation(orbx_matcher* m,
                                             const orbx_keypoint* k1, const uint8_t* d1, const uint8_t* free1, const uint8_t* stereo1, int n1,
                                             const int32_t* fv1_nodes, const int32_t* fv1_start, const int32_t* fv1_feat, int nfv1,
                                             const orbx_keypoint* k2, const uint8_t* d2, const uint8_t* free2, const uint8_t* stereo2, int n2,
                                             const int32_t* fv2_nodes, const int32_t* fv2_start, const int32_t* fv2_feat, int nfv2,
                                             const float* F12, float ep_x, float ep_y, const float* scale_factors2,
                                             const float* level_sigma2_2, int nlevels, int only_stereo, int coarse, int check_ori,
                                             int32_t* matches12, int* nmatches)
{
    if (!m || n1 < 0 || n2 < 0 || n2 > 65535 || nfv1 < 0 || nfv2 < 0 || !matches12 || !F12 || !scale_factors2 || !level_sigma2_2 ||
        nlevels < 1 || nlevels > ORBX_MAX_LEVELS) return ORBX_E_INVALID;
    if (nmatches) *nmatches = 0;
    for (int i = 0; i < n1; i++) matches12[i] = -1;
    if (n1 == 0 || n2 == 0 || nfv1 == 0 || nfv2 == 0) return ORBX_OK;
    if (!free2) return ORBX_E_INVALID;
    TriArgs T;
    uint8_t* dx = nullptr;
    int rc = bow_stage(m, "orbx_search_for_triangulation", k1, d1, free1, n1, fv1_nodes, fv1_start, fv1_feat, nfv1,
                       k2, d2, free2, n2, fv2_nodes, fv2_start, fv2_feat, nfv2, (size_t)n1 + n2 + 512, &T.B, &dx);
    if (rc) return rc;
    cudaStream_t s = m->stream;
    T.stereo1 = nullptr; T.stereo2 = nullptr;
    if (stereo1) { CKM(cudaMemcpyAsync(dx, stereo1, n1, cudaMemcpyHostToDevice, s)); T.stereo1 = dx; }
    uint8_t* dx2 = dx + (((size_t)n1 + 255) & ~(size_t)255);
    if (stereo2) { CKM(cudaMemcpyAsync(dx2, stereo2, n2, cudaMemcpyHostToDevice, s)); T.stereo2 = dx2; }
    for (int i = 0; i < 9; i++) T.F[i] = F12[i];
    T.ep_x = ep_x; T.ep_y = ep_y;
    for (int l = 0; l < ORBX_MAX_LEVELS; l++) { T.scale2[l] = l < nlevels ? scale_factors2[l] : 0.f; T.sigma2_2[l] = l < nlevels ? level_sigma2_2[l] : 0.f; }
    T.only_stereo = only_stereo; T.coarse = coarse;
    T.B.check_ori = check_ori;
    k_triangulation_match<<<(nfv1 + 7) / 8, 256, 0, s>>>(T); ORBX_COUNT_LAUNCH(1);
    return bow_finish(m, T.B, matches12, nmatches);
}

// MapPoint::ComputeDistinctiveDescriptors for a batch of map points (see include/orbx.h).  Host pointers, synchronous.
extern "C" int orbx_distinctive_descriptors(orbx_matcher* m, const uint8_t* desc, const int32_t* offsets, int npoints, int32_t* best)
{
    if (!m || npoints < 0 || (npoints > 0 && (!offsets || !best))) return ORBX_E_INVALID;
    if (npoints == 0) return ORBX_OK;
    if (offsets[0] != 0) return ORBX_E_INVALID;
    for (int p = 0; p < npoints; p++) if (offsets[p] > offsets[p + 1]) { orbx_set_error("%s%s", "orbx_distinctive_descriptors: offsets must be non-decreasing", ""); return ORBX_E_INVALID; }
    const int total = offsets[npoints];
    if (total > 0 && !desc) return ORBX_E_INVALID;
    CKM(cudaSetDevice(m->p.device));
    cudaStream_t s = m->stream;
    const size_t o_d = 0, o_o = ((size_t)(total > 0 ? total : 1) * 32 + 255) & ~(size_t)255;
    const size_t o_b = o_o + ((sizeof(int32_t) * ((size_t)npoints + 1) + 255) & ~(size_t)255);
    const size_t bytes = o_b + sizeof(int32_t) * (size_t)npoints;
    { const int rcs = orbx_m_gen_scratch(m, bytes); if (rcs) return rcs; }
    uint8_t* B = m->d_gen;
    if (total) CKM(cudaMemcpyAsync(B + o_d, desc, (size_t)total * 32, cudaMemcpyHostToDevice, s));
    CKM(cudaMemcpyAsync(B + o_o, offsets, sizeof(int32_t) * ((size_t)npoints + 1), cudaMemcpyHostToDevice, s));
    k_distinctive<<<(npoints + 7) / 8, 256, 0, s>>>(B + o_d, reinterpret_cast<const int32_t*>(B + o_o), npoints, reinterpret_cast<int32_t*>(B + o_b));
    ORBX_COUNT_LAUNCH(1);
    CKM(cudaGetLastError());
    CKM(cudaMemcpyAsync(best, B + o_b, sizeof(int32_t) * (size_t)npoints, cudaMemcpyDeviceToHost, s));
    CKM(cudaStreamSynchronize(s));
    return ORBX_OK;
}

