// orbx_match.cu - 256-bit Hamming matching: brute-force kNN-2, grid-windowed searches, stereo band search.
//
// Replaces (R/ = src/orb_slam3_ros/orb_slam3/):
//   ORBmatcher::DescriptorDistance                      R/src/ORBmatcher.cc:2358-2374
//   cv::BFMatcher(NORM_HAMMING).knnMatch(k=2)           R/src/Frame.cc:1127-1137 (+ server cross-agent matching)
//   Frame::AssignFeaturesToGrid / GetFeaturesInArea     R/src/Frame.cc:360-391, 628-709
//   ORBmatcher::SearchForInitialization                 R/src/ORBmatcher.cc:702-817
//   ORBmatcher::SearchByProjection (2 overloads)        R/src/ORBmatcher.cc:44-214, 1970-2186
//   ORBmatcher::ComputeThreeMaxima                      R/src/ORBmatcher.cc:2312-2353
//   Frame::ComputeStereoMatches (descriptor search)     R/src/Frame.cc:785-868
//
// Structure of the windowed searches: all (query, candidate) distances are order-free and computed in
// parallel (one warp per query, candidates in the reference's visit order: grid column ix, then row iy,
// then insertion order) into a CSR pool; the order-dependent bookkeeping of the reference (a keypoint taken
// by an earlier query is skipped / stolen back) is replayed by one warp per frame pair over that pool.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "orbx_internal.h"

#define CKM(call)                                                                         \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            orbx_set_error("%s failed: %s", #call, cudaGetErrorString(e_));               \
            return ORBX_E_CUDA;                                                           \
        }                                                                                 \
    } while (0)

#define ORBX_MAX_CHUNKS 8

namespace {

constexpr int GC = ORBX_GRID_COLS, GR = ORBX_GRID_ROWS, NCELL = GC * GR;

struct PairDesc {
    const orbx_keypoint* k1; const uint8_t* d1;       // query frame (SearchForInitialization) or unused
    const orbx_keypoint* k2; const uint8_t* d2;       // searched frame
    const float* uright2;                             // may be null
    const orbx_proj_query* q; const uint8_t* qdesc;   // queries
    int n1, n2, nq;
};

__device__ __forceinline__ int hamming256(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1)
{
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// The same distance with carry-save compression for the popc-bound brute-force kernel: three full adders (2 LOP3 each)
// fold 7 of the 8 difference words into 2 "ones" and 3 "twos" words, so a pair costs 5 POPC (the 16-lane XU pipe) instead
// of 8, at the price of 6 LOP3 on the 64-lane ALU pipe: d = popc(s3) + popc(x7) + 2 * (popc(c1) + popc(c2) + popc(c3)).
__device__ __forceinline__ unsigned xor3(unsigned a, unsigned b, unsigned c)
{
    unsigned r; asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
__device__ __forceinline__ unsigned maj3(unsigned a, unsigned b, unsigned c)
{
    unsigned r; asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
__device__ __forceinline__ int hamming256_csa(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1)
{
    const unsigned x0 = a0.x ^ b0.x, x1 = a0.y ^ b0.y, x2 = a0.z ^ b0.z, x3 = a0.w ^ b0.w;
    const unsigned x4 = a1.x ^ b1.x, x5 = a1.y ^ b1.y, x6 = a1.z ^ b1.z, x7 = a1.w ^ b1.w;
    const unsigned s1 = xor3(x0, x1, x2), c1 = maj3(x0, x1, x2);
    const unsigned s2 = xor3(x3, x4, x5), c2 = maj3(x3, x4, x5);
    const unsigned s3 = xor3(s1, s2, x6), c3 = maj3(s1, s2, x6);
    return __popc(s3) + __popc(x7) + 2 * (__popc(c1) + __popc(c2) + __popc(c3));
}

__global__ void k_hamming_pairs(const uint4* a, const uint4* b, int n, int* out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = hamming256(a[2 * i], a[2 * i + 1], b[2 * i], b[2 * i + 1]);
}

// ---------------------------------------------------------------------------------------------------
// brute-force kNN-2.  One thread = one query descriptor held in 8 registers; train descriptors stream
// through shared memory in tiles and are read by all threads as broadcasts.  Top-2 by (distance, index).
// ---------------------------------------------------------------------------------------------------
constexpr int BF_NT = 128;      // queries per CTA
constexpr int BF_TILE = 256;    // train descriptors per tile (8 KB)

struct BfArgs {
    // direct mode
    const uint8_t* q; const uint8_t* t; int nq; long long nt;
    // slot mode (q == nullptr): descriptors of result slots a[p] / b[p]
    const uint8_t* desc; const int* n; const int* a; const int* b; int cap;
    int32_t* idx; int32_t* dist;          // [pair][out_stride][2]
    int out_stride; int idx_base;
    long long chunk;                      // train descriptors per split
    int nsplit;
    int32_t* part_idx; int32_t* part_dist;   // [pair][split][out_stride][2] when nsplit > 1
};

__global__ void __launch_bounds__(BF_NT) k_bf_knn2(BfArgs A)
{
    __shared__ uint4 tile[BF_TILE * 2];
    const int p = blockIdx.z, split = blockIdx.y;
    const uint8_t* q; const uint8_t* t; int nq; long long nt;
    if (A.q) { q = A.q; t = A.t; nq = A.nq; nt = A.nt; }
    else {
        const int sa = A.a[p], sb = A.b[p];
        q = A.desc + (long long)sa * A.cap * 32; nq = A.n[sa];
        t = A.desc + (long long)sb * A.cap * 32; nt = A.n[sb];
    }
    if ((long long)blockIdx.x * BF_NT >= nq) return;
    const int qi = blockIdx.x * BF_NT + threadIdx.x;
    const bool live = qi < nq;
    uint4 q0 = make_uint4(0, 0, 0, 0), q1 = q0;
    if (live) { q0 = reinterpret_cast<const uint4*>(q)[2 * qi]; q1 = reinterpret_cast<const uint4*>(q)[2 * qi + 1]; }
    int d0 = 0x7fffffff, d1 = 0x7fffffff, i0 = -1, i1 = -1;
    const long long t_begin = (long long)split * A.chunk;
    long long t_end = t_begin + A.chunk; if (t_end > nt) t_end = nt;
    for (long long base = t_begin; base < t_end; base += BF_TILE) {
        const int cnt = (int)min((long long)BF_TILE, t_end - base);
        __syncthreads();
        for (int k = threadIdx.x; k < cnt * 2; k += BF_NT) tile[k] = __ldg(reinterpret_cast<const uint4*>(t) + base * 2 + k);
        __syncthreads();
        if (live) {
#pragma unroll 4
            for (int j = 0; j < cnt; j++) {
                const int d = hamming256_csa(q0, q1, tile[2 * j], tile[2 * j + 1]);
                if (d < d1) {
                    const int id = (int)(base + j) + A.idx_base;
                    if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = id; }
                    else { d1 = d; i1 = id; }
                }
            }
        }
    }
    if (!live) return;
    int32_t* oi; int32_t* od;
    if (A.nsplit > 1) {
        const long long o = (((long long)p * A.nsplit + split) * A.out_stride + qi) * 2;
        oi = A.part_idx + o; od = A.part_dist + o;
    } else {
        const long long o = ((long long)p * A.out_stride + qi) * 2;
        oi = A.idx + o; od = A.dist + o;
    }
    oi[0] = i0; oi[1] = i1;
    od[0] = i0 >= 0 ? d0 : -1; od[1] = i1 >= 0 ? d1 : -1;
}

// merge partial top-2 tables: parts laid out [pair][part][stride][2]; lexicographic (dist, idx)
__global__ void k_knn2_merge(const int32_t* pidx, const int32_t* pdist, int nparts, int stride, int nq_fixed,
                             const int* n, const int* a, int32_t* idx, int32_t* dist)
{
    const int p = blockIdx.y;
    const int nq = n ? n[a[p]] : nq_fixed;
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    int d0 = 0x7fffffff, d1 = 0x7fffffff, i0 = -1, i1 = -1;
    for (int s = 0; s < nparts; s++) {
        const long long o = (((long long)p * nparts + s) * stride + qi) * 2;
        for (int k = 0; k < 2; k++) {
            const int id = pidx[o + k], d = pdist[o + k];
            if (id < 0) continue;
            if (d < d0 || (d == d0 && id < i0)) { d1 = d0; i1 = i0; d0 = d; i0 = id; }
            else if (d < d1 || (d == d1 && id < i1)) { d1 = d; i1 = id; }
        }
    }
    const long long o = ((long long)p * stride + qi) * 2;
    idx[o] = i0; idx[o + 1] = i1;
    dist[o] = i0 >= 0 ? d0 : -1; dist[o + 1] = i1 >= 0 ? d1 : -1;
}

// ---------------------------------------------------------------------------------------------------
// windowed searches
// ---------------------------------------------------------------------------------------------------
struct WinBufs {
    PairDesc* pairs;              // [P]
    orbx_proj_query* q;           // [P][K]   (queries synthesised for SearchForInitialization)
    uint16_t* items;              // [P][K]   keypoint indices sorted by (cell, index)
    float4* skp;                  // [P][K]   the same order as records (x, y, octave, index): one load per candidate
    int* cell_start;              // [P][NCELL+1]
    int* q_off; int* q_cnt;       // [P][K]
    uint32_t* pool; int* pool_used;   // [P][POOL], [P]
    uint8_t* bin_of;              // [P][K]
    uint2* top2;                  // [P][K]   best / second pool entry of every query by (distance, list rank), 0xFFFFFFFF = none
    int K, POOL;
    float minX, maxX, minY, maxY, wInv, hInv;
    unsigned* err;
};

// fills PairDesc for result slots and synthesises the SearchForInitialization queries:
// level-0 keypoints of F1 search a window around vbPrevMatched = their own position, levels [0,0]
__global__ void k_setup_slot_pairs(WinBufs W, const orbx_keypoint* kps, const uint8_t* desc, const int* n,
                                   const int* a, const int* b, int cap, float window)
{
    const int p = blockIdx.x;
    const int sa = a[p], sb = b[p];
    const orbx_keypoint* k1 = kps + (long long)sa * cap;
    if (threadIdx.x == 0) {
        PairDesc d;
        d.k1 = k1; d.d1 = desc + (long long)sa * cap * 32;
        d.k2 = kps + (long long)sb * cap; d.d2 = desc + (long long)sb * cap * 32;
        d.uright2 = nullptr; d.q = W.q + (long long)p * W.K; d.qdesc = d.d1;
        d.n1 = n[sa]; d.n2 = n[sb]; d.nq = n[sa];
        W.pairs[p] = d;
    }
    const int n1 = n[sa];
    for (int i = threadIdx.x; i < n1 && i < W.K; i += blockDim.x) {
        orbx_proj_query q;
        q.u = k1[i].x; q.v = k1[i].y; q.r = window; q.minl = 0; q.maxl = 0; q.ur = 0.f; q.angle = k1[i].angle;
        q.valid = k1[i].octave > 0 ? 0 : 1;
        W.q[(long long)p * W.K + i] = q;
    }
}

// host-API variant: queries for SearchForInitialization from explicit prev_xy
__global__ void k_make_init_queries(orbx_proj_query* q, const orbx_keypoint* k1, const float* prev_xy, int n1, float window)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n1) return;
    orbx_proj_query o;
    o.u = prev_xy[2 * i]; o.v = prev_xy[2 * i + 1]; o.r = window; o.minl = 0; o.maxl = 0; o.ur = 0.f;
    o.angle = k1[i].angle; o.valid = k1[i].octave > 0 ? 0 : 1;
    q[i] = o;
}

constexpr int GRID_NT = 512;

// Frame::AssignFeaturesToGrid: keypoints sorted by (cell, index) == per-cell vectors in push_back order.
// cell = ix*GR + iy so that the cells (ix, iy0..iy1) visited by GetFeaturesInArea are one contiguous range.
__global__ void __launch_bounds__(GRID_NT) k_grid_build(WinBufs W, int npad_max)
{
    extern __shared__ uint32_t keys[];     // [npad]
    const int p = blockIdx.x;
    const PairDesc P = W.pairs[p];
    const int n = min(P.n2, W.K);
    int npad = 1; while (npad < n) npad <<= 1;
    if (npad > npad_max) npad = npad_max;
    for (int i = threadIdx.x; i < npad; i += GRID_NT) {
        uint32_t key = 0xFFFFFFFFu;
        if (i < n) {
            const orbx_keypoint kp = P.k2[i];
            const int px = (int)roundf(__fmul_rn(__fsub_rn(kp.x, W.minX), W.wInv));      // PosInGrid: round, not floor
            const int py = (int)roundf(__fmul_rn(__fsub_rn(kp.y, W.minY), W.hInv));
            if (px >= 0 && px < GC && py >= 0 && py < GR) key = ((uint32_t)(px * GR + py) << 16) | (uint32_t)i;
        }
        keys[i] = key;
    }
    __syncthreads();
    for (int k = 2; k <= npad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (npad >> 1); t += GRID_NT) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const bool up = (i & k) == 0;
                const uint32_t x = keys[i], y = keys[l];
                if ((x > y) == up) { keys[i] = y; keys[l] = x; }
            }
            __syncthreads();
        }
    uint16_t* items = W.items + (long long)p * W.K;
    float4* skp = W.skp + (long long)p * W.K;
    for (int i = threadIdx.x; i < n; i += GRID_NT) {
        const uint32_t key = keys[i];
        const int idx = (int)(key & 0xFFFF);
        items[i] = (uint16_t)idx;
        if (key != 0xFFFFFFFFu) {
            const orbx_keypoint kp = P.k2[idx];
            skp[i] = make_float4(kp.x, kp.y, __int_as_float(kp.octave), __int_as_float(idx));
        }
    }
    int* cs = W.cell_start + (long long)p * (NCELL + 1);
    for (int c = threadIdx.x; c <= NCELL; c += GRID_NT) {
        // first sorted position whose cell >= c (out-of-grid keys sort last)
        int lo = 0, hi = n;
        const uint32_t kc = (uint32_t)c << 16;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (keys[mid] < kc) lo = mid + 1; else hi = mid; }
        cs[c] = lo;
    }
}

constexpr int CAND_WARPS = 8;

// warp-wide top-2 by (dist, rank): each lane holds its local best (d0, k0, e0) and second (d1, k1, e1); ranks are unique,
// so (dist << 16 | rank) keys are unique and two REDUX.MIN + two ballots replace a 5-round shuffle tree.
__device__ __forceinline__ void warp_top2(int& d0, int& k0, uint32_t& e0, int& d1, uint32_t& e1, int& k1)
{
    const unsigned key0 = d0 == 0x7fffffff ? 0xFFFFFFFFu : (((unsigned)d0 << 16) | (unsigned)k0);
    const unsigned key1 = d1 == 0x7fffffff ? 0xFFFFFFFFu : (((unsigned)d1 << 16) | (unsigned)k1);
    const unsigned B = __reduce_min_sync(0xffffffffu, key0);
    const unsigned c2 = key0 == B ? key1 : key0;
    const unsigned S = __reduce_min_sync(0xffffffffu, c2);
    const uint32_t sel = (key0 == S) ? e0 : e1;
    const unsigned wb = __ballot_sync(0xffffffffu, key0 == B), ws = __ballot_sync(0xffffffffu, c2 == S);
    const uint32_t be = __shfl_sync(0xffffffffu, e0, __ffs(wb) - 1);
    const uint32_t se = __shfl_sync(0xffffffffu, sel, __ffs(ws) - 1);
    if (B == 0xFFFFFFFFu) { d0 = 0x7fffffff; k0 = 0x7fffffff; e0 = 0; } else { d0 = (int)(B >> 16); k0 = (int)(B & 0xFFFF); e0 = be; }
    if (S == 0xFFFFFFFFu) { d1 = 0x7fffffff; k1 = 0x7fffffff; e1 = 0; } else { d1 = (int)(S >> 16); k1 = (int)(S & 0xFFFF); e1 = se; }
}

// Frame::GetFeaturesInArea + descriptor distances, one warp per query.  Pool entry: i2 | dist << 16 | octave << 25
__global__ void __launch_bounds__(CAND_WARPS * 32) k_window_candidates(WinBufs W)
{
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.y;
    const PairDesc P = W.pairs[p];
    const int qi = blockIdx.x * CAND_WARPS + (threadIdx.x >> 5);
    if (qi >= P.nq || qi >= W.K) return;
    int* q_off = W.q_off + (long long)p * W.K + qi;
    int* q_cnt = W.q_cnt + (long long)p * W.K + qi;
    const orbx_proj_query Q = P.q[qi];
    bool any = Q.valid != 0;
    int cx0 = 0, cx1 = -1, cy0 = 0, cy1 = -1;
    if (any) {
        // :639-660, all fp32
        cx0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(Q.u, W.minX), Q.r), W.wInv)));
        cx1 = min(GC - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(Q.u, W.minX), Q.r), W.wInv)));
        cy0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(Q.v, W.minY), Q.r), W.hInv)));
        cy1 = min(GR - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(Q.v, W.minY), Q.r), W.hInv)));
        if (cx0 >= GC || cx1 < 0 || cy0 >= GR || cy1 < 0) any = false;
    }
    if (!any) { if (lane == 0) { *q_off = 0; *q_cnt = 0; } return; }
    const bool check_levels = (Q.minl > 0) || (Q.maxl >= 0);
    const int* cs = W.cell_start + (long long)p * (NCELL + 1);
    const float4* skp = W.skp + (long long)p * W.K;
    const uint4 q0 = reinterpret_cast<const uint4*>(P.qdesc)[2 * qi], q1 = reinterpret_cast<const uint4*>(P.qdesc)[2 * qi + 1];

    // The cells (ix, cy0..cy1) of one grid column are one contiguous range of the sorted records, so the window is
    // nr <= 64 ranges.  Their bounds are fetched by the lanes in parallel, prefix-summed, and the concatenation of
    // the ranges (= the reference's visit order) is walked 32 records per step: pass 0 counts the records that pass
    // the octave / window / stereo tests, pass 1 computes their distances into the reserved CSR segment.
    __shared__ int s_rs[CAND_WARPS][GC], s_rp[CAND_WARPS][GC + 1];
    int* rs = s_rs[threadIdx.x >> 5]; int* rp = s_rp[threadIdx.x >> 5];
    const int nr = cx1 - cx0 + 1;
    int T = 0;
    for (int r0 = 0; r0 < nr; r0 += 32) {
        const int r = r0 + lane;
        int st = 0, len = 0;
        if (r < nr) { st = cs[(cx0 + r) * GR + cy0]; len = cs[(cx0 + r) * GR + cy1 + 1] - st; }
        int inc = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (r < nr) { rs[r] = st; rp[r] = T + inc - len; }
        T += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) rp[nr] = T;
    __syncwarp();

    int base = 0, total = 0;
    int td0 = 0x7fffffff, tk0 = 0x7fffffff, td1 = 0x7fffffff, tk1 = 0x7fffffff;      // unfiltered top-2 by (distance, rank)
    uint32_t te0 = 0, te1 = 0;
    for (int pass = 0; pass < 2; pass++) {
        int pos = 0;
        for (int j0 = 0; j0 < T; j0 += 32) {
            const int j = j0 + lane;
            bool ok = false; int i2 = 0, oct = 0;
            if (j < T) {
                int lo = 0, hi = nr - 1;                      // last range whose prefix <= j
                while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (rp[mid] <= j) lo = mid; else hi = mid - 1; }
                const float4 rec = __ldg(skp + rs[lo] + (j - rp[lo]));
                oct = __float_as_int(rec.z); i2 = __float_as_int(rec.w);
                ok = true;
                if (check_levels) {
                    if (oct < Q.minl) ok = false;
                    if (Q.maxl >= 0 && oct > Q.maxl) ok = false;
                }
                const float dx = __fsub_rn(rec.x, Q.u), dy = __fsub_rn(rec.y, Q.v);
                if (!(fabsf(dx) < Q.r && fabsf(dy) < Q.r)) ok = false;
                // stereo gate (ORBmatcher.cc:93-98 / :2049-2055): not order dependent, applied here
                if (ok && P.uright2) {
                    const float ur2 = P.uright2[i2];
                    if (ur2 > 0 && fabsf(__fsub_rn(Q.ur, ur2)) > Q.r) ok = false;
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, ok);
            if (pass == 1 && ok) {
                const int o = pos + __popc(bal & ((1u << lane) - 1));
                const uint4 t0 = reinterpret_cast<const uint4*>(P.d2)[2 * i2], t1 = reinterpret_cast<const uint4*>(P.d2)[2 * i2 + 1];
                const int d = hamming256(q0, q1, t0, t1);
                const uint32_t e = (uint32_t)i2 | ((uint32_t)d << 16) | ((uint32_t)oct << 25);
                if (base + o < W.POOL) W.pool[(long long)p * W.POOL + base + o] = e;
                if (d < td0) { td1 = td0; tk1 = tk0; te1 = te0; td0 = d; tk0 = o; te0 = e; }      // ranks grow per lane: first minimum kept
                else if (d < td1) { td1 = d; tk1 = o; te1 = e; }
            }
            pos += __popc(bal);
        }
        if (pass == 0) {
            total = pos;
            if (lane == 0) base = total ? atomicAdd(W.pool_used + p, total) : 0;
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base + total > W.POOL) {
                if (lane == 0) { atomicOr(W.err, ORBX_DEVERR_POOL_OVERFLOW); *q_off = 0; *q_cnt = 0; }
                return;
            }
            if (lane == 0) { *q_off = base; *q_cnt = total; }
            if (total == 0) return;
        }
    }
    // the resolve kernel starts from these two and rescans the list only when one of them has been taken meanwhile
    warp_top2(td0, tk0, te0, td1, te1, tk1);
    if (lane == 0)
        W.top2[(long long)p * W.K + qi] = make_uint2(td0 == 0x7fffffff ? 0xFFFFFFFFu : te0, td1 == 0x7fffffff ? 0xFFFFFFFFu : te1);
}


__device__ __forceinline__ int rot_bin(float a1, float a2)
{
    const float factor = 1.0f / ORBX_HISTO_LENGTH;
    float rot = __fsub_rn(a1, a2);
    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
    int bin = (int)roundf(__fmul_rn(rot, factor));
    if (bin == ORBX_HISTO_LENGTH) bin = 0;
    return bin;
}

// ORBmatcher::ComputeThreeMaxima on bin counts; all lanes compute the same result
__device__ __forceinline__ void three_maxima(const int* hist, int& ind1, int& ind2, int& ind3)
{
    int max1 = 0, max2 = 0, max3 = 0;
    ind1 = ind2 = ind3 = -1;
    for (int i = 0; i < ORBX_HISTO_LENGTH; i++) {
        const int s = hist[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if ((float)max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if ((float)max3 < 0.1f * (float)max1) { ind3 = -1; }
}

// mode 2 = SearchForInitialization, 0 / 1 = SearchByProjection overloads (see orbx.h)
// out: mode 2 -> matches12 [P][K] (+ prev_xy update when prev != null); modes 0/1 -> assigned [P][K] (in/out)
__global__ void __launch_bounds__(32) k_window_resolve(WinBufs W, int mode, float nnratio, int check_ori,
                                                     int32_t* out, int32_t* nmatches, float* prev_xy)
{
    extern __shared__ int s_mem[];
    const int lane = threadIdx.x;
    const int p = blockIdx.x;
    const PairDesc P = W.pairs[p];
    const int nq = min(P.nq, W.K), n2 = min(P.n2, W.K);
    int* matchedDist = s_mem;                 // [K] (mode 2)
    int* matches21 = s_mem + W.K;             // [K] (mode 2)
    __shared__ int hist[ORBX_HISTO_LENGTH];
    int32_t* res = out + (long long)p * W.K;
    uint8_t* bin_of = W.bin_of + (long long)p * W.K;
    const uint32_t* pool = W.pool + (long long)p * W.POOL;
    const int* q_off = W.q_off + (long long)p * W.K;
    const int* q_cnt = W.q_cnt + (long long)p * W.K;
    if (lane < ORBX_HISTO_LENGTH) hist[lane] = 0;
    if (mode == 2) {
        for (int i = lane; i < n2; i += 32) { matchedDist[i] = 0x7fffffff; matches21[i] = -1; }
        for (int i = lane; i < nq; i += 32) { res[i] = -1; bin_of[i] = 0xFF; }
    } else {
        for (int i = lane; i < n2; i += 32) bin_of[i] = 0xFF;
    }
    __syncwarp();

    // Queries are visited in order; their CSR headers and their unfiltered (best, second) entries from k_window_candidates
    // are fetched 32 at a time and empty queries are skipped by ballot.  The order-dependent rule only REMOVES candidates
    // (:741 / :89-91), so when neither of the two entries has been taken meanwhile they are still the best and the second
    // of the filtered list and the query costs two shared-memory lookups; otherwise its list is rescanned.
    const uint2* top2 = W.top2 + (long long)p * W.K;
    for (int qb = 0; qb < nq; qb += 32) {
      const int my_q = qb + lane;
      const int my_cnt = my_q < nq ? q_cnt[my_q] : 0;
      const int my_off = my_q < nq ? q_off[my_q] : 0;
      const uint2 my_t2 = (my_q < nq && my_cnt > 0) ? top2[my_q] : make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
      // the two angles of the rotation histogram, fetched with the headers instead of inside the sequential rule
      float my_a1 = 0.f, my_a2 = 0.f;
      if (check_ori && mode != 1 && my_t2.x != 0xFFFFFFFFu) { my_a1 = P.q[my_q].angle; my_a2 = P.k2[my_t2.x & 0xFFFF].angle; }
      unsigned live = __ballot_sync(0xffffffffu, my_cnt > 0);
      while (live) {
        const int src = __ffs(live) - 1;
        live &= live - 1;
        const int i = qb + src;
        uint32_t e0 = __shfl_sync(0xffffffffu, my_t2.x, src), e1 = __shfl_sync(0xffffffffu, my_t2.y, src);
        int d0 = e0 == 0xFFFFFFFFu ? 0x7fffffff : (int)((e0 >> 16) & 0x1FF);
        int d1 = e1 == 0xFFFFFFFFu ? 0x7fffffff : (int)((e1 >> 16) & 0x1FF);
        const float a1 = __shfl_sync(0xffffffffu, my_a1, src);
        float a2 = __shfl_sync(0xffffffffu, my_a2, src);
        bool taken = false;
        if (mode == 2) {
            taken = (d0 != 0x7fffffff && matchedDist[e0 & 0xFFFF] <= d0) || (d1 != 0x7fffffff && matchedDist[e1 & 0xFFFF] <= d1);
        } else {
            taken = (d0 != 0x7fffffff && res[e0 & 0xFFFF] >= 0) || (mode == 1 && d1 != 0x7fffffff && res[e1 & 0xFFFF] >= 0);
        }
        if (taken) {                                     // warp-uniform: rescan the query's list with the skip rule
            const int cnt = __shfl_sync(0xffffffffu, my_cnt, src), off = __shfl_sync(0xffffffffu, my_off, src);
            int k0 = 0x7fffffff, k1 = 0x7fffffff;
            d0 = 0x7fffffff; d1 = 0x7fffffff; e0 = 0; e1 = 0;
            for (int k = lane; k < cnt; k += 32) {
                const uint32_t e = pool[off + k];
                const int i2 = e & 0xFFFF, d = (e >> 16) & 0x1FF;
                const bool skip = (mode == 2) ? (matchedDist[i2] <= d)      // ORBmatcher.cc:741
                                              : (res[i2] >= 0);             // occupied keypoint, :89-91 / :2045-2047
                if (skip) continue;
                if (d < d0) { d1 = d0; k1 = k0; e1 = e0; d0 = d; k0 = k; e0 = e; }
                else if (d < d1) { d1 = d; k1 = k; e1 = e; }
            }
            warp_top2(d0, k0, e0, d1, e1, k1);
            if (check_ori && mode != 1 && d0 != 0x7fffffff) a2 = P.k2[e0 & 0xFFFF].angle;
        }
        // every lane now holds the same (best, second); lane 0 applies the sequential rule
        if (mode == 2) {
            if (d0 <= ORBX_TH_LOW && (float)d0 < (float)d1 * nnratio) {     // :756-758 (INT_MAX second -> float)
                const int i2 = e0 & 0xFFFF;
                if (lane == 0) {
                    if (matches21[i2] >= 0) res[matches21[i2]] = -1;
                    res[i] = i2; matches21[i2] = i; matchedDist[i2] = d0;
                    if (check_ori) {
                        const int bin = rot_bin(a1, a2);
                        bin_of[i] = (uint8_t)bin; hist[bin]++;
                    }
                }
            }
        } else if (mode == 0) {
            // best starts at 256 and must be <= TH_HIGH (:2038, :2068)
            if (d0 <= ORBX_TH_HIGH) {
                const int i2 = e0 & 0xFFFF;
                if (lane == 0) {
                    res[i2] = i;
                    if (check_ori) { const int bin = rot_bin(a1, a2); bin_of[i2] = (uint8_t)bin; hist[bin]++; }
                }
            }
        } else {
            // :78-141: best/second start at 256 with level -1
            const int bd = d0 < 256 ? d0 : 256, bd2 = d1 < 256 ? d1 : 256;
            const int bl = d0 < 256 ? (int)(e0 >> 25) : -1, bl2 = d1 < 256 ? (int)(e1 >> 25) : -1;
            if (bd <= ORBX_TH_HIGH && !(bl == bl2 && (float)bd > nnratio * (float)bd2)) {
                if (lane == 0) res[e0 & 0xFFFF] = i;
            }
        }
        __syncwarp();
      }
    }
    __syncwarp();
    // rotation consistency (:786-809 / :2163-2183)
    if (check_ori && mode != 1) {
        int ind1, ind2, ind3;
        three_maxima(hist, ind1, ind2, ind3);
        const int lim = mode == 2 ? nq : n2;
#pragma unroll 4
        for (int i = lane; i < lim; i += 32) {
            const int b = bin_of[i];
            if (b != 0xFF && b != ind1 && b != ind2 && b != ind3) res[i] = -1;
        }
        __syncwarp();
    }
    int cntm = 0;
    const int lim = mode == 2 ? nq : n2;
#pragma unroll 4
    for (int i = lane; i < lim; i += 32) {
        const int m = res[i];
        if (mode == 2) {
            if (m >= 0) { cntm++; if (prev_xy) { prev_xy[((long long)p * W.K + i) * 2] = P.k2[m].x; prev_xy[((long long)p * W.K + i) * 2 + 1] = P.k2[m].y; } }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cntm += __shfl_xor_sync(0xffffffffu, cntm, o);
    if (lane == 0 && mode == 2) nmatches[p] = cntm;
}

// modes 0/1 count matches as "queries that own a keypoint they took in this call"
__global__ void k_count_new_assigned(const int32_t* before, const int32_t* after, int n, int32_t* count)
{
    __shared__ int s;
    if (threadIdx.x == 0) s = 0;
    __syncthreads();
    int c = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) c += (after[i] >= 0 && before[i] < 0) ? 1 : 0;
    atomicAdd(&s, c);
    __syncthreads();
    if (threadIdx.x == 0) *count = s;
}

// ---- Frame::ComputeStereoMatches, batched: blockIdx.y = stereo pair of the batch ----
struct StereoArgs {
    const orbx_keypoint* kL; const uint8_t* dL; const int32_t* nL; int capL;     // left results: [slot][capL]
    const orbx_keypoint* kR; const uint8_t* dR; const int32_t* nR; int capR;     // right results
    int slotL0, slotR0;                 // result slot of pair 0 (pair p uses slot*0 + p); nL/nR == nullptr: counts in nl1/nr1
    int nl1, nr1;
    int nrows;                          // level-0 rows
    float minD, maxD, mbf;
    float sf[ORBX_MAX_LEVELS];          // mvScaleFactors
    int32_t* best_idx; int32_t* best_dist;      // [pair][capL]
    float* uright; float* depth; int32_t* sad;  // [pair][ostride]
    int ostride;
};

// descriptor search (R/src/Frame.cc:785-868), one CTA per stereo pair:
//   1. vRowIndices (:798-812): every right keypoint is listed under the level-0 rows [floor(y - r), ceil(y + r)],
//      r = 2 * scale(octave), as a CSR table (row histogram in shared memory, block scan, fill into `lists`);
//   2. one warp per left keypoint walks the list of its own row (:826-865): octave gate, disparity gate, Hamming distance;
//      best = smallest distance, ties -> smallest right index (the reference visits a row's list in index order).
constexpr int STEREO_NT = 1024;
__device__ __forceinline__ void stereo_row_range(const orbx_keypoint& R, const float* sf, int nrows, int& minr, int& maxr)
{
    const float r = __fmul_rn(2.0f, sf[R.octave]);
    maxr = (int)ceilf(__fadd_rn(R.y, r)); minr = (int)floorf(__fsub_rn(R.y, r));
    if (minr < 0) minr = 0;
    if (maxr > nrows - 1) maxr = nrows - 1;
}

__global__ void __launch_bounds__(STEREO_NT) k_stereo_band(StereoArgs A, int32_t* lists, int list_cap)
{
    extern __shared__ int s_rows[];                 // [nrows + 1] starts, [nrows + 1] fill cursors
    __shared__ int s_warp[STEREO_NT / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, p = blockIdx.x;
    const int nl = A.nL ? A.nL[A.slotL0 + p] : A.nl1, nr = A.nR ? A.nR[A.slotR0 + p] : A.nr1;
    const int nrows = A.nrows;
    int* start = s_rows; int* cur = s_rows + nrows + 1;
    const orbx_keypoint* kl = A.kL + (size_t)(A.slotL0 + p) * A.capL; const uint8_t* dl = A.dL + (size_t)(A.slotL0 + p) * A.capL * 32;
    const orbx_keypoint* kr = A.kR + (size_t)(A.slotR0 + p) * A.capR; const uint8_t* dr = A.dR + (size_t)(A.slotR0 + p) * A.capR * 32;
    int32_t* list = lists + (size_t)p * list_cap;
    for (int i = tid; i <= nrows; i += STEREO_NT) start[i] = 0;
    __syncthreads();
    for (int iR = tid; iR < nr; iR += STEREO_NT) {
        int minr, maxr;
        stereo_row_range(kr[iR], A.sf, nrows, minr, maxr);
        for (int y = minr; y <= maxr; y++) atomicAdd(&start[y], 1);
    }
    __syncthreads();
    {   // exclusive scan of the row histogram: each thread owns a run of consecutive rows
        const int per = (nrows + STEREO_NT) / STEREO_NT;
        const int r0 = tid * per, r1 = min(r0 + per, nrows + 1);
        int sum = 0;
        for (int r = r0; r < r1; r++) sum += start[r];
        int inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int w = s_warp[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
            s_warp[lane] = wi - w;
        }
        __syncthreads();
        int run = s_warp[warp] + inc - sum;
        for (int r = r0; r < r1; r++) { const int c = start[r]; start[r] = run; cur[r] = run; run += c; }
    }
    __syncthreads();
    for (int iR = tid; iR < nr; iR += STEREO_NT) {
        int minr, maxr;
        stereo_row_range(kr[iR], A.sf, nrows, minr, maxr);
        for (int y = minr; y <= maxr; y++) { const int o = atomicAdd(&cur[y], 1); if (o < list_cap) list[o] = iR; }
    }
    __syncthreads();
    for (int iL = warp; iL < nl; iL += STEREO_NT / 32) {
        const orbx_keypoint L = kl[iL];
        const int row = (int)L.y;
        const float minU = __fsub_rn(L.x, A.maxD), maxU = __fsub_rn(L.x, A.minD);
        int bd = ORBX_TH_HIGH, bi = 0x7fffffff;
        if (row >= 0 && row < nrows && !(maxU < 0)) {
            const uint4 q0 = reinterpret_cast<const uint4*>(dl)[2 * iL], q1 = reinterpret_cast<const uint4*>(dl)[2 * iL + 1];
            const int k1 = min(start[row + 1], list_cap);
            for (int k = start[row] + lane; k < k1; k += 32) {
                const int iR = list[k];
                const int oct = kr[iR].octave; const float xr = kr[iR].x;
                if (oct < L.octave - 1 || oct > L.octave + 1) continue;
                if (!(xr >= minU && xr <= maxU)) continue;
                const uint4 t0 = reinterpret_cast<const uint4*>(dr)[2 * iR], t1 = reinterpret_cast<const uint4*>(dr)[2 * iR + 1];
                const int d = hamming256(q0, q1, t0, t1);
                if (d < bd || (d == bd && iR < bi && d < ORBX_TH_HIGH)) { bd = d; bi = iR; }    // strict '<' from TH_HIGH (:829)
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const int od = __shfl_xor_sync(0xffffffffu, bd, o), oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
        }
        if (lane == 0) { A.best_idx[(size_t)p * A.capL + iL] = bi == 0x7fffffff ? -1 : bi; A.best_dist[(size_t)p * A.capL + iL] = bd; }
    }
}

// rows a right keypoint can be listed under: 2 * ceil(2 * largest scale factor) + 3
static int stereo_rows_per_kp(const float* sf, int nlevels)
{
    float mx = 1.0f;
    for (int l = 0; l < nlevels; l++) if (sf[l] > mx) mx = sf[l];
    return 2 * (int)ceilf(2.0f * mx) + 3;
}

// sub-pixel refinement (R/src/Frame.cc:871-946): one warp per left keypoint.
// 11x11 patches around the keypoint (left) and around the matched right keypoint shifted by incR = -5..5, both centred on
// their own middle pixel, L1 distance per shift, parabola through the best shift and its neighbours.
__global__ void __launch_bounds__(256) k_stereo_refine(StereoArgs A, OrbxPyrView L, OrbxPyrView R)
{
    __shared__ int s_d[8][12];
    __shared__ uint8_t s_patch[8][11 * 32];
    const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5, p = blockIdx.y;
    const int iL = blockIdx.x * 8 + wq;
    const int nl = A.nL ? A.nL[A.slotL0 + p] : A.nl1;
    if (iL >= nl) return;
    const orbx_keypoint* kl = A.kL + (size_t)(A.slotL0 + p) * A.capL;
    const orbx_keypoint* kr = A.kR + (size_t)(A.slotR0 + p) * A.capR;
    float out_u = -1.0f, out_z = -1.0f; int out_s = -1;
    const int bi = A.best_idx[(size_t)p * A.capL + iL];
    const int thOrbDist = (ORBX_TH_HIGH + ORBX_TH_LOW) / 2;
    if (bi >= 0 && A.best_dist[(size_t)p * A.capL + iL] < thOrbDist) {
        const orbx_keypoint kp = kl[iL];
        const int oct = kp.octave;
        const float uL = kp.x;
        const float uR0 = kr[bi].x;
        const float sf = L.inv_scale[oct];
        const float scaleduL = roundf(__fmul_rn(kp.x, sf)), scaledvL = roundf(__fmul_rn(kp.y, sf)), scaleduR0 = roundf(__fmul_rn(uR0, sf));
        const int w = 5, Lr = 5;
        const float iniu = scaleduR0 + Lr - w, endu = scaleduR0 + Lr + w + 1;
        if (!(iniu < 0 || endu >= (float)R.w[oct])) {
            const int r0 = (int)(scaledvL - w), c0 = (int)(scaleduL - w), cr0 = (int)(scaleduR0 - w);
            const uint8_t* imL = L.lv[oct] + (long long)p * L.fstride[oct]; const int pL = L.pitch[oct];
            const uint8_t* imR = R.lv[oct] + (long long)p * R.fstride[oct]; const int pR = R.pitch[oct];
            if (lane < 11) s_d[wq][lane] = 0;
            // stage the 11x11 left patch and the 11x21 right strip (all 11 shifts) in shared memory: 32 bytes per row
            uint8_t* sp = s_patch[wq];
#pragma unroll
            for (int r = 0; r < 11; r++)
                sp[r * 32 + lane] = lane < 11 ? imL[(long long)(r0 + r) * pL + c0 + lane]
                                              : imR[(long long)(r0 + r) * pR + cr0 - Lr + (lane - 11)];
            __syncwarp();
            const int ctrL = sp[w * 32 + w];
            // 121 (shift, row) tasks of 11 pixels each: |(a - ctrL) - (b - ctrR)| = |(a + ctrR - ctrL) - b|
            for (int t = lane; t < 121; t += 32) {
                const int inc = t / 11, r = t - inc * 11;
                const int kd = (int)sp[w * 32 + 11 + inc + w] - ctrL;
                const uint8_t* a = sp + r * 32;
                const uint8_t* b = sp + r * 32 + 11 + inc;
                unsigned acc = 0;
#pragma unroll
                for (int c = 0; c < 11; c++) acc = __sad((int)a[c] + kd, (int)b[c], acc);
                atomicAdd(&s_d[wq][inc], (int)acc);
            }
            __syncwarp();
            int bestDist = 0x7fffffff, bestinc = 0;
            for (int k = 0; k < 11; k++) { const int d = s_d[wq][k]; if (d < bestDist) { bestDist = d; bestinc = k - Lr; } }   // first minimum wins (:905-909)
            if (bestinc != -Lr && bestinc != Lr) {
                const float d1 = (float)s_d[wq][Lr + bestinc - 1], d2 = (float)s_d[wq][Lr + bestinc], d3 = (float)s_d[wq][Lr + bestinc + 1];
                const float deltaR = __fdiv_rn(__fsub_rn(d1, d3), __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2))));
                if (!(deltaR < -1 || deltaR > 1)) {
                    float bestuR = __fmul_rn(L.scale[oct], __fadd_rn(__fadd_rn(scaleduR0, (float)bestinc), deltaR));
                    float disparity = __fsub_rn(uL, bestuR);
                    if (disparity >= A.minD && disparity < A.maxD) {
                        if (disparity <= 0) { disparity = 0.01f; bestuR = (float)((double)uL - 0.01); }
                        out_z = __fdiv_rn(A.mbf, disparity); out_u = bestuR; out_s = bestDist;
                    }
                }
            }
        }
    }
    if (lane == 0) {
        const size_t o = (size_t)p * A.ostride + iL;
        A.uright[o] = out_u; A.depth[o] = out_z; A.sad[o] = out_s;
    }
}

// median-based outlier cut (R/src/Frame.cc:949-962): one CTA per pair; bitonic sort of the SAD distances of the matched keypoints
__global__ void __launch_bounds__(1024) k_stereo_outliers(StereoArgs A, int npad)
{
    extern __shared__ int s_v[];
    __shared__ int s_n;
    const int p = blockIdx.x;
    const int nl = A.nL ? A.nL[A.slotL0 + p] : A.nl1;
    float* uright = A.uright + (size_t)p * A.ostride; float* depth = A.depth + (size_t)p * A.ostride;
    const int32_t* sad = A.sad + (size_t)p * A.ostride;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    int local = 0;
    for (int i = threadIdx.x; i < nl; i += 1024) local += sad[i] >= 0;
    if (local) atomicAdd(&s_n, local);
    for (int i = threadIdx.x; i < npad; i += 1024) s_v[i] = (i < nl && sad[i] >= 0) ? sad[i] : 0x7fffffff;
    __syncthreads();
    const int n = s_n;
    if (n == 0) return;
    for (int k = 2; k <= npad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (npad >> 1); t += 1024) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
                const bool up = (i & k) == 0;
                const int x = s_v[i], y = s_v[l];
                if ((x > y) == up) { s_v[i] = y; s_v[l] = x; }
            }
            __syncthreads();
        }
    const float median = (float)s_v[n / 2];
    const float thDist = __fmul_rn(1.5f * 1.4f, median);
    for (int i = threadIdx.x; i < nl; i += 1024)
        if (sad[i] >= 0 && !((float)sad[i] < thDist)) { uright[i] = -1.0f; depth[i] = -1.0f; }
}

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

// register-only throughput probes
__global__ void k_popc_probe(unsigned seed, int iters, unsigned* sink)
{
    unsigned a0 = seed + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, a4 = a0 * 11, a5 = a0 * 13, a6 = a0 * 17, a7 = a0 * 19;
    for (int i = 0; i < iters; i++) {
        a0 = __popc(a0) + a1; a1 = __popc(a1) + a2; a2 = __popc(a2) + a3; a3 = __popc(a3) + a4;
        a4 = __popc(a4) + a5; a5 = __popc(a5) + a6; a6 = __popc(a6) + a7; a7 = __popc(a7) + a0;
    }
    if ((a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7) == 0x12345678u) *sink = a0;
}
__global__ void k_lop3_probe(unsigned seed, int iters, unsigned* sink)
{
    unsigned a0 = seed + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, a4 = a0 * 11, a5 = a0 * 13, a6 = a0 * 17, a7 = a0 * 19;
    for (int i = 0; i < iters; i++) {
        a0 = (a0 & a1) ^ a2; a1 = (a1 & a2) ^ a3; a2 = (a2 & a3) ^ a4; a3 = (a3 & a4) ^ a5;
        a4 = (a4 & a5) ^ a6; a5 = (a5 & a6) ^ a7; a6 = (a6 & a7) ^ a0; a7 = (a7 & a0) ^ a1;
    }
    if ((a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7) == 0x12345678u) *sink = a0;
}

}  // namespace

// ===================================================================================================
struct orbx_matcher {
    orbx_matcher_params p;
    int K, P, POOL;
    cudaStream_t stream;
    WinBufs W;
    // device staging for the host-pointer APIs (two frames + queries)
    orbx_keypoint* d_k1; orbx_keypoint* d_k2; uint8_t* d_d1; uint8_t* d_d2; uint8_t* d_qdesc; float* d_uright;
    float* d_prev; int32_t* d_out; int32_t* d_out2; int32_t* d_nm; float* d_sf;
    int32_t* d_knn_idx; int32_t* d_knn_dist;
    int32_t* d_part_idx; int32_t* d_part_dist; size_t part_elems;
    uint8_t* d_bfq; uint8_t* d_bft; size_t bfq_bytes, bft_bytes;
    unsigned* h_err;
    int32_t* d_pair_a; int32_t* d_pair_b;
    uint8_t* d_gen; size_t gen_bytes;
    uint8_t* d_st; size_t st_bytes;          // stereo scratch
    int32_t* h_mono2; int mono2_cap;         // pinned monoIndex landing zone of the stereo pipeline (2 x batch)
    cudaStream_t s_h2d, s_d2h, s_match; cudaEvent_t ev[2 * ORBX_MAX_CHUNKS]; cudaEvent_t ev_ext[ORBX_MAX_CHUNKS]; cudaEvent_t ev_start;
    cudaEvent_t ev_r[2 * ORBX_MAX_CHUNKS];      // right camera of the stereo pipeline: [c] copy done, [MAX + c] extraction done
    std::vector<void*> allocs;
};

static int m_alloc(orbx_matcher* m, void** p, size_t bytes)
{
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) { orbx_set_error("%s: %s", "cudaMalloc", cudaGetErrorString(e)); return ORBX_E_NOMEM; }
    m->allocs.push_back(*p);
    return ORBX_OK;
}

extern "C" int orbx_matcher_create(const orbx_matcher_params* p, orbx_matcher** out)
{
    if (!p || !out || p->max_keypoints < 1 || p->max_keypoints > 24000 || p->max_batch < 1) {
        orbx_set_error("%s%s", "orbx_matcher_create: invalid parameters (max_keypoints must be 1..24000)", "");
        return ORBX_E_INVALID;
    }
    int ndev = 0;
    CKM(cudaGetDeviceCount(&ndev));
    if (p->device < 0 || p->device >= ndev) { orbx_set_error("%s%s", "no such CUDA device", ""); return ORBX_E_CUDA; }
    CKM(cudaSetDevice(p->device));
    orbx_matcher* m = new orbx_matcher();
    m->p = *p;
    m->K = p->max_keypoints; m->P = p->max_batch;
    m->POOL = p->max_candidates > 0 ? p->max_candidates : 131072;
    CKM(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
    WinBufs& W = m->W;
    memset(&W, 0, sizeof(W));
    W.K = m->K; W.POOL = m->POOL;
    const size_t K = m->K, P = m->P;
    int rc;
#define MA(ptr, bytes) if ((rc = m_alloc(m, (void**)&(ptr), (bytes)))) return rc
    MA(W.pairs, sizeof(PairDesc) * P);
    MA(W.q, sizeof(orbx_proj_query) * K * P);
    MA(W.items, sizeof(uint16_t) * K * P);
    MA(W.skp, sizeof(float4) * K * P);
    MA(W.cell_start, sizeof(int) * (NCELL + 1) * P);
    MA(W.q_off, sizeof(int) * K * P);
    MA(W.q_cnt, sizeof(int) * K * P);
    MA(W.pool, sizeof(uint32_t) * (size_t)m->POOL * P);
    MA(W.pool_used, sizeof(int) * P);
    MA(W.bin_of, K * P);
    MA(W.top2, sizeof(uint2) * K * P);
    MA(W.err, sizeof(unsigned));
    MA(m->d_k1, sizeof(orbx_keypoint) * K); MA(m->d_k2, sizeof(orbx_keypoint) * K);
    MA(m->d_d1, 32 * K); MA(m->d_d2, 32 * K); MA(m->d_qdesc, 32 * K); MA(m->d_uright, sizeof(float) * K);
    MA(m->d_prev, sizeof(float) * 2 * K * P); MA(m->d_out, sizeof(int32_t) * K * P); MA(m->d_out2, sizeof(int32_t) * K);
    MA(m->d_nm, sizeof(int32_t) * P); MA(m->d_sf, sizeof(float) * ORBX_MAX_LEVELS);
    MA(m->d_knn_idx, sizeof(int32_t) * 2 * K); MA(m->d_knn_dist, sizeof(int32_t) * 2 * K);
#undef MA
    m->d_part_idx = m->d_part_dist = nullptr; m->part_elems = 0;
    m->d_bfq = m->d_bft = nullptr; m->bfq_bytes = m->bft_bytes = 0;
    m->d_pair_a = m->d_pair_b = nullptr;
    m->d_gen = nullptr; m->gen_bytes = 0;
    m->d_st = nullptr; m->st_bytes = 0;
    m->h_mono2 = nullptr; m->mono2_cap = 0;
    m->s_h2d = m->s_d2h = nullptr;
    CKM(cudaMemset(W.err, 0, sizeof(unsigned)));
    CKM(cudaMallocHost((void**)&m->h_err, sizeof(unsigned)));
    if (2 * K * sizeof(int) > 48 * 1024)
        CKM(cudaFuncSetAttribute(k_window_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * K * sizeof(int))));
    {
        size_t npad = 1; while (npad < K) npad <<= 1;
        if (npad * sizeof(uint32_t) > 48 * 1024)
            CKM(cudaFuncSetAttribute(k_grid_build, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(npad * sizeof(uint32_t))));
    }
    *out = m;
    return ORBX_OK;
}

extern "C" void orbx_matcher_destroy(orbx_matcher* m)
{
    if (!m) return;
    cudaSetDevice(m->p.device);
    cudaStreamSynchronize(m->stream);
    for (void* p : m->allocs) cudaFree(p);
    if (m->d_part_idx) cudaFree(m->d_part_idx);
    if (m->d_part_dist) cudaFree(m->d_part_dist);
    if (m->d_bfq) cudaFree(m->d_bfq);
    if (m->d_bft) cudaFree(m->d_bft);
    if (m->d_pair_a) cudaFree(m->d_pair_a);
    if (m->d_gen) cudaFree(m->d_gen);
    if (m->d_st) cudaFree(m->d_st);
    if (m->h_mono2) cudaFreeHost(m->h_mono2);
    if (m->s_h2d) {
        cudaStreamDestroy(m->s_h2d); cudaStreamDestroy(m->s_d2h); cudaStreamDestroy(m->s_match);
        for (int i = 0; i < ORBX_MAX_CHUNKS; i++) cudaEventDestroy(m->ev_ext[i]);
        for (int i = 0; i < 2 * ORBX_MAX_CHUNKS; i++) { cudaEventDestroy(m->ev[i]); cudaEventDestroy(m->ev_r[i]); }
        cudaEventDestroy(m->ev_start);
    }
    cudaFreeHost(m->h_err);
    cudaStreamDestroy(m->stream);
    delete m;
}

static int m_check_err(orbx_matcher* m, cudaStream_t s)
{
    CKM(cudaMemcpyAsync(m->h_err, m->W.err, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
    CKM(cudaStreamSynchronize(s));
    if (*m->h_err) {
        char buf[32]; snprintf(buf, sizeof(buf), "0x%x", *m->h_err);
        orbx_set_error("matcher device capacity error flags %s%s", buf, " (raise max_candidates)");
        cudaMemsetAsync(m->W.err, 0, sizeof(unsigned), s);
        return ORBX_E_CAPACITY;
    }
    return ORBX_OK;
}

extern "C" int orbx_matcher_sync(orbx_matcher* m, void* stream)
{
    if (!m) return ORBX_E_INVALID;
    return m_check_err(m, stream ? (cudaStream_t)stream : m->stream);
}

extern "C" int orbx_hamming_pairs(orbx_matcher* m, const uint8_t* a, const uint8_t* b, int n, int32_t* out)
{
    if (!m || n < 0 || (n > 0 && (!a || !b || !out))) return ORBX_E_INVALID;
    if (n == 0) return ORBX_OK;
    CKM(cudaSetDevice(m->p.device));
    for (int done = 0; done < n; done += m->K) {
        const int c = n - done < m->K ? n - done : m->K;
        CKM(cudaMemcpyAsync(m->d_d1, a + (size_t)done * 32, (size_t)c * 32, cudaMemcpyHostToDevice, m->stream));
        CKM(cudaMemcpyAsync(m->d_d2, b + (size_t)done * 32, (size_t)c * 32, cudaMemcpyHostToDevice, m->stream));
        k_hamming_pairs<<<(c + 255) / 256, 256, 0, m->stream>>>((const uint4*)m->d_d1, (const uint4*)m->d_d2, c, m->d_out); ORBX_COUNT_LAUNCH(1);
        CKM(cudaMemcpyAsync(out + done, m->d_out, sizeof(int32_t) * c, cudaMemcpyDeviceToHost, m->stream));
        CKM(cudaStreamSynchronize(m->stream));
    }
    return ORBX_OK;
}

static int ensure_parts(orbx_matcher* m, size_t elems)
{
    if (elems <= m->part_elems) return ORBX_OK;
    if (m->d_part_idx) cudaFree(m->d_part_idx);
    if (m->d_part_dist) cudaFree(m->d_part_dist);
    CKM(cudaMalloc((void**)&m->d_part_idx, sizeof(int32_t) * elems));
    CKM(cudaMalloc((void**)&m->d_part_dist, sizeof(int32_t) * elems));
    m->part_elems = elems;
    return ORBX_OK;
}

static int bf_launch(orbx_matcher* m, BfArgs A, int npairs, int nq_max, long long nt_max, cudaStream_t s)
{
    // split the train set so that a small query set still fills the 148 SMs
    const int qblocks = (nq_max + BF_NT - 1) / BF_NT;
    int nsplit = 1;
    const long long tiles = (nt_max + BF_TILE - 1) / BF_TILE;
    if (tiles > 0) {
        const int want = 148 * 8;
        while ((long long)qblocks * npairs * nsplit < want && nsplit * 2 <= tiles && nsplit < 1024) nsplit *= 2;
    }
    long long chunk = (nt_max + nsplit - 1) / nsplit;
    chunk = (chunk + BF_TILE - 1) / BF_TILE * BF_TILE;
    if (chunk < BF_TILE) chunk = BF_TILE;
    A.chunk = chunk; A.nsplit = nsplit;
    if (nsplit > 1) {
        int rc = ensure_parts(m, (size_t)npairs * nsplit * A.out_stride * 2);
        if (rc) return rc;
        A.part_idx = m->d_part_idx; A.part_dist = m->d_part_dist;
    }
    if (qblocks == 0 || npairs == 0) return ORBX_OK;
    dim3 grid(qblocks, nsplit, npairs);
    k_bf_knn2<<<grid, BF_NT, 0, s>>>(A); ORBX_COUNT_LAUNCH(1);
    if (nsplit > 1) {
        dim3 mg((nq_max + 127) / 128, npairs);
        k_knn2_merge<<<mg, 128, 0, s>>>(A.part_idx, A.part_dist, nsplit, A.out_stride, A.nq, A.q ? nullptr : A.n, A.a, A.idx, A.dist); ORBX_COUNT_LAUNCH(1);
    }
    CKM(cudaGetLastError());
    return ORBX_OK;
}

extern "C" int orbx_bf_knn2_device(orbx_matcher* m, const uint8_t* d_q, int nq, const uint8_t* d_t, long long nt,
                                   int32_t* d_idx, int32_t* d_dist, int idx_base, void* stream)
{
    if (!m || nq < 0 || nt < 0 || !d_idx || !d_dist) return ORBX_E_INVALID;
    if (nq == 0) return ORBX_OK;
    CKM(cudaSetDevice(m->p.device));
    cudaStream_t s = stream ? (cudaStream_t)stream : m->stream;
    BfArgs A{};
    A.q = d_q; A.t = d_t; A.nq = nq; A.nt = nt; A.idx = d_idx; A.dist = d_dist; A.out_stride = nq; A.idx_base = idx_base;
    return bf_launch(m, A, 1, nq, nt, s);
}

extern "C" int orbx_bf_knn2(orbx_matcher* m, const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx, int32_t* dist)
{
    if (!m || nq < 0 || nt < 0 || (nq > 0 && (!q || !idx || !dist)) || (nt > 0 && !t)) return ORBX_E_INVALID;
    if (nq == 0) return ORBX_OK;
    CKM(cudaSetDevice(m->p.device));
    if ((size_t)nq * 32 > m->bfq_bytes) { if (m->d_bfq) cudaFree(m->d_bfq); m->bfq_bytes = (size_t)nq * 32; CKM(cudaMalloc((void**)&m->d_bfq, m->bfq_bytes + (size_t)nq * 16)); }
    if ((size_t)nt * 32 > m->bft_bytes) { if (m->d_bft) cudaFree(m->d_bft); m->bft_bytes = (size_t)nt * 32; CKM(cudaMalloc((void**)&m->d_bft, m->bft_bytes + 32)); }
    // result tables live behind the query copy
    int32_t* d_idx = reinterpret_cast<int32_t*>(m->d_bfq + (size_t)nq * 32);
    int32_t* d_dist = d_idx + 2 * nq;
    CKM(cudaMemcpyAsync(m->d_bfq, q, (size_t)nq * 32, cudaMemcpyHostToDevice, m->stream));
    if (nt) CKM(cudaMemcpyAsync(m->d_bft, t, (size_t)nt * 32, cudaMemcpyHostToDevice, m->stream));
    int rc = orbx_bf_knn2_device(m, m->d_bfq, nq, m->d_bft, nt, d_idx, d_dist, 0, m->stream);
    if (rc) return rc;
    CKM(cudaMemcpyAsync(idx, d_idx, sizeof(int32_t) * 2 * nq, cudaMemcpyDeviceToHost, m->stream));
    CKM(cudaMemcpyAsync(dist, d_dist, sizeof(int32_t) * 2 * nq, cudaMemcpyDeviceToHost, m->stream));
    CKM(cudaStreamSynchronize(m->stream));
    return ORBX_OK;
}

extern "C" int orbx_knn2_merge_device(orbx_matcher* m, const int32_t* d_idx_parts, const int32_t* d_dist_parts, int nparts,
                                      int nq, int32_t* d_idx, int32_t* d_dist, void* stream)
{
    if (!m || nparts < 1 || nq < 0) return ORBX_E_INVALID;
    if (nq == 0) return ORBX_OK;
    CKM(cudaSetDevice(m->p.device));
    cudaStream_t s = stream ? (cudaStream_t)stream : m->stream;
    dim3 mg((nq + 127) / 128, 1);
    k_knn2_merge<<<mg, 128, 0, s>>>(d_idx_parts, d_dist_parts, nparts, nq, nq, nullptr, nullptr, d_idx, d_dist); ORBX_COUNT_LAUNCH(1);
    CKM(cudaGetLastError());
    return ORBX_OK;
}

static void set_bounds(orbx_matcher* m, const float bounds[4])
{
    WinBufs& W = m->W;
    W.minX = bounds[0]; W.maxX = bounds[1]; W.minY = bounds[2]; W.maxY = bounds[3];
    W.wInv = (float)GC / (W.maxX - W.minX);      // R/src/Frame.cc:318-319
    W.hInv = (float)GR / (W.maxY - W.minY);
}

// the matcher's per-pair scratch starting at pair `pb` (chunks of one batch may be matched concurrently)
static WinBufs shifted_pairs(const orbx_matcher* m, int pb)
{
    WinBufs W = m->W;
    const long long K = m->K;
    W.pairs += pb; W.q += pb * K; W.items += pb * K; W.skp += pb * K; W.cell_start += (long long)pb * (NCELL + 1);
    W.q_off += pb * K; W.q_cnt += pb * K; W.pool += (long long)pb * m->POOL; W.pool_used += pb; W.bin_of += pb * K; W.top2 += pb * K;
    return W;
}

static int run_window(orbx_matcher* m, const WinBufs& W, int npairs, int nq_max, int mode, float nnratio, int check_ori,
                      int32_t* d_out, int32_t* d_nm, float* d_prev, cudaStream_t s)
{
    int npad = 1; while (npad < m->K) npad <<= 1;
    CKM(cudaMemsetAsync(W.pool_used, 0, sizeof(int) * npairs, s));
    k_grid_build<<<npairs, GRID_NT, sizeof(uint32_t) * npad, s>>>(W, npad); ORBX_COUNT_LAUNCH(1);
    dim3 cg((nq_max + CAND_WARPS - 1) / CAND_WARPS, npairs);
    if (nq_max > 0) { k_window_candidates<<<cg, CAND_WARPS * 32, 0, s>>>(W); ORBX_COUNT_LAUNCH(1); }
    k_window_resolve<<<npairs, 32, 2 * m->K * sizeof(int), s>>>(W, mode, nnratio, check_ori, d_out, d_nm, d_prev); ORBX_COUNT_LAUNCH(1);
    CKM(cudaGetLastError());
    return ORBX_OK;
}

extern "C" int orbx_search_for_initialization(orbx_matcher* m, const orbx_keypoint* k1, const uint8_t* d1, int n1,
                                              const orbx_keypoint* k2, const uint8_t* d2, int n2, const float bounds[4],
                                              float* prev_xy, int32_t* matches12, int window, float nnratio, int check_ori,
                                              int* nmatches)
{
    if (!m || n1 < 0 || n2 < 0 || n1 > m->K || n2 > m->K || !bounds || (n1 > 0 && (!k1 || !d1 || !prev_xy || !matches12)) ||
        (n2 > 0 && (!k2 || !d2))) {
        orbx_set_error("%s%s", "orbx_search_for_initialization: invalid arguments / more keypoints than max_keypoints", "");
        return ORBX_E_INVALID;
    }
    if (nmatches) *nmatches = 0;
    if (n1 == 0) return ORBX_OK;
    CKM(cudaSetDevice(m->p.device));
    cudaStream_t s = m->stream;
    set_bounds(m, bounds);
    CKM(cudaMemcpyAsync(m->d_k1, k1, sizeof(orbx_keypoint) * n1, cudaMemcpyHostToDevice, s));
    CKM(cudaMemcpyAsync(m->d_d1, d1, (size_t)32 * n1, cudaMemcpyHostToDevice, s));
    if (n2) {
        CKM(cudaMemcpyAsync(m->d_k2, k2, sizeof(orbx_keypoint) * n2, cudaMemcpyHostToDevice, s));
        CKM(cudaMemcpyAsync(m->d_d2, d2, (size_t)32 * n2, cudaMemcpyHostToDevice, s));
    }
    CKM(cudaMemcpyAsync(m->d_prev, prev_xy, sizeof(float) * 2 * n1, cudaMemcpyHostToDevice, s));
    PairDesc pd{};
    pd.k1 = m->d_k1; pd.d1 = m->d_d1; pd.k2 = m->d_k2; pd.d2 = m->d_d2; pd.uright2 = nullptr;
    pd.q = m->W.q; pd.qdesc = m->d_d1; pd.n1 = n1; pd.n2 = n2; pd.nq = n1;
    CKM(cudaMemcpyAsync(m->W.pairs, &pd, sizeof(pd), cudaMemcpyHostToDevice, s));
    k_make_init_queries<<<(n1 + 255) / 256, 256, 0, s>>>(m->W.q, m->d_k1, m->d_prev, n1, (float)window); ORBX_COUNT_LAUNCH(1);
    int rc = run_window(m, m->W, 1, n1, 2, nnratio, check_ori, m->d_out, m->d_nm, m->d_prev, s);
    if (rc) return rc;
    int nm = 0;
    CKM(cudaMemcpyAsync(matches12, m->d_out, sizeof(int32_t) * n1, cudaMemcpyDeviceToHost, s));
    CKM(cudaMemcpyAsync(prev_xy, m->d_prev, sizeof(float) * 2 * n1, cudaMemcpyDeviceToHost, s));
    CKM(cudaMemcpyAsync(&nm, m->d_nm, sizeof(int), cudaMemcpyDeviceToHost, s));
    rc = m_check_err(m, s);
    if (rc) return rc;
    if (nmatches) *nmatches = nm;
    return ORBX_OK;
}

extern "C" int orbx_search_by_projection(orbx_matcher* m, int mode, const orbx_proj_query* q, const uint8_t* qdesc, int nq,
                                         const orbx_keypoint* k2, const uint8_t* d2, const float* uright2, int n2,
                                         const float bounds[4], int32_t* assigned, float nnratio, int check_ori, int* nmatches)
{
    if (!m || (mode != 0 && mode != 1) || nq < 0 || n2 < 0 || nq > m->K || n2 > m->K || !bounds ||
        (nq > 0 && (!q || !qdesc)) || (n2 > 0 && (!k2 || !d2 || !assigned))) {
        orbx_set_error("%s%s", "orbx_search_by_projection: invalid arguments / more keypoints than max_keypoints", "");
        return ORBX_E_INVALID;
    }
    if (nmatches) *nmatches = 0;
    if (nq == 0 || n2 == 0) return ORBX_OK;
    CKM(cudaSetDevice(m->p.device));
    cudaStream_t s = m->stream;
    set_bounds(m, bounds);
    CKM(cudaMemcpyAsync(m->W.q, q, sizeof(orbx_proj_query) * nq, cudaMemcpyHostToDevice, s));
    CKM(cudaMemcpyAsync(m->d_qdesc, qdesc, (size_t)32 * nq, cudaMemcpyHostToDevice, s));
    CKM(cudaMemcpyAsync(m->d_k2, k2, sizeof(orbx_keypoint) * n2, cudaMemcpyHostToDevice, s));
    CKM(cudaMemcpyAsync(m->d_d2, d2, (size_t)32 * n2, cudaMemcpyHostToDevice, s));
    if (uright2) CKM(cudaMemcpyAsync(m->d_uright, uright2, sizeof(float) * n2, cudaMemcpyHostToDevice, s));
    CKM(cudaMemcpyAsync(m->d_out, assigned, sizeof(int32_t) * n2, cudaMemcpyHostToDevice, s));
    CKM(cudaMemcpyAsync(m->d_out2, assigned, sizeof(int32_t) * n2, cudaMemcpyHostToDevice, s));
    PairDesc pd{};
    pd.k1 = nullptr; pd.d1 = nullptr; pd.k2 = m->d_k2; pd.d2 = m->d_d2; pd.uright2 = uright2 ? m->d_uright : nullptr;
    pd.q = m->W.q; pd.qdesc = m->d_qdesc; pd.n1 = nq; pd.n2 = n2; pd.nq = nq;
    CKM(cudaMemcpyAsync(m->W.pairs, &pd, sizeof(pd), cudaMemcpyHostToDevice, s));
    int rc = run_window(m, m->W, 1, nq, mode, nnratio, check_ori, m->d_out, m->d_nm, nullptr, s);
    if (rc) return rc;
    k_count_new_assigned<<<1, 256, 0, s>>>(m->d_out2, m->d_out, n2, m->d_nm); ORBX_COUNT_LAUNCH(1);
    int nm = 0;
    CKM(cudaMemcpyAsync(assigned, m->d_out, sizeof(int32_t) * n2, cudaMemcpyDeviceToHost, s));
    CKM(cudaMemcpyAsync(&nm, m->d_nm, sizeof(int), cudaMemcpyDeviceToHost, s));
    rc = m_check_err(m, s);
    if (rc) return rc;
    if (nmatches) *nmatches = nm;
    return ORBX_OK;
}

static int match_slots_impl(orbx_matcher* m, orbx_extractor* ex, const int32_t* a, const int32_t* b, int npairs, int pair_base,
                            const float bounds[4], int window, float nnratio, int check_ori,
                            int32_t* d_matches12, int32_t* d_nmatches, int32_t* d_knn_idx, int32_t* d_knn_dist, cudaStream_t s)
{
    orbx_keypoint* kps; uint8_t* desc; int32_t* n; int32_t* mono; int cap, slots;
    int rc = orbx_extractor_results_device(ex, &kps, &desc, &n, &mono, &cap, &slots);
    if (rc) return rc;
    if (cap > m->K) { orbx_set_error("%s%s", "matcher max_keypoints smaller than the extractor's result capacity", ""); return ORBX_E_INVALID; }
    CKM(cudaSetDevice(m->p.device));
    set_bounds(m, bounds);
    const WinBufs W = shifted_pairs(m, pair_base);
    // NOTE: outputs use row stride K (= matcher max_keypoints)
    k_setup_slot_pairs<<<npairs, 256, 0, s>>>(W, kps, desc, n, a, b, cap, (float)window); ORBX_COUNT_LAUNCH(1);
    rc = run_window(m, W, npairs, cap, 2, nnratio, check_ori, d_matches12, d_nmatches, nullptr, s);
    if (rc) return rc;
    if (d_knn_idx && d_knn_dist) {
        BfArgs A{};
        A.q = nullptr; A.desc = desc; A.n = n; A.a = a; A.b = b; A.cap = cap;
        A.idx = d_knn_idx; A.dist = d_knn_dist; A.out_stride = m->K; A.idx_base = 0;
        rc = bf_launch(m, A, npairs, cap, cap, s);
        if (rc) return rc;
    }
    return ORBX_OK;
}

extern "C" int orbx_match_slots_device(orbx_matcher* m, orbx_extractor* ex, const int32_t* a, const int32_t* b, int npairs,
                                       const float bounds[4], int window, float nnratio, int check_ori,
                                       int32_t* d_matches12, int32_t* d_nmatches, int32_t* d_knn_idx, int32_t* d_knn_dist,
                                       void* stream)
{
    if (!m || !ex || !a || !b || npairs < 1 || npairs > m->P || !bounds || !d_matches12 || !d_nmatches) return ORBX_E_INVALID;
    return match_slots_impl(m, ex, a, b, npairs, 0, bounds, window, nnratio, check_ori, d_matches12, d_nmatches, d_knn_idx, d_knn_dist,
                            stream ? (cudaStream_t)stream : m->stream);
}

static int ensure_pipeline(orbx_matcher* m)
{
    if (!m->s_h2d) {
        CKM(cudaStreamCreateWithFlags(&m->s_h2d, cudaStreamNonBlocking));
        CKM(cudaStreamCreateWithFlags(&m->s_d2h, cudaStreamNonBlocking));
        CKM(cudaStreamCreateWithFlags(&m->s_match, cudaStreamNonBlocking));
        for (int i = 0; i < 2 * ORBX_MAX_CHUNKS; i++) CKM(cudaEventCreateWithFlags(&m->ev[i], cudaEventDisableTiming));
        for (int i = 0; i < ORBX_MAX_CHUNKS; i++) CKM(cudaEventCreateWithFlags(&m->ev_ext[i], cudaEventDisableTiming));
        for (int i = 0; i < 2 * ORBX_MAX_CHUNKS; i++) CKM(cudaEventCreateWithFlags(&m->ev_r[i], cudaEventDisableTiming));
        CKM(cudaEventCreateWithFlags(&m->ev_start, cudaEventDisableTiming));
    }
    if (!m->d_pair_a) {
        std::vector<int32_t> a(m->P), b(m->P);
        for (int i = 0; i < m->P; i++) { a[i] = i; b[i] = i + 1; }
        CKM(cudaMalloc((void**)&m->d_pair_a, sizeof(int32_t) * m->P * 2));
        m->d_pair_b = m->d_pair_a + m->P;
        CKM(cudaMemcpy(m->d_pair_a, a.data(), sizeof(int32_t) * m->P, cudaMemcpyHostToDevice));
        CKM(cudaMemcpy(m->d_pair_b, b.data(), sizeof(int32_t) * m->P, cudaMemcpyHostToDevice));
    }
    return ORBX_OK;
}

// One call = one tracking step over a batch of frames: ORBextractor::operator() on every frame (result slots
// 1..batch), SearchForInitialization (+ optional BF kNN-2) of every frame against its predecessor (slot i-1 -> i; slot 0
// holds the last frame of the previous call).  The batch is cut into chunks that flow through up to four streams
// (H2D | extraction kernels | matcher kernels | D2H): copies of chunk c+1 / c-1 and the latency-bound matcher kernels of
// chunk c-1 overlap the extraction of chunk c.  host == true: imgs/outputs are host buffers; else imgs is a device pointer
// and nothing is copied back (results stay in the slots / the caller's device arrays).
static int extract_match_pipeline(orbx_extractor* ex, orbx_matcher* m, bool host, const uint8_t* imgs, int batch, int width,
                                  int height, int stride, size_t frame_stride, int lap0, int lap1,
                                  const float bounds[4], int window, float nnratio, int check_ori,
                                  orbx_keypoint* kps, uint8_t* desc, int cap, int32_t* n, int32_t* mono_index,
                                  int32_t* matches12, int32_t* nmatches,
                                  int32_t* d_matches12, int32_t* d_nmatches, int32_t* d_knn_idx, int32_t* d_knn_dist,
                                  cudaStream_t s)
{
    int rc = orbx_ex_configure(ex, width, height);
    if (rc) return rc;
    CKM(cudaSetDevice(m->p.device));
    if ((rc = ensure_pipeline(m))) return rc;
    const bool direct = host && orbx_ex_can_fetch_direct(ex, kps, desc, cap, n, mono_index);
    // host path: chunks hide the PCIe copies under the kernels.  The call is H2D-bound in the middle, so what is left is
    // the fill (first chunk's copy) and the drain (last chunk's kernels + D2H): the first and last chunk are small
    // (1/16 of the batch each), the rest is split evenly.  Device path: chunking only shrinks the kernels and makes the
    // two kernel streams fight for the SMs (measured: 1 chunk 5.7 ms, 4 chunks 6.2 ms per 512 frames): one chunk.
    int nchunks = host ? (batch >= 64 ? 6 : 1) : 1;
    if (const char* e = getenv(host ? "ORBX_HOST_CHUNKS" : "ORBX_DEVICE_CHUNKS")) { const int v = atoi(e); if (v >= 1 && v <= ORBX_MAX_CHUNKS) nchunks = v; }
    if (nchunks > batch) nchunks = batch;
    int c_f0[ORBX_MAX_CHUNKS], c_cnt[ORBX_MAX_CHUNKS];
    {
        const bool taper = host && nchunks >= 4 && batch >= 16 * nchunks && !getenv("ORBX_UNIFORM_CHUNKS");
        const int edge = taper ? batch / 16 : 0;
        const int mid = taper ? nchunks - 2 : nchunks, rest = batch - 2 * edge;
        int f = 0, k = 0;
        if (taper) { c_f0[k] = 0; c_cnt[k++] = edge; f = edge; }
        for (int i = 0; i < mid; i++) {
            const int cnt = rest / mid + (i < rest % mid ? 1 : 0);
            c_f0[k] = f; c_cnt[k++] = cnt; f += cnt;
        }
        if (taper) { c_f0[k] = f; c_cnt[k++] = edge; }
        nchunks = k;
    }
    int32_t* dm12 = host ? m->d_out : d_matches12;
    int32_t* dnm = host ? m->d_nm : d_nmatches;
    // the side streams must not run ahead of work already queued on the kernel stream (previous call's carry)
    CKM(cudaEventRecord(m->ev_start, s));
    CKM(cudaStreamWaitEvent(m->s_match, m->ev_start, 0));
    if (host) {
        CKM(cudaStreamWaitEvent(m->s_h2d, m->ev_start, 0));
        CKM(cudaStreamWaitEvent(m->s_d2h, m->ev_start, 0));
        for (int c = 0; c < nchunks; c++) {
            const int f0 = c_f0[c], cnt = c_cnt[c];
            rc = orbx_ex_stage_input(ex, imgs, f0, cnt, width, height, stride, frame_stride, m->s_h2d);
            if (rc) return rc;
            CKM(cudaEventRecord(m->ev[c], m->s_h2d));
        }
    }
    for (int c = 0; c < nchunks; c++) {
        const int f0 = c_f0[c], cnt = c_cnt[c];
        if (host) {
            CKM(cudaStreamWaitEvent(s, m->ev[c], 0));
            rc = orbx_ex_run_staged(ex, f0, cnt, lap0, lap1, 1 + f0, s);
        } else {
            rc = orbx_ex_run_device(ex, imgs, stride, (long long)frame_stride, f0, cnt, lap0, lap1, 1 + f0, s);
        }
        if (rc) return rc;
        CKM(cudaEventRecord(m->ev_ext[c], s));
        // pairs (slot f0+i, slot f0+i+1) are matched on a second kernel stream, concurrently with the extraction of the
        // next chunk; each chunk owns its slice of the pair scratch
        CKM(cudaStreamWaitEvent(m->s_match, m->ev_ext[c], 0));
        rc = match_slots_impl(m, ex, m->d_pair_a + f0, m->d_pair_b + f0, cnt, f0, bounds, window, nnratio, check_ori,
                              dm12 + (size_t)f0 * m->K, dnm + f0,
                              d_knn_idx ? d_knn_idx + (size_t)f0 * m->K * 2 : nullptr, d_knn_dist ? d_knn_dist + (size_t)f0 * m->K * 2 : nullptr,
                              m->s_match);
        if (rc) return rc;
        CKM(cudaEventRecord(m->ev[ORBX_MAX_CHUNKS + c], m->s_match));
        if (host) {
            CKM(cudaStreamWaitEvent(m->s_d2h, m->ev[ORBX_MAX_CHUNKS + c], 0));
            rc = orbx_ex_fetch_async(ex, 1 + f0, cnt, f0, kps, desc, cap, n, mono_index, m->s_d2h, direct);
            if (rc) return rc;
            // matches: device rows have stride K; host rows have stride cap
            if (matches12) CKM(cudaMemcpy2DAsync(matches12 + (size_t)f0 * cap, sizeof(int32_t) * cap, dm12 + (size_t)f0 * m->K, sizeof(int32_t) * m->K,
                                                 sizeof(int32_t) * (cap < m->K ? cap : m->K), cnt, cudaMemcpyDeviceToHost, m->s_d2h));
            if (nmatches) CKM(cudaMemcpyAsync(nmatches + f0, dnm + f0, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, m->s_d2h));
        }
    }
    // slot 0 is read by the first chunk's matcher: carry the last frame over only after every matcher finished; the
    // caller's stream `s` thereby also waits for the matcher stream
    CKM(cudaStreamWaitEvent(s, m->ev[ORBX_MAX_CHUNKS + nchunks - 1], 0));
    rc = orbx_extractor_copy_slot(ex, batch, 0, s);
    if (rc) return rc;
    if (!host) return ORBX_OK;                       // asynchronous: the caller synchronises `s`
    CKM(cudaStreamSynchronize(m->s_d2h));
    CKM(cudaStreamSynchronize(m->s_match));
    rc = m_check_err(m, s);      // synchronises the kernel stream
    if (rc) return rc;
    return orbx_ex_fetch_finish(ex, batch, kps, desc, cap, n, mono_index, direct);
}

extern "C" int orbx_extract_match_batch(orbx_extractor* ex, orbx_matcher* m, const uint8_t* imgs, int batch, int width,
                                        int height, int stride, size_t frame_stride, int lap0, int lap1,
                                        const float bounds[4], int window, float nnratio, int check_ori,
                                        orbx_keypoint* kps, uint8_t* desc, int cap, int32_t* n, int32_t* mono_index,
                                        int32_t* matches12, int32_t* nmatches)
{
    if (!ex || !m || !imgs || batch < 1 || batch > m->P || !bounds) return ORBX_E_INVALID;
    if (width <= 0 || height <= 0) return ORBX_E_EMPTY;
    return extract_match_pipeline(ex, m, true, imgs, batch, width, height, stride, frame_stride, lap0, lap1, bounds, window, nnratio,
                                  check_ori, kps, desc, cap, n, mono_index, matches12, nmatches, nullptr, nullptr, nullptr, nullptr,
                                  orbx_ex_stream(ex));
}

extern "C" int orbx_extract_match_batch_device(orbx_extractor* ex, orbx_matcher* m, const uint8_t* d_imgs, int batch, int width,
                                               int height, int stride, size_t frame_stride, int lap0, int lap1,
                                               const float bounds[4], int window, float nnratio, int check_ori,
                                               int32_t* d_matches12, int32_t* d_nmatches, int32_t* d_knn_idx, int32_t* d_knn_dist,
                                               void* stream)
{
    if (!ex || !m || !d_imgs || batch < 1 || batch > m->P || !bounds || !d_matches12 || !d_nmatches || stride < width) return ORBX_E_INVALID;
    if (width <= 0 || height <= 0) return ORBX_E_EMPTY;
    return extract_match_pipeline(ex, m, false, d_imgs, batch, width, height, stride, frame_stride, lap0, lap1, bounds, window, nnratio,
                                  check_ori, nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr, d_matches12, d_nmatches, d_knn_idx,
                                  d_knn_dist, stream ? (cudaStream_t)stream : orbx_ex_stream(ex));
}

static int stereo_scratch(orbx_matcher* m, size_t bytes);

extern "C" int orbx_stereo_band_match(orbx_matcher* m, const orbx_keypoint* kl, const uint8_t* dl, int nl,
                                      const orbx_keypoint* kr, const uint8_t* dr, int nr, const float* scale_factors,
                                      int nlevels, int nrows, float min_d, float max_d, int32_t* best_idx, int32_t* best_dist)
{
    if (!m || nl < 0 || nr < 0 || nl > m->K || nr > m->K || nlevels < 1 || nlevels > ORBX_MAX_LEVELS || !scale_factors ||
        (nl > 0 && (!kl || !dl || !best_idx || !best_dist)) || (nr > 0 && (!kr || !dr))) return ORBX_E_INVALID;
    if (nl == 0) return ORBX_OK;
    CKM(cudaSetDevice(m->p.device));
    cudaStream_t s = m->stream;
    CKM(cudaMemcpyAsync(m->d_k1, kl, sizeof(orbx_keypoint) * nl, cudaMemcpyHostToDevice, s));
    CKM(cudaMemcpyAsync(m->d_d1, dl, (size_t)32 * nl, cudaMemcpyHostToDevice, s));
    if (nr) {
        CKM(cudaMemcpyAsync(m->d_k2, kr, sizeof(orbx_keypoint) * nr, cudaMemcpyHostToDevice, s));
        CKM(cudaMemcpyAsync(m->d_d2, dr, (size_t)32 * nr, cudaMemcpyHostToDevice, s));
    }
    {
        StereoArgs A{};
        A.kL = m->d_k1; A.dL = m->d_d1; A.capL = nl; A.kR = m->d_k2; A.dR = m->d_d2; A.capR = nr; A.nl1 = nl; A.nr1 = nr;
        A.nrows = nrows; A.minD = min_d; A.maxD = max_d;
        for (int l = 0; l < nlevels && l < ORBX_MAX_LEVELS; l++) A.sf[l] = scale_factors[l];
        A.best_idx = m->d_out; A.best_dist = m->d_out2;
        const int list_cap = (nr > 0 ? nr : 1) * stereo_rows_per_kp(A.sf, nlevels < ORBX_MAX_LEVELS ? nlevels : ORBX_MAX_LEVELS);
        int rc = stereo_scratch(m, sizeof(int32_t) * (size_t)list_cap);
        if (rc) return rc;
        const size_t smem = sizeof(int) * 2 * ((size_t)nrows + 1);
        if (smem > 200 * 1024) return ORBX_E_INVALID;
        if (smem > 48 * 1024) CKM(cudaFuncSetAttribute(k_stereo_band, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_stereo_band<<<1, STEREO_NT, smem, s>>>(A, reinterpret_cast<int32_t*>(m->d_st), list_cap); ORBX_COUNT_LAUNCH(1);
    }
    CKM(cudaGetLastError());
    CKM(cudaMemcpyAsync(best_idx, m->d_out, sizeof(int32_t) * nl, cudaMemcpyDeviceToHost, s));
    CKM(cudaMemcpyAsync(best_dist, m->d_out2, sizeof(int32_t) * nl, cudaMemcpyDeviceToHost, s));
    CKM(cudaStreamSynchronize(s));
    return ORBX_OK;
}

// Frame::ComputeStereoMatches (R/src/Frame.cc:785-962) on two extractors' device-resident results and pyramids
// Device scratch of one stereo call: row lists, best index / distance, and (when the caller's outputs live on the host)
// mvuRight / mvDepth / SAD rows.
struct StereoScratch { int32_t* lists; int list_cap; int32_t* best; float* u; float* z; int32_t* sad; };
static int stereo_prepare(orbx_matcher* m, orbx_extractor* left, orbx_extractor* right, int count, StereoScratch* S)
{
    const int capL = orbx_ex_out_cap(left), capR = orbx_ex_out_cap(right);
    OrbxPyrView v;
    int rc = orbx_ex_pyramid_view(left, 0, &v);
    if (rc) return rc;
    S->list_cap = capR * stereo_rows_per_kp(v.scale, v.nlevels);
    const size_t n_lists = (size_t)count * S->list_cap, n_row = (size_t)count * capL;
    if ((rc = stereo_scratch(m, sizeof(int32_t) * (n_lists + 5 * n_row)))) return rc;
    S->lists = reinterpret_cast<int32_t*>(m->d_st);
    S->best = S->lists + n_lists;
    S->u = reinterpret_cast<float*>(S->best + 2 * n_row); S->z = S->u + n_row;
    S->sad = reinterpret_cast<int32_t*>(S->z + n_row);
    return ORBX_OK;
}

// Launches the three stereo kernels for `count` pairs: pair p = (left slot slot_l + p, frame frame_l + p) x (right ...).
// Outputs have row stride `ostride`.
static int stereo_launch(orbx_matcher* m, orbx_extractor* left, orbx_extractor* right, int slot_l, int slot_r, int frame_l, int frame_r,
                         int count, float mb, float mbf, const StereoScratch& S, float* d_u, float* d_z, int32_t* d_sad, int ostride,
                         cudaStream_t s)
{
    orbx_keypoint *kL, *kR; uint8_t *dL, *dR; int32_t *nL, *nR; int capL, capR, slotsL, slotsR;
    int rc = orbx_extractor_results_device(left, &kL, &dL, &nL, nullptr, &capL, &slotsL);
    if (rc) return rc;
    if ((rc = orbx_extractor_results_device(right, &kR, &dR, &nR, nullptr, &capR, &slotsR))) return rc;
    if (count <= 0 || slot_l < 0 || slot_l + count > slotsL || slot_r < 0 || slot_r + count > slotsR || ostride < capL) return ORBX_E_INVALID;
    OrbxPyrView vL, vR, tmp;
    if ((rc = orbx_ex_pyramid_view(left, frame_l, &vL)) || (rc = orbx_ex_pyramid_view(right, frame_r, &vR))) return rc;
    if ((rc = orbx_ex_pyramid_view(left, frame_l + count - 1, &tmp)) || (rc = orbx_ex_pyramid_view(right, frame_r + count - 1, &tmp))) return rc;
    StereoArgs A{};
    A.kL = kL; A.dL = dL; A.nL = nL; A.capL = capL; A.kR = kR; A.dR = dR; A.nR = nR; A.capR = capR;
    A.slotL0 = slot_l; A.slotR0 = slot_r;
    A.nrows = vL.h[0]; A.minD = 0.0f; A.maxD = mbf / mb; A.mbf = mbf;      // minZ = mb (R/src/Frame.cc:815-818)
    for (int l = 0; l < vL.nlevels; l++) A.sf[l] = vL.scale[l];
    A.best_idx = S.best; A.best_dist = S.best + (size_t)count * capL;
    A.uright = d_u; A.depth = d_z; A.sad = d_sad; A.ostride = ostride;
    {
        const size_t smem = sizeof(int) * 2 * ((size_t)A.nrows + 1);
        if (smem > 200 * 1024) return ORBX_E_INVALID;
        if (smem > 48 * 1024) CKM(cudaFuncSetAttribute(k_stereo_band, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_stereo_band<<<count, STEREO_NT, smem, s>>>(A, S.lists, S.list_cap); ORBX_COUNT_LAUNCH(1);
    }
    k_stereo_refine<<<dim3((capL + 7) / 8, count), 256, 0, s>>>(A, vL, vR); ORBX_COUNT_LAUNCH(1);
    int npad = 1; while (npad < capL) npad <<= 1;
    if (sizeof(int) * npad > 48 * 1024) CKM(cudaFuncSetAttribute(k_stereo_outliers, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(int) * npad)));
    k_stereo_outliers<<<count, 1024, sizeof(int) * npad, s>>>(A, npad); ORBX_COUNT_LAUNCH(1);
    CKM(cudaGetLastError());
    return ORBX_OK;
}

static int stereo_scratch(orbx_matcher* m, size_t bytes)
{
    if (bytes <= m->st_bytes) return ORBX_OK;
    if (m->d_st) { cudaDeviceSynchronize(); cudaFree(m->d_st); }
    m->d_st = nullptr; m->st_bytes = 0;
    CKM(cudaMalloc((void**)&m->d_st, bytes));
    m->st_bytes = bytes;
    return ORBX_OK;
}

extern "C" int orbx_stereo_matches(orbx_matcher* m, orbx_extractor* left, orbx_extractor* right, int slot_l, int slot_r,
                                   int frame_l, int frame_r, float mb, float mbf, float* uright, float* depth,
                                   int32_t* sad_dist, int cap, int* n_left)
{
    if (!m || !left || !right || !uright || !depth) return ORBX_E_INVALID;
    orbx_keypoint* kL; uint8_t* dL; int32_t* nL; int capL, slotsL;
    int rc = orbx_extractor_results_device(left, &kL, &dL, &nL, nullptr, &capL, &slotsL);
    if (rc) return rc;
    if (slot_l < 0 || slot_l >= slotsL) return ORBX_E_INVALID;
    CKM(cudaSetDevice(m->p.device));
    // the extractors run on their own streams: wait for both, then work on the matcher's stream
    CKM(cudaStreamSynchronize(orbx_ex_stream(left)));
    CKM(cudaStreamSynchronize(orbx_ex_stream(right)));
    cudaStream_t s = m->stream;
    int nl = 0;
    CKM(cudaMemcpyAsync(&nl, nL + slot_l, sizeof(int), cudaMemcpyDeviceToHost, s));
    CKM(cudaStreamSynchronize(s));
    if (n_left) *n_left = nl;
    if (nl > cap) { orbx_set_error("%s%s", "orbx_stereo_matches: output capacity too small", ""); return ORBX_E_CAPACITY; }
    if (nl == 0) return ORBX_OK;
    StereoScratch S;
    if ((rc = stereo_prepare(m, left, right, 1, &S))) return rc;
    if ((rc = stereo_launch(m, left, right, slot_l, slot_r, frame_l, frame_r, 1, mb, mbf, S, S.u, S.z, S.sad, capL, s))) return rc;
    CKM(cudaMemcpyAsync(uright, S.u, sizeof(float) * nl, cudaMemcpyDeviceToHost, s));
    CKM(cudaMemcpyAsync(depth, S.z, sizeof(float) * nl, cudaMemcpyDeviceToHost, s));
    if (sad_dist) CKM(cudaMemcpyAsync(sad_dist, S.sad, sizeof(int32_t) * nl, cudaMemcpyDeviceToHost, s));
    CKM(cudaStreamSynchronize(s));
    return ORBX_OK;
}

// batched form on device buffers: pair p = slot / frame (first + p) of both extractors; everything is enqueued on `stream`
// (the stream the two orbx_extract_batch_device calls used); outputs are [count][capacity of the left extractor]
extern "C" int orbx_stereo_matches_batch_device(orbx_matcher* m, orbx_extractor* left, orbx_extractor* right, int first, int count,
                                                float mb, float mbf, float* d_uright, float* d_depth, int32_t* d_sad, void* stream)
{
    if (!m || !left || !right || !d_uright || !d_depth || count <= 0) return ORBX_E_INVALID;
    CKM(cudaSetDevice(m->p.device));
    cudaStream_t s = stream ? (cudaStream_t)stream : m->stream;
    StereoScratch S;
    int rc = stereo_prepare(m, left, right, count, &S);
    if (rc) return rc;
    return stereo_launch(m, left, right, first, first, first, first, count, mb, mbf, S, d_uright, d_depth, d_sad ? d_sad : S.sad,
                         orbx_ex_out_cap(left), s);
}

// batched form with host outputs: uright/depth are [count][cap] (rows beyond a frame's keypoint count are unspecified)
extern "C" int orbx_stereo_matches_batch(orbx_matcher* m, orbx_extractor* left, orbx_extractor* right, int first, int count,
                                         float mb, float mbf, float* uright, float* depth, int cap)
{
    if (!m || !left || !right || !uright || !depth || count <= 0 || cap <= 0) return ORBX_E_INVALID;
    CKM(cudaSetDevice(m->p.device));
    CKM(cudaStreamSynchronize(orbx_ex_stream(left)));
    CKM(cudaStreamSynchronize(orbx_ex_stream(right)));
    cudaStream_t s = m->stream;
    const int capL = orbx_ex_out_cap(left);
    StereoScratch S;
    int rc = stereo_prepare(m, left, right, count, &S);
    if (rc) return rc;
    if ((rc = stereo_launch(m, left, right, first, first, first, first, count, mb, mbf, S, S.u, S.z, S.sad, capL, s))) return rc;
    const int wcopy = cap < capL ? cap : capL;
    CKM(cudaMemcpy2DAsync(uright, sizeof(float) * cap, S.u, sizeof(float) * capL, sizeof(float) * wcopy, count, cudaMemcpyDeviceToHost, s));
    CKM(cudaMemcpy2DAsync(depth, sizeof(float) * cap, S.z, sizeof(float) * capL, sizeof(float) * wcopy, count, cudaMemcpyDeviceToHost, s));
    CKM(cudaStreamSynchronize(s));
    return ORBX_OK;
}

// One call = a batch of stereo frames from host memory to host results: both cameras extracted (Frame.cc:92-95 runs the
// two extractors side by side) and Frame::ComputeStereoMatches for every pair.  The batch is cut into chunks that flow
// through five streams (H2D | left extractor | right extractor | stereo kernels | D2H), so the copies and the stereo
// kernels of one chunk hide under the extraction of its neighbours.
extern "C" int orbx_extract_stereo_batch(orbx_matcher* m, orbx_extractor* left, orbx_extractor* right,
                                         const uint8_t* imgs_left, const uint8_t* imgs_right, int batch, int width, int height,
                                         int stride, size_t frame_stride, float mb, float mbf,
                                         orbx_keypoint* kps_l, uint8_t* desc_l, int32_t* n_l,
                                         orbx_keypoint* kps_r, uint8_t* desc_r, int32_t* n_r, int cap,
                                         float* uright, float* depth)
{
    if (!m || !left || !right || left == right || !imgs_left || !imgs_right || batch < 1 || !uright || !depth || cap <= 0) return ORBX_E_INVALID;
    if (width <= 0 || height <= 0) return ORBX_E_EMPTY;
    int rc;
    if ((rc = orbx_ex_configure(left, width, height)) || (rc = orbx_ex_configure(right, width, height))) return rc;
    if (orbx_ex_device(left) != m->p.device || orbx_ex_device(right) != m->p.device) return ORBX_E_INVALID;
    int slotsL = 0, slotsR = 0, capL = 0, capR = 0;
    { orbx_keypoint* k; uint8_t* d; int32_t* n;
      if ((rc = orbx_extractor_results_device(left, &k, &d, &n, nullptr, &capL, &slotsL)) || (rc = orbx_extractor_results_device(right, &k, &d, &n, nullptr, &capR, &slotsR))) return rc; }
    if (batch > slotsL - 1 || batch > slotsR - 1) { orbx_set_error("%s%s", "orbx_extract_stereo_batch: batch larger than max_batch of an extractor", ""); return ORBX_E_INVALID; }
    CKM(cudaSetDevice(m->p.device));
    if ((rc = ensure_pipeline(m))) return rc;
    cudaStream_t sL = orbx_ex_stream(left), sR = orbx_ex_stream(right);
    if (m->mono2_cap < batch) {
        if (m->h_mono2) cudaFreeHost(m->h_mono2);
        m->h_mono2 = nullptr; m->mono2_cap = 0;
        CKM(cudaMallocHost((void**)&m->h_mono2, sizeof(int32_t) * 2 * (size_t)batch));
        m->mono2_cap = batch;
    }
    int32_t* mono_l = m->h_mono2; int32_t* mono_r = m->h_mono2 + batch;      // monoIndex is not part of this call's results
    const bool directL = orbx_ex_can_fetch_direct(left, kps_l, desc_l, cap, n_l, mono_l), directR = orbx_ex_can_fetch_direct(right, kps_r, desc_r, cap, n_r, mono_r);
    int nchunks = batch >= 48 ? 6 : (batch >= 16 ? 4 : 1);      // measured on C2 / C3: 6 chunks 88.9 k / 62.0 k frames/s, 4: 87.6 / 61.2, 8: 86.6 / 59.1
    if (const char* e = getenv("ORBX_HOST_CHUNKS")) { const int v = atoi(e); if (v >= 1 && v <= ORBX_MAX_CHUNKS) nchunks = v; }
    if (nchunks > batch) nchunks = batch;
    const int per = (batch + nchunks - 1) / nchunks;
    nchunks = (batch + per - 1) / per;
    StereoScratch S{};
    // the side streams start after whatever the caller queued on the extractors' streams
    CKM(cudaEventRecord(m->ev_start, sL));
    CKM(cudaStreamWaitEvent(m->s_h2d, m->ev_start, 0));
    CKM(cudaStreamWaitEvent(m->s_match, m->ev_start, 0));
    CKM(cudaStreamWaitEvent(m->s_d2h, m->ev_start, 0));
    for (int c = 0; c < nchunks; c++) {
        const int f0 = c * per, cnt = f0 + per <= batch ? per : batch - f0;
        if ((rc = orbx_ex_stage_input(left, imgs_left, f0, cnt, width, height, stride, frame_stride, m->s_h2d))) return rc;
        CKM(cudaEventRecord(m->ev[c], m->s_h2d));
        if ((rc = orbx_ex_stage_input(right, imgs_right, f0, cnt, width, height, stride, frame_stride, m->s_h2d))) return rc;
        CKM(cudaEventRecord(m->ev_r[c], m->s_h2d));
    }
    for (int c = 0; c < nchunks; c++) {
        const int f0 = c * per, cnt = f0 + per <= batch ? per : batch - f0;
        CKM(cudaStreamWaitEvent(sL, m->ev[c], 0));
        if ((rc = orbx_ex_run_staged(left, f0, cnt, 0, 0, f0, sL))) return rc;
        CKM(cudaEventRecord(m->ev_ext[c], sL));
        CKM(cudaStreamWaitEvent(sR, m->ev_r[c], 0));
        if ((rc = orbx_ex_run_staged(right, f0, cnt, 0, 0, f0, sR))) return rc;
        CKM(cudaEventRecord(m->ev_r[ORBX_MAX_CHUNKS + c], sR));
        if (c == 0 && (rc = stereo_prepare(m, left, right, batch, &S))) return rc;      // needs the geometry of a queued batch
        CKM(cudaStreamWaitEvent(m->s_match, m->ev_ext[c], 0));
        CKM(cudaStreamWaitEvent(m->s_match, m->ev_r[ORBX_MAX_CHUNKS + c], 0));
        StereoScratch C = S;                                   // this chunk's slice of the scratch
        C.lists = S.lists + (size_t)f0 * S.list_cap; C.best = S.best + 2 * (size_t)f0 * capL;
        float* du = S.u + (size_t)f0 * capL; float* dz = S.z + (size_t)f0 * capL; int32_t* ds = S.sad + (size_t)f0 * capL;
        if ((rc = stereo_launch(m, left, right, f0, f0, f0, f0, cnt, mb, mbf, C, du, dz, ds, capL, m->s_match))) return rc;
        CKM(cudaEventRecord(m->ev[ORBX_MAX_CHUNKS + c], m->s_match));
        CKM(cudaStreamWaitEvent(m->s_d2h, m->ev[ORBX_MAX_CHUNKS + c], 0));
        if ((rc = orbx_ex_fetch_async(left, f0, cnt, f0, kps_l, desc_l, cap, n_l, mono_l, m->s_d2h, directL))) return rc;
        if ((rc = orbx_ex_fetch_async(right, f0, cnt, f0, kps_r, desc_r, cap, n_r, mono_r, m->s_d2h, directR))) return rc;
        const int wcopy = cap < capL ? cap : capL;
        CKM(cudaMemcpy2DAsync(uright + (size_t)f0 * cap, sizeof(float) * cap, du, sizeof(float) * capL, sizeof(float) * wcopy, cnt, cudaMemcpyDeviceToHost, m->s_d2h));
        CKM(cudaMemcpy2DAsync(depth + (size_t)f0 * cap, sizeof(float) * cap, dz, sizeof(float) * capL, sizeof(float) * wcopy, cnt, cudaMemcpyDeviceToHost, m->s_d2h));
    }
    CKM(cudaStreamSynchronize(m->s_d2h));
    CKM(cudaStreamSynchronize(sL));
    CKM(cudaStreamSynchronize(sR));
    if ((rc = orbx_ex_fetch_finish(left, batch, kps_l, desc_l, cap, n_l, nullptr, directL))) return rc;
    return orbx_ex_fetch_finish(right, batch, kps_r, desc_r, cap, n_r, nullptr, directR);
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
    if (bytes > m->gen_bytes) {
        if (m->d_gen) cudaFree(m->d_gen);
    if (m->d_st) cudaFree(m->d_st);
    if (m->h_mono2) cudaFreeHost(m->h_mono2);
        CKM(cudaMalloc((void**)&m->d_gen, bytes)); m->gen_bytes = bytes;
    }
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
    if (!k1 || !d1 || !valid1 || !k2 || !d2 || !fv1_nodes || !fv1_start || !fv1_feat || !fv2_nodes || !fv2_start || !fv2_feat) return ORBX_E_INVALID;
    const int nf1 = fv1_start[nfv1], nf2 = fv2_start[nfv2];
    if (nf1 < 0 || nf2 < 0 || fv1_start[0] != 0 || fv2_start[0] != 0) return ORBX_E_INVALID;
    for (int i = 0; i < nfv1; i++) if (fv1_start[i] > fv1_start[i + 1] || (i && fv1_nodes[i - 1] >= fv1_nodes[i])) { orbx_set_error("%s%s", "orbx_search_by_bow: FeatureVector 1 must be sorted by node id", ""); return ORBX_E_INVALID; }
    for (int i = 0; i < nfv2; i++) if (fv2_start[i] > fv2_start[i + 1] || (i && fv2_nodes[i - 1] >= fv2_nodes[i])) { orbx_set_error("%s%s", "orbx_search_by_bow: FeatureVector 2 must be sorted by node id", ""); return ORBX_E_INVALID; }
    for (int i = 0; i < nf1; i++) if ((unsigned)fv1_feat[i] >= (unsigned)n1) { orbx_set_error("%s%s", "orbx_search_by_bow: feature index out of range", ""); return ORBX_E_INVALID; }
    for (int i = 0; i < nf2; i++) if ((unsigned)fv2_feat[i] >= (unsigned)n2) { orbx_set_error("%s%s", "orbx_search_by_bow: feature index out of range", ""); return ORBX_E_INVALID; }
    CKM(cudaSetDevice(m->p.device));
    cudaStream_t s = m->stream;
    // one scratch block: [k1][d1][valid1][k2][d2][valid2][fv tables][outputs]
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_k1 = take(sizeof(orbx_keypoint) * n1), o_d1 = take((size_t)32 * n1), o_v1 = take(n1);
    const size_t o_k2 = take(sizeof(orbx_keypoint) * n2), o_d2 = take((size_t)32 * n2), o_v2 = take(n2);
    const size_t o_n1 = take(sizeof(int32_t) * nfv1), o_s1 = take(sizeof(int32_t) * (nfv1 + 1)), o_f1 = take(sizeof(int32_t) * (nf1 + 1));
    const size_t o_n2 = take(sizeof(int32_t) * nfv2), o_s2 = take(sizeof(int32_t) * (nfv2 + 1)), o_f2 = take(sizeof(int32_t) * (nf2 + 1));
    const size_t o_m = take(sizeof(int32_t) * n1), o_c = take(n2), o_b = take(n1), o_h = take(sizeof(int32_t) * (ORBX_HISTO_LENGTH + 2));
    if (off > m->gen_bytes) {
        if (m->d_gen) cudaFree(m->d_gen);
        m->d_gen = nullptr; m->gen_bytes = 0;
        CKM(cudaMalloc((void**)&m->d_gen, off)); m->gen_bytes = off;
    }
    uint8_t* B = m->d_gen;
    CKM(cudaMemcpyAsync(B + o_k1, k1, sizeof(orbx_keypoint) * n1, cudaMemcpyHostToDevice, s));
    CKM(cudaMemcpyAsync(B + o_d1, d1, (size_t)32 * n1, cudaMemcpyHostToDevice, s));
    CKM(cudaMemcpyAsync(B + o_v1, valid1, n1, cudaMemcpyHostToDevice, s));
    CKM(cudaMemcpyAsync(B + o_k2, k2, sizeof(orbx_keypoint) * n2, cudaMemcpyHostToDevice, s));
    CKM(cudaMemcpyAsync(B + o_d2, d2, (size_t)32 * n2, cudaMemcpyHostToDevice, s));
    if (valid2) CKM(cudaMemcpyAsync(B + o_v2, valid2, n2, cudaMemcpyHostToDevice, s));
    CKM(cudaMemcpyAsync(B + o_n1, fv1_nodes, sizeof(int32_t) * nfv1, cudaMemcpyHostToDevice, s));
    CKM(cudaMemcpyAsync(B + o_s1, fv1_start, sizeof(int32_t) * (nfv1 + 1), cudaMemcpyHostToDevice, s));
    if (nf1) CKM(cudaMemcpyAsync(B + o_f1, fv1_feat, sizeof(int32_t) * nf1, cudaMemcpyHostToDevice, s));
    CKM(cudaMemcpyAsync(B + o_n2, fv2_nodes, sizeof(int32_t) * nfv2, cudaMemcpyHostToDevice, s));
    CKM(cudaMemcpyAsync(B + o_s2, fv2_start, sizeof(int32_t) * (nfv2 + 1), cudaMemcpyHostToDevice, s));
    if (nf2) CKM(cudaMemcpyAsync(B + o_f2, fv2_feat, sizeof(int32_t) * nf2, cudaMemcpyHostToDevice, s));
    CKM(cudaMemsetAsync(B + o_m, 0xFF, sizeof(int32_t) * n1, s));
    CKM(cudaMemsetAsync(B + o_c, 0, n2, s));
    CKM(cudaMemsetAsync(B + o_h, 0, sizeof(int32_t) * (ORBX_HISTO_LENGTH + 2), s));
    BowArgs A;
    A.mode = mode;
    A.k1 = reinterpret_cast<const orbx_keypoint*>(B + o_k1); A.d1 = B + o_d1; A.valid1 = B + o_v1; A.n1 = n1;
    A.fv1_nodes = reinterpret_cast<const int32_t*>(B + o_n1); A.fv1_start = reinterpret_cast<const int32_t*>(B + o_s1);
    A.fv1_feat = reinterpret_cast<const int32_t*>(B + o_f1); A.nfv1 = nfv1;
    A.k2 = reinterpret_cast<const orbx_keypoint*>(B + o_k2); A.d2 = B + o_d2; A.valid2 = valid2 ? B + o_v2 : nullptr; A.n2 = n2;
    A.fv2_nodes = reinterpret_cast<const int32_t*>(B + o_n2); A.fv2_start = reinterpret_cast<const int32_t*>(B + o_s2);
    A.fv2_feat = reinterpret_cast<const int32_t*>(B + o_f2); A.nfv2 = nfv2;
    A.nnratio = nnratio; A.check_ori = check_ori;
    A.matches12 = reinterpret_cast<int32_t*>(B + o_m); A.claimed2 = B + o_c; A.bin_of = B + o_b; A.hist = reinterpret_cast<int32_t*>(B + o_h);
    k_bow_match<<<(nfv1 + 7) / 8, 256, 0, s>>>(A); ORBX_COUNT_LAUNCH(1);
    k_bow_finish<<<1, 256, 0, s>>>(A, A.hist + ORBX_HISTO_LENGTH + 1); ORBX_COUNT_LAUNCH(1);
    CKM(cudaGetLastError());
    int nm = 0;
    CKM(cudaMemcpyAsync(matches12, A.matches12, sizeof(int32_t) * n1, cudaMemcpyDeviceToHost, s));
    CKM(cudaMemcpyAsync(&nm, A.hist + ORBX_HISTO_LENGTH + 1, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CKM(cudaStreamSynchronize(s));
    if (nmatches) *nmatches = nm;
    return ORBX_OK;
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
    if (bytes > m->gen_bytes) {
        if (m->d_gen) cudaFree(m->d_gen);
        m->d_gen = nullptr; m->gen_bytes = 0;
        CKM(cudaMalloc((void**)&m->d_gen, bytes)); m->gen_bytes = bytes;
    }
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

extern "C" int orbx_popc_peak(int device, double* popc_per_s, double* lop3_per_s)
{
    CKM(cudaSetDevice(device));
    unsigned* sink; CKM(cudaMalloc((void**)&sink, 4));
    cudaEvent_t e0, e1; CKM(cudaEventCreate(&e0)); CKM(cudaEventCreate(&e1));
    const int iters = 4096, blocks = 148 * 8, threads = 256;
    for (int which = 0; which < 2; which++) {
        float best = 1e30f;
        for (int rep = 0; rep < 5; rep++) {
            CKM(cudaEventRecord(e0));
            if (which == 0) k_popc_probe<<<blocks, threads>>>(rep, iters, sink); else k_lop3_probe<<<blocks, threads>>>(rep, iters, sink); ORBX_COUNT_LAUNCH(1);
            CKM(cudaEventRecord(e1));
            CKM(cudaEventSynchronize(e1));
            float ms; CKM(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        // popc probe: 8 popc per iteration; lop3 probe: 8 LOP3 per iteration (and-xor fuses into one LOP3)
        const double ops = (double)iters * 8 * blocks * threads;
        if (which == 0 && popc_per_s) *popc_per_s = ops / (best * 1e-3);
        if (which == 1 && lop3_per_s) *lop3_per_s = ops / (best * 1e-3);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
    return ORBX_OK;
}
