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
#include "orbx_match_internal.h"
#include "orbx_bfknn_tc.cuh"

namespace {

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

// The same search on the tensor cores (orbx_bfknn_tc.cuh): grid (query tiles of 128, train splits, pairs), one CTA per SM.
__global__ void __launch_bounds__(bftc::NT, 1) k_bf_knn2_tc(BfArgs A)
{
    extern __shared__ __align__(128) uint8_t bf_smem[];
    const int p = blockIdx.z, split = blockIdx.y;
    const uint8_t* q; const uint8_t* t; int nq; long long nt;
    if (A.q) { q = A.q; t = A.t; nq = A.nq; nt = A.nt; }
    else {
        const int sa = A.a[p], sb = A.b[p];
        q = A.desc + (long long)sa * A.cap * 32; nq = A.n[sa];
        t = A.desc + (long long)sb * A.cap * 32; nt = A.n[sb];
    }
    if ((long long)blockIdx.x * bftc::MQ >= nq) return;
    const long long t_begin = (long long)split * A.chunk;
    long long t_end = t_begin + A.chunk; if (t_end > nt) t_end = nt;
    int32_t* oi; int32_t* od;
    if (A.nsplit > 1) {
        const long long o = (((long long)p * A.nsplit + split) * A.out_stride) * 2;
        oi = A.part_idx + o; od = A.part_dist + o;
    } else {
        const long long o = ((long long)p * A.out_stride) * 2;
        oi = A.idx + o; od = A.dist + o;
    }
    bftc::bf_tile_body(q, nq, blockIdx.x * bftc::MQ, t, t_begin, t_end, A.idx_base, oi, od, bf_smem);
}

// The warp-specialised form (producers / MMA issuer / consumers, no block-wide barrier per tile): the default.
__global__ void __launch_bounds__(bftc::ws::NT, 1) k_bf_knn2_tcws(BfArgs A)
{
    extern __shared__ __align__(128) uint8_t bf_smem[];
    const int p = blockIdx.z, split = blockIdx.y;
    orbx_pdl_prologue();
    const uint8_t* q; const uint8_t* t; int nq; long long nt;
    if (A.q) { q = A.q; t = A.t; nq = A.nq; nt = A.nt; }
    else {
        const int sa = A.a[p], sb = A.b[p];
        q = A.desc + (long long)sa * A.cap * 32; nq = A.n[sa];
        t = A.desc + (long long)sb * A.cap * 32; nt = A.n[sb];
    }
    if ((long long)blockIdx.x * bftc::MQ >= nq) return;
    const long long t_begin = (long long)split * A.chunk;
    long long t_end = t_begin + A.chunk; if (t_end > nt) t_end = nt;
    int32_t* oi; int32_t* od;
    if (A.nsplit > 1) {
        const long long o = (((long long)p * A.nsplit + split) * A.out_stride) * 2;
        oi = A.part_idx + o; od = A.part_dist + o;
    } else {
        const long long o = ((long long)p * A.out_stride) * 2;
        oi = A.idx + o; od = A.dist + o;
    }
    bftc::ws::bf_tile_body(q, nq, blockIdx.x * bftc::MQ, t, t_begin, t_end, A.idx_base, oi, od, bf_smem);
}

// The small results of a single-chunk call in one place: mail = {extractor error word, matcher error word, n[batch], monoIndex[batch],
// nmatches[batch]}; `mail` is mapped pinned host memory, so the stores ARE the transfer (complete when the kernel is).
__global__ void k_mailbox(const int* ex_n, const int* ex_mono, const int* nm, const unsigned* ex_err, const unsigned* m_err, int batch, int* mail)
{
    orbx_pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { mail[0] = (int)*ex_err; mail[1] = (int)*m_err; }
    if (i < batch) { mail[2 + i] = ex_n[i]; mail[2 + batch + i] = ex_mono[i]; mail[2 + 2 * batch + i] = nm ? nm[i] : 0; }
}

// merge partial top-2 tables: parts laid out [pair][part][stride][2]; lexicographic (dist, idx)
__global__ void k_knn2_merge(const int32_t* pidx, const int32_t* pdist, int nparts, int stride, int nq_fixed,
                             const int* n, const int* a, int32_t* idx, int32_t* dist)
{
    const int p = blockIdx.y;
    orbx_pdl_prologue();
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
// fills PairDesc for result slots and synthesises the SearchForInitialization queries:
// level-0 keypoints of F1 search a window around vbPrevMatched = their own position, levels [0,0]
__global__ void k_setup_slot_pairs(WinBufs W, const orbx_keypoint* kps, const uint8_t* desc, const int* n,
                                   const int* a, const int* b, int cap, float window)
{
    const int p = blockIdx.x;
    orbx_pdl_prologue();
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
__global__ void __launch_bounds__(GRID_NT) k_grid_build(WinBufs W, int kmax, int only_level)
{
    // The records of a pair sorted by (cell, keypoint index) = the reference's visit order inside a cell (insertion order).  A stable
    // counting sort over the NCELL cells: histogram with shared-memory atomics, block-wide exclusive scan (also the cell_start table
    // the candidate kernel reads), then ONE warp places the keypoints in index order, 32 per step: match.any groups the lanes of a
    // cell, the lowest lane of a group advances the cell's cursor by the group size, a lane's place is cursor + its rank in the
    // group.  (The first version sorted 32-bit keys with a 55-step bitonic network: 15 us for one pair.)
    extern __shared__ int g_sm[];
    int* cur = g_sm;                                                        // [NCELL + 2] counts -> exclusive starts -> cursors
    unsigned short* cell_of = reinterpret_cast<unsigned short*>(cur + NCELL + 2);     // [kmax] cell of keypoint i (NCELL = outside the grid)
    unsigned short* order = cell_of + kmax;                                 // [kmax] keypoint index at sorted position
    __shared__ int s_wsum[GRID_NT / 32];
    const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    orbx_pdl_prologue();
    if (tid == 0) W.pool_used[p] = 0;       // the candidate pool of the pair starts empty (k_window_candidates fills it)
    const PairDesc P = W.pairs[p];
    const int n = min(min(P.n2, W.K), kmax);
    for (int c = tid; c < NCELL + 2; c += GRID_NT) cur[c] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += GRID_NT) {
        const orbx_keypoint kp = P.k2[i];
        const int px = (int)roundf(__fmul_rn(__fsub_rn(kp.x, W.minX), W.wInv));      // PosInGrid: round, not floor
        const int py = (int)roundf(__fmul_rn(__fsub_rn(kp.y, W.minY), W.hInv));
        // only_level >= 0: every query of the search asks for that octave alone (SearchForInitialization: level1 == 0 on both sides,
        // :714-718), so the other keypoints never enter the grid; dropping them keeps the visit order of the rest
        const int c = (px >= 0 && px < GC && py >= 0 && py < GR && (only_level < 0 || kp.octave == only_level)) ? px * GR + py : NCELL;
        cell_of[i] = (unsigned short)c;
        atomicAdd(&cur[c], 1);
    }
    __syncthreads();
    // exclusive scan of the NCELL + 1 counts: PER consecutive entries per thread, warp scan, block scan
    constexpr int PER = (NCELL + 1 + GRID_NT - 1) / GRID_NT;
    int v[PER], sum = 0;
#pragma unroll
    for (int k = 0; k < PER; k++) { const int c = tid * PER + k; v[k] = c <= NCELL ? cur[c] : 0; sum += v[k]; }
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < GRID_NT / 32 ? s_wsum[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
        if (lane < GRID_NT / 32) s_wsum[lane] = w;                          // inclusive over warps
    }
    __syncthreads();
    int run = inc - sum + (warp ? s_wsum[warp - 1] : 0);
    int* cs = W.cell_start + (long long)p * (NCELL + 1);
#pragma unroll
    for (int k = 0; k < PER; k++) {
        const int c = tid * PER + k;
        if (c <= NCELL) { cur[c] = run; cs[c] = run; }                      // cs[c] = first sorted position whose cell >= c; out-of-grid last
        run += v[k];
    }
    __syncthreads();
    if (warp == 0) {
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + lane;
            const bool act = i < n;
            const int c = act ? (int)cell_of[i] : NCELL + 1;                // idle lanes form their own group on an unused cursor
            const unsigned grp = __match_any_sync(0xffffffffu, c);
            const int leader = __ffs(grp) - 1;
            int base = 0;
            if (lane == leader) { base = cur[c]; cur[c] = base + __popc(grp); }
            base = __shfl_sync(0xffffffffu, base, leader);
            if (act) order[base + __popc(grp & ((1u << lane) - 1u))] = (unsigned short)i;
            __syncwarp();
        }
    }
    __syncthreads();
    const int n_in = cs[NCELL];                                             // (written above by this CTA; visible after the barrier)
    uint16_t* items = W.items + (long long)p * W.K;
    float4* skp = W.skp + (long long)p * W.K;
    for (int i = tid; i < n; i += GRID_NT) {
        if (i < n_in) {
            const int idx = order[i];
            items[i] = (uint16_t)idx;
            const orbx_keypoint kp = P.k2[idx];
            skp[i] = make_float4(kp.x, kp.y, __int_as_float(kp.octave), __int_as_float(idx));
        } else {
            items[i] = 0xFFFF;                                              // outside the grid: never visited
        }
    }
}

constexpr int CAND_WARPS = 8;
constexpr int CAND_CACHE = 4;           // warp-steps (x 32 records) of a window whose test outcome is kept in registers between the two passes

// Frame::GetFeaturesInArea + descriptor distances, one warp per query.  Pool entry: i2 | dist << 16 | octave << 25
__global__ void __launch_bounds__(CAND_WARPS * 32, 6) k_window_candidates(WinBufs W)
{
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.y;
    orbx_pdl_prologue();
    const PairDesc P = W.pairs[p];
    const int qi = blockIdx.x * CAND_WARPS + (threadIdx.x >> 5);
    if (qi >= P.nq || qi >= W.K) return;
    int* q_off = W.q_off + (long long)p * W.K + qi;
    int* q_cnt = W.q_cnt + (long long)p * W.K + qi;
    const orbx_proj_query Q = P.q[qi];
    bool any = (Q.valid & 1) != 0;
    int cx0 = 0, cx1 = -1, cy0 = 0, cy1 = -1;
    if (any) {
        // :639-660, all fp32
        cx0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(Q.u, W.qminX), Q.r), W.wInv)));
        cx1 = min(GC - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(Q.u, W.qminX), Q.r), W.wInv)));
        cy0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(Q.v, W.qminY), Q.r), W.hInv)));
        cy1 = min(GR - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(Q.v, W.qminY), Q.r), W.hInv)));
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
    // pass 0 counts the records that pass the tests and keeps the outcome of the first CAND_CACHE warp-steps in registers
    // (index | octave << 16, or ~0), so that pass 1 does not search, load and test those records a second time
    uint32_t cache[CAND_CACHE];
#pragma unroll
    for (int c = 0; c < CAND_CACHE; c++) cache[c] = 0xFFFFFFFFu;
    auto test_record = [&](int j, int& i2, int& oct) -> bool {
        int lo = 0, hi = nr - 1;                          // last range whose prefix <= j
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (rp[mid] <= j) lo = mid; else hi = mid - 1; }
        const float4 rec = __ldg(skp + rs[lo] + (j - rp[lo]));
        oct = __float_as_int(rec.z); i2 = __float_as_int(rec.w);
        bool ok = true;
        if (check_levels) {
            if (oct < Q.minl) ok = false;
            if (Q.maxl >= 0 && oct > Q.maxl) ok = false;
        }
        const float dx = __fsub_rn(rec.x, Q.u), dy = __fsub_rn(rec.y, Q.v);
        if (!(fabsf(dx) < Q.r && fabsf(dy) < Q.r)) ok = false;
        // Fuse (ORBmatcher.cc:1525-1552): e2 * invSigma2 (float) against the double chi-square bound; a keypoint with a right
        // coordinate (mvuRight >= 0) adds the squared right-image error and uses the 3-dof bound
        if (ok && W.gate_chi2 > 0.0) {
            float e2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
            double bound = W.gate_chi2;
            if (W.gate_chi2_stereo > 0.0 && P.uright2) {
                const float ur2 = P.uright2[i2];
                if (ur2 >= 0) { const float er = __fsub_rn(Q.ur, ur2); e2 = __fadd_rn(e2, __fmul_rn(er, er)); bound = W.gate_chi2_stereo; }
            }
            if ((double)__fmul_rn(e2, W.inv_sigma2[min(oct, ORBX_MAX_LEVELS - 1)]) > bound) ok = false;
        }
        // stereo gate (ORBmatcher.cc:93-98 / :2049-2055): not order dependent, applied here
        if (ok && P.uright2 && !(W.gate_chi2 > 0.0)) {
            const float ur2 = P.uright2[i2];
            if (ur2 > 0 && fabsf(__fsub_rn(Q.ur, ur2)) > Q.r) ok = false;
        }
        return ok;
    };
    {
        int pos = 0;
#pragma unroll
        for (int c = 0; c < CAND_CACHE; c++) {
            const int j = c * 32 + lane;
            if (c * 32 < T) {                               // warp-uniform
                int i2 = 0, oct = 0;
                const bool ok = j < T && test_record(j, i2, oct);
                if (ok) cache[c] = (uint32_t)i2 | ((uint32_t)oct << 16);
                pos += __popc(__ballot_sync(0xffffffffu, ok));
            }
        }
        for (int j0 = CAND_CACHE * 32; j0 < T; j0 += 32) {
            const int j = j0 + lane;
            int i2 = 0, oct = 0;
            const bool ok = j < T && test_record(j, i2, oct);
            pos += __popc(__ballot_sync(0xffffffffu, ok));
        }
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
    {
        int pos = 0;
        auto emit = [&](bool ok, int i2, int oct) {
            const unsigned bal = __ballot_sync(0xffffffffu, ok);
            if (ok) {
                const int o = pos + __popc(bal & ((1u << lane) - 1));
                const uint4 t0 = reinterpret_cast<const uint4*>(P.d2)[2 * i2], t1 = reinterpret_cast<const uint4*>(P.d2)[2 * i2 + 1];
                const int d = hamming256(q0, q1, t0, t1);
                const uint32_t e = (uint32_t)i2 | ((uint32_t)d << 16) | ((uint32_t)oct << 25);
                if (base + o < W.POOL) W.pool[(long long)p * W.POOL + base + o] = e;
                if (d < td0) { td1 = td0; tk1 = tk0; te1 = te0; td0 = d; tk0 = o; te0 = e; }      // ranks grow per lane: first minimum kept
                else if (d < td1) { td1 = d; tk1 = o; te1 = e; }
            }
            pos += __popc(bal);
        };
#pragma unroll
        for (int c = 0; c < CAND_CACHE; c++)
            if (c * 32 < T) emit(cache[c] != 0xFFFFFFFFu, (int)(cache[c] & 0xFFFF), (int)((cache[c] >> 16) & 0xFF));
        for (int j0 = CAND_CACHE * 32; j0 < T; j0 += 32) {
            const int j = j0 + lane;
            int i2 = 0, oct = 0;
            const bool ok = j < T && test_record(j, i2, oct);
            emit(ok, i2, oct);
        }
    }
    // the resolve kernel starts from these two and rescans the list only when one of them has been taken meanwhile
    warp_top2(td0, tk0, te0, td1, te1, tk1);
    if (lane == 0)
        W.top2[(long long)p * W.K + qi] = make_uint2(td0 == 0x7fffffff ? 0xFFFFFFFFu : te0, td1 == 0x7fffffff ? 0xFFFFFFFFu : te1);
}


// mode 2 = SearchForInitialization, 0 / 1 = SearchByProjection overloads (see orbx.h)
// out: mode 2 -> matches12 [P][K] (+ prev_xy update when prev != null); modes 0/1 -> assigned [P][K] (in/out)
constexpr int INIT_UNRESOLVED = -1;       // nmatches[p] marker: the parallel resolve gave up on pair p, the sequential kernel takes it
constexpr int INIT_NT = 256;             // threads per pair in a batch; a launch over few pairs uses INIT_NT_FEW (one query per thread: latency)
constexpr int INIT_NT_FEW = 1024;
constexpr int INIT_CLAIMS = 4;            // claims kept per keypoint (more -> sequential fallback)
constexpr int INIT_MAX_ROUNDS = 32;
constexpr int INIT_MAX_K = 8192;          // shared memory of the parallel resolve: 6 ints per keypoint (192 KB at 8192)
constexpr int PROJ_MAX_K = 16384;         // k_proj_resolve: 3 ints per keypoint

// ---- k_init_resolve: SearchForInitialization's order-dependent part (R/src/ORBmatcher.cc:721-784), in parallel --------------
// The reference visits the queries in order; query i skips a candidate i2 whose recorded match distance is <= its own
// (vMatchedDistance[i2] <= dist, :741), i.e. it depends on the ACCEPTED claims of the queries before it.  vMatchedDistance[i2]
// only ever decreases, so the state query i sees is "min distance over the claims on i2 by queries j < i".
// The sequential process is the unique fixed point of  X -> F(X),  F(X)[i] = decision of query i under the claims X[j], j < i
// (by induction on i: after k rounds the first k decisions are final), so iterating F from the empty set converges to exactly
// the reference's result, and "no decision changed in a round" proves the fixed point is reached.  Dependency chains are short
// (a claim only matters to later queries that share the keypoint), so a handful of rounds suffices; every round is one
// candidate-list scan per query, all queries in parallel: ~10 us instead of the 120 us of one warp walking 1000 queries in order.
// Steals (:760-764): the LAST claimer of a keypoint owns it; the rotation histogram counts every accepted claim, stolen or not,
// as the reference's rotHist lists do (:772-779).  A keypoint with more than INIT_CLAIMS claimers, or no convergence within
// INIT_MAX_ROUNDS, hands the pair to the sequential kernel (nmatches[p] = INIT_UNRESOLVED): still exact, never silently wrong.
__global__ void __launch_bounds__(INIT_NT_FEW) k_init_resolve(WinBufs W, float nnratio, int check_ori, int32_t* out, int32_t* nmatches, float* prev_xy)
{
    extern __shared__ int s_mem[];
    __shared__ int hist[ORBX_HISTO_LENGTH];
    __shared__ int s_chg[2];                  // a decision changed in this round (indexed by round parity: reset two rounds later)
    __shared__ int s_over;                    // claim-list overflow
    __shared__ int s_count;
    const int tid = threadIdx.x, NTH = blockDim.x;
    const int p = blockIdx.x;
    orbx_pdl_prologue();
    const PairDesc P = W.pairs[p];
    const int nq = min(P.nq, W.K), n2 = min(P.n2, W.K);
    uint32_t* dec = reinterpret_cast<uint32_t*>(s_mem);              // [K] by query: i2 | dist << 16, or NONE
    int* cnt = s_mem + W.K;                                          // [K] by keypoint: number of claims
    uint32_t* lists = reinterpret_cast<uint32_t*>(s_mem + 2 * W.K);  // [K][INIT_CLAIMS]: claiming query | dist << 16
    constexpr uint32_t NONE = 0xFFFFFFFFu;
    int32_t* res = out + (long long)p * W.K;
    uint8_t* bin_of = W.bin_of + (long long)p * W.K;
    const uint32_t* pool = W.pool + (long long)p * W.POOL;
    const int* q_off = W.q_off + (long long)p * W.K;
    const int* q_cnt = W.q_cnt + (long long)p * W.K;
    for (int i = tid; i < nq; i += NTH) dec[i] = NONE;
    if (tid < ORBX_HISTO_LENGTH) hist[tid] = 0;
    if (tid == 0) { s_over = 0; s_count = 0; }
    bool converged = false;
    for (int round = 0; round < INIT_MAX_ROUNDS; round++) {
        for (int i = tid; i < n2; i += NTH) cnt[i] = 0;
        if (tid == 0) s_chg[round & 1] = 0;
        __syncthreads();
        // claims of the current decisions, per keypoint
        for (int i = tid; i < nq; i += NTH) {
            const uint32_t d = dec[i];
            if (d != NONE) {
                const int i2 = d & 0xFFFF;
                const int slot = atomicAdd(&cnt[i2], 1);
                if (slot < INIT_CLAIMS) lists[i2 * INIT_CLAIMS + slot] = (uint32_t)i | (d & 0xFFFF0000u);
                else s_over = 1;
            }
        }
        __syncthreads();
        if (s_over) break;
        // every query decides again under the claims of the queries before it
        for (int i = tid; i < nq; i += NTH) {
            const int c = q_cnt[i];
            if (c <= 0) continue;
            const int off = q_off[i];
            int best = 0x7fffffff, second = 0x7fffffff, bidx = -1;
            for (int k = 0; k < c; k++) {
                const uint32_t e = __ldg(pool + off + k);
                const int i2 = e & 0xFFFF, d = (e >> 16) & 0x1FF;
                if (d >= second) continue;                       // cannot change (best, second): skip the claim lookup
                const int nc = min(cnt[i2], INIT_CLAIMS);
                bool blocked = false;
                for (int t = 0; t < nc; t++) {
                    const uint32_t le = lists[i2 * INIT_CLAIMS + t];
                    if ((int)(le & 0xFFFF) < i && (int)(le >> 16) <= d) blocked = true;      // vMatchedDistance[i2] <= dist (:741)
                }
                if (blocked) continue;
                if (d < best) { second = best; best = d; bidx = i2; }
                else if (d < second) second = d;
            }
            uint32_t nd = NONE;
            if (best <= ORBX_TH_LOW && (float)best < (float)second * nnratio) nd = (uint32_t)bidx | ((uint32_t)best << 16);   // :756-758
            if (nd != dec[i]) { dec[i] = nd; s_chg[round & 1] = 1; }
        }
        __syncthreads();
        if (!s_chg[round & 1]) { converged = true; break; }
    }
    if (!converged) { if (tid == 0) nmatches[p] = INIT_UNRESOLVED; return; }
    // the claim lists are those of the final decisions (the last round changed nothing)
    for (int i = tid; i < nq; i += NTH) {
        const uint32_t d = dec[i];
        int r = -1; uint8_t bin = 0xFF;
        if (d != NONE) {
            const int i2 = d & 0xFFFF;
            const int nc = min(cnt[i2], INIT_CLAIMS);
            int owner = -1;
            for (int t = 0; t < nc; t++) owner = max(owner, (int)(lists[i2 * INIT_CLAIMS + t] & 0xFFFF));
            if (owner == i) r = i2;                              // later claimers steal (:760-764)
            if (check_ori) { bin = (uint8_t)rot_bin(P.q[i].angle, P.k2[i2].angle); atomicAdd(&hist[bin], 1); }
        }
        res[i] = r; bin_of[i] = bin;
    }
    __syncthreads();
    int ind1 = -1, ind2 = -1, ind3 = -1;
    if (check_ori) three_maxima(hist, ind1, ind2, ind3);
    int mine = 0;
    for (int i = tid; i < nq; i += NTH) {
        int m = res[i];
        if (check_ori) {
            const int b = bin_of[i];
            if (b != 0xFF && b != ind1 && b != ind2 && b != ind3) { m = -1; res[i] = -1; }
        }
        if (m >= 0) {
            mine++;
            if (prev_xy) { prev_xy[((long long)p * W.K + i) * 2] = P.k2[m].x; prev_xy[((long long)p * W.K + i) * 2 + 1] = P.k2[m].y; }
        }
    }
    if (mine) atomicAdd(&s_count, mine);
    __syncthreads();
    if (tid == 0) nmatches[p] = s_count;
}

// ---- SearchByProjection modes 0 / 1 as a parallel fixed point ----
// The sequential rule (:89-91 / :2045-2047): query i may not take a keypoint that is occupied, i.e. held by a MapPoint with
// observations before the call (res[i2] >= 0) or claimed by an OCCUPYING query j < i (valid bit 1 clear); everything else about a
// query's decision (best / second, thresholds, the level rule of mode 1) depends on its own candidate list only.  So the visit
// order matters only through  minOcc[i2] = the smallest occupying query index that claims i2,  and the sequential result is the
// unique fixed point of  X -> F(X),  F(X)[i] = decision of query i when i2 is barred iff minOcc_X[i2] < i  (induction on i, as for
// k_init_resolve).  A round = one atomicMin per deciding query + one candidate-list scan per query, all queries in parallel; chains
// are short (a query that loses its keypoint moves to another one), so a handful of rounds replaces one warp walking the queries in
// order (0.55 ms for 1000 MapPoints against 1000 keypoints).  Final state: a keypoint belongs to the LAST query that claimed it
// (non-occupying claimers are overwritten, :2063), every accepting query counts as a match and enters the rotation histogram.
// No convergence within INIT_MAX_ROUNDS -> nmatches[p] = INIT_UNRESOLVED and the sequential kernel takes the pair (res untouched).
__global__ void __launch_bounds__(INIT_NT_FEW) k_proj_resolve(WinBufs W, int mode, float nnratio, int check_ori, int max_dist, int32_t* out, int32_t* nmatches)
{
    extern __shared__ int s_mem[];
    __shared__ int hist[ORBX_HISTO_LENGTH];
    __shared__ int s_chg[2];
    __shared__ int s_acc, s_cull;
    const int tid = threadIdx.x, NTH = blockDim.x;
    const int p = blockIdx.x;
    orbx_pdl_prologue();
    const PairDesc P = W.pairs[p];
    const int nq = min(P.nq, W.K), n2 = min(P.n2, W.K);
    uint32_t* dec = reinterpret_cast<uint32_t*>(s_mem);              // [K] by query: claimed keypoint, or NONE
    int* base = s_mem + W.K;                                         // [K] by keypoint: -1 occupied before the call, else INT_MAX; later: owner
    int* min_occ = s_mem + 2 * W.K;                                  // [K] by keypoint: smallest occupying claimer; later: claim_of by query
    constexpr uint32_t NONE = 0xFFFFFFFFu;
    int32_t* res = out + (long long)p * W.K;
    uint8_t* bin_of = W.bin_of + (long long)p * W.K;
    const uint32_t* pool = W.pool + (long long)p * W.POOL;
    const int* q_off = W.q_off + (long long)p * W.K;
    const int* q_cnt = W.q_cnt + (long long)p * W.K;
    for (int i = tid; i < nq; i += NTH) dec[i] = NONE;
    for (int i = tid; i < n2; i += NTH) base[i] = res[i] >= 0 ? -1 : 0x7fffffff;
    if (tid < ORBX_HISTO_LENGTH) hist[tid] = 0;
    if (tid == 0) { s_acc = 0; s_cull = 0; }
    bool converged = false;
    for (int round = 0; round < INIT_MAX_ROUNDS; round++) {
        for (int i = tid; i < n2; i += NTH) min_occ[i] = base[i];
        if (tid == 0) s_chg[round & 1] = 0;
        __syncthreads();
        for (int i = tid; i < nq; i += NTH) {
            const uint32_t d = dec[i];
            if (d != NONE && !(P.q[i].valid & 2)) atomicMin(&min_occ[d], i);
        }
        __syncthreads();
        for (int i = tid; i < nq; i += NTH) {
            const int c = q_cnt[i];
            if (c <= 0) continue;
            const int off = q_off[i];
            int d0 = 0x7fffffff, d1 = 0x7fffffff; uint32_t e0 = 0, e1 = 0;
            for (int k = 0; k < c; k++) {
                const uint32_t e = __ldg(pool + off + k);
                const int d = (e >> 16) & 0x1FF;
                if (d >= d1) continue;                               // cannot change (best, second)
                if (min_occ[e & 0xFFFF] < i) continue;               // occupied for this query
                if (d < d0) { d1 = d0; e1 = e0; d0 = d; e0 = e; }
                else { d1 = d; e1 = e; }
            }
            uint32_t nd = NONE;
            if (mode == 0) {
                if (d0 <= max_dist) nd = e0 & 0xFFFF;                // :2038, :2068 (or the caller's bound)
            } else {
                const int bd = d0 < 256 ? d0 : 256, bd2 = d1 < 256 ? d1 : 256;      // :78-141: best / second start at 256 with level -1
                const int bl = d0 < 256 ? (int)(e0 >> 25) : -1, bl2 = d1 < 256 ? (int)(e1 >> 25) : -1;
                if (bd <= ORBX_TH_HIGH && !(bl == bl2 && (float)bd > nnratio * (float)bd2)) nd = e0 & 0xFFFF;
            }
            if (nd != dec[i]) { dec[i] = nd; s_chg[round & 1] = 1; }
        }
        __syncthreads();
        if (!s_chg[round & 1]) { converged = true; break; }
    }
    if (!converged) { if (tid == 0) nmatches[p] = INIT_UNRESOLVED; return; }
    int* owner = base; int* claim_of = min_occ;
    for (int i = tid; i < n2; i += NTH) owner[i] = -1;
    __syncthreads();
    int mine = 0;
    for (int i = tid; i < nq; i += NTH) {
        const uint32_t d = dec[i];
        int c = -1; uint8_t bin = 0xFF;
        if (d != NONE) {
            c = (int)d; mine++;
            atomicMax(&owner[c], i);
            if (check_ori && mode != 1) { bin = (uint8_t)rot_bin(P.q[i].angle, P.k2[c].angle); atomicAdd(&hist[bin], 1); }
        }
        claim_of[i] = c; bin_of[i] = bin;
    }
    if (mine) atomicAdd(&s_acc, mine);
    __syncthreads();
    for (int i = tid; i < n2; i += NTH) if (owner[i] >= 0) res[i] = owner[i];
    __syncthreads();
    if (check_ori && mode != 1) {
        // rotation consistency (:2163-2183): a keypoint claimed by a query of a rejected bin is cleared (-2: the caller NULLs the slot)
        int ind1, ind2, ind3;
        three_maxima(hist, ind1, ind2, ind3);
        int culled = 0;
        for (int i = tid; i < nq; i += NTH) {
            const int b = bin_of[i];
            if (b != 0xFF && b != ind1 && b != ind2 && b != ind3) { res[claim_of[i]] = -2; culled++; }
        }
        if (culled) atomicAdd(&s_cull, culled);
        __syncthreads();
    }
    if (tid == 0) nmatches[p] = s_acc - s_cull;
}

__global__ void __launch_bounds__(32) k_window_resolve(WinBufs W, int mode, float nnratio, int check_ori, int max_dist,
                                                     int32_t* out, int32_t* nmatches, float* prev_xy, int only_unresolved)
{
    extern __shared__ int s_mem[];
    const int lane = threadIdx.x;
    const int p = blockIdx.x;
    orbx_pdl_prologue();
    if (only_unresolved && nmatches[p] != INIT_UNRESOLVED) return;        // k_init_resolve finished this pair
    const PairDesc P = W.pairs[p];
    const int nq = min(P.nq, W.K), n2 = min(P.n2, W.K);
    int* matchedDist = s_mem;                 // [K] (mode 2)
    int* matches21 = s_mem + W.K;             // [K] (mode 2)
    // modes 0 / 1: "occupied" (= holds a MapPoint with Observations() > 0, :89-91 / :2045-2047) is kept apart from the owner in
    // res[]: a claim by a 0-observation MapPoint (query valid bit 1) leaves the keypoint free for later queries, which may take
    // it over; every accepting query counts as a match and enters the rotation histogram on its own (:2068-2090)
    int* occ = s_mem;                         // [K] by keypoint
    int* claim_of = s_mem + W.K;              // [K] by query: the keypoint the query claimed, or -1
    int accepted = 0;
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
        for (int i = lane; i < n2; i += 32) occ[i] = res[i] >= 0;
        for (int i = lane; i < nq; i += 32) { claim_of[i] = -1; bin_of[i] = 0xFF; }
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
            taken = (d0 != 0x7fffffff && occ[e0 & 0xFFFF]) || (mode == 1 && d1 != 0x7fffffff && occ[e1 & 0xFFFF]);
        }
        if (taken) {                                     // warp-uniform: rescan the query's list with the skip rule
            const int cnt = __shfl_sync(0xffffffffu, my_cnt, src), off = __shfl_sync(0xffffffffu, my_off, src);
            int k0 = 0x7fffffff, k1 = 0x7fffffff;
            d0 = 0x7fffffff; d1 = 0x7fffffff; e0 = 0; e1 = 0;
            for (int k = lane; k < cnt; k += 32) {
                const uint32_t e = pool[off + k];
                const int i2 = e & 0xFFFF, d = (e >> 16) & 0x1FF;
                const bool skip = (mode == 2) ? (matchedDist[i2] <= d)      // ORBmatcher.cc:741
                                              : (occ[i2] != 0);             // occupied keypoint, :89-91 / :2045-2047
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
            // best starts at 256 and must be <= TH_HIGH (:2038, :2068), or the caller's bound for the other overloads
            if (d0 <= max_dist) {
                const int i2 = e0 & 0xFFFF;
                if (lane == 0) {
                    res[i2] = i; claim_of[i] = i2; accepted++;
                    if (!(P.q[i].valid & 2)) occ[i2] = 1;
                    if (check_ori) { const int bin = rot_bin(a1, a2); bin_of[i] = (uint8_t)bin; hist[bin]++; }
                }
            }
        } else {
            // :78-141: best/second start at 256 with level -1
            const int bd = d0 < 256 ? d0 : 256, bd2 = d1 < 256 ? d1 : 256;
            const int bl = d0 < 256 ? (int)(e0 >> 25) : -1, bl2 = d1 < 256 ? (int)(e1 >> 25) : -1;
            if (bd <= ORBX_TH_HIGH && !(bl == bl2 && (float)bd > nnratio * (float)bd2)) {
                if (lane == 0) { res[e0 & 0xFFFF] = i; accepted++; if (!(P.q[i].valid & 2)) occ[e0 & 0xFFFF] = 1; }
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
        // both tables are indexed by query; mode 0 clears the CLAIMED keypoint (a keypoint claimed twice is cleared when either claim
        // falls into a rejected bin, exactly as the reference's rotHist lists do)
        int culled = 0;
#pragma unroll 4
        for (int i = lane; i < nq; i += 32) {
            const int b = bin_of[i];
            if (b != 0xFF && b != ind1 && b != ind2 && b != ind3) {
                if (mode == 2) res[i] = -1; else { res[claim_of[i]] = -2; culled++; }      // -2: claimed, then cleared (the caller NULLs the slot)
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) culled += __shfl_xor_sync(0xffffffffu, culled, o);
        accepted -= culled;                               // only lane 0's value is used
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
    if (lane == 0) nmatches[p] = mode == 2 ? cntm : accepted;
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
static int m_alloc(orbx_matcher* m, void** p, size_t bytes)
{
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) { orbx_set_error("%s: %s", "cudaMalloc", cudaGetErrorString(e)); return ORBX_E_NOMEM; }
    m->allocs.push_back(*p);
    if (cudaMemset(*p, 0, bytes) != cudaSuccess) { orbx_set_error("%s%s", "cudaMemset failed", ""); return ORBX_E_CUDA; }
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
    if (cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) != cudaSuccess) {
        orbx_set_error("%s%s", "cudaStreamCreate failed", ""); delete m; return ORBX_E_CUDA;
    }
    WinBufs& W = m->W;
    memset(&W, 0, sizeof(W));
    W.K = m->K; W.POOL = m->POOL;
    const size_t K = m->K, P = m->P;
    int rc;
#define MA(ptr, bytes) if ((rc = m_alloc(m, (void**)&(ptr), (bytes)))) { orbx_matcher_destroy(m); return rc; }
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
    m->d_pipe_knn = nullptr; m->pipe_knn_elems = 0;
    m->d_gen = nullptr; m->gen_bytes = 0;
    m->d_kps_src = nullptr;
    m->cam_set = false; m->cam_ndist = 0; m->d_kps_un = nullptr; m->kps_un_elems = 0;
    m->d_st = nullptr; m->st_bytes = 0;
    m->h_mono2 = nullptr; m->mono2_cap = 0;
    m->s_h2d = m->s_d2h = nullptr;
    m->h_err = nullptr;
#define CKD(call)                                                                         \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            orbx_set_error("%s failed: %s", #call, cudaGetErrorString(e_));               \
            orbx_matcher_destroy(m);                                                      \
            return ORBX_E_CUDA;                                                           \
        }                                                                                 \
    } while (0)
    CKD(cudaMemset(W.err, 0, sizeof(unsigned)));
    CKD(cudaMallocHost((void**)&m->h_err, sizeof(unsigned)));
    CKD(ORBX_OPTIN_SMEM(k_window_resolve));
    CKD(ORBX_OPTIN_SMEM(k_grid_build));
#undef CKD
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
    if (m->d_pipe_knn) cudaFree(m->d_pipe_knn);
    if (m->d_gen) cudaFree(m->d_gen);
    if (m->d_kps_un) cudaFree(m->d_kps_un);
    if (m->d_st) cudaFree(m->d_st);
    if (m->h_mono2) cudaFreeHost(m->h_mono2);
    if (m->s_h2d) {
        cudaStreamDestroy(m->s_h2d); cudaStreamDestroy(m->s_d2h); cudaStreamDestroy(m->s_match);
        if (m->st_init) for (int i = 0; i < 2; i++) { cudaEventDestroy(m->st[i].ev_kernels); cudaEventDestroy(m->st[i].ev_host); cudaFreeHost(m->st[i].h_err); }
        for (int k = 0; k < 2; k++) { if (m->lg.exec[k]) cudaGraphExecDestroy(m->lg.exec[k]); if (m->lg.graph[k]) cudaGraphDestroy(m->lg.graph[k]); }
        if (m->h_mail) { cudaFreeHost(m->h_mail); cudaEventDestroy(m->ev_ex); cudaEventDestroy(m->ev_done); cudaEventDestroy(m->ev_exd2h); cudaEventDestroy(m->ev_carry); }
        for (int i = 0; i < ORBX_MAX_CHUNKS; i++) cudaEventDestroy(m->ev_ext[i]);
        for (int i = 0; i < 2 * ORBX_MAX_CHUNKS; i++) { cudaEventDestroy(m->ev[i]); cudaEventDestroy(m->ev_r[i]); }
        cudaEventDestroy(m->ev_start);
    }
    if (m->s_bf) { cudaStreamDestroy(m->s_bf); cudaEventDestroy(m->ev_bf_fork); cudaEventDestroy(m->ev_bf_join); }
    if (m->d_pack) { cudaFree(m->d_pack); cudaFreeHost(m->h_pack); }
    if (m->h_gen) cudaFreeHost(m->h_gen);
    if (m->h_err) cudaFreeHost(m->h_err);
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
    static const bool use_popc = getenv("ORBX_BF_POPC") != nullptr;      // the popc-pipe kernel of round 1 (kept for comparison runs)
    if (npairs == 0 || nq_max <= 0) return ORBX_OK;
    const int want = 148 * 8;
    int qblocks, nsplit = 1; long long chunk;
    if (!use_popc) {
        // tensor-core kernel: 256 queries x 128 train rows per tile, one CTA per SM; split the train set so that about 8 CTAs per
        // SM exist, but keep at least 4 tiles per CTA (the query tile and the TMEM allocation are per CTA)
        qblocks = (nq_max + bftc::MQ - 1) / bftc::MQ;
        const long long tiles = (nt_max + bftc::N - 1) / bftc::N;
        const long long per = (long long)qblocks * npairs;
        if (per < want && tiles > 4) {
            long long ns = (want + per - 1) / per;
            if (ns > tiles / 4) ns = tiles / 4;
            if (ns > 4096) ns = 4096;
            nsplit = (int)(ns < 1 ? 1 : ns);
        }
        const long long tpc = (tiles + nsplit - 1) / nsplit;              // tiles per CTA
        chunk = (tpc < 1 ? 1 : tpc) * bftc::N;
        nsplit = (int)((nt_max + chunk - 1) / chunk); if (nsplit < 1) nsplit = 1;
    } else {
        // split the train set so that a small query set still fills the 148 SMs
        qblocks = (nq_max + BF_NT - 1) / BF_NT;
        const long long tiles = (nt_max + BF_TILE - 1) / BF_TILE;
        if (tiles > 0)
            while ((long long)qblocks * npairs * nsplit < want && nsplit * 2 <= tiles && nsplit < 1024) nsplit *= 2;
        chunk = (nt_max + nsplit - 1) / nsplit;
        chunk = (chunk + BF_TILE - 1) / BF_TILE * BF_TILE;
        if (chunk < BF_TILE) chunk = BF_TILE;
    }
    A.chunk = chunk; A.nsplit = nsplit;
    if (nsplit > 1) {
        int rc = ensure_parts(m, (size_t)npairs * nsplit * A.out_stride * 2);
        if (rc) return rc;
        A.part_idx = m->d_part_idx; A.part_dist = m->d_part_dist;
    }
    dim3 grid(qblocks, nsplit, npairs);
    static const bool lockstep = getenv("ORBX_BF_LOCKSTEP") != nullptr;   // the first tensor-core kernel (all warps do everything), for comparison
    if (!use_popc && !lockstep) {
        CKM(ORBX_OPTIN_SMEM(k_bf_knn2_tcws));
        orbx_launch_pdl(k_bf_knn2_tcws, grid, dim3(bftc::ws::NT), (size_t)bftc::ws::SMEM_BYTES, s, A); ORBX_COUNT_LAUNCH(1);
    } else if (!use_popc) {
        CKM(ORBX_OPTIN_SMEM(k_bf_knn2_tc));
        k_bf_knn2_tc<<<grid, bftc::NT, bftc::SMEM_BYTES, s>>>(A); ORBX_COUNT_LAUNCH(1);
    } else {
        k_bf_knn2<<<grid, BF_NT, 0, s>>>(A); ORBX_COUNT_LAUNCH(1);
    }
    if (nsplit > 1) {
        dim3 mg((nq_max + 127) / 128, npairs);
        orbx_launch_pdl(k_knn2_merge, mg, dim3(128), 0, s, A.part_idx, A.part_dist, nsplit, A.out_stride, A.nq, A.q ? nullptr : A.n, A.a, A.idx, A.dist); ORBX_COUNT_LAUNCH(1);
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
    W.qminX = W.minX; W.qminY = W.minY;
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

// mode 3: the independent best of every query is the first entry of its top-2 record
__global__ void k_best_from_top2(WinBufs W, int nq, int32_t* best_idx, int32_t* best_dist)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const unsigned e = W.q_cnt[i] > 0 ? W.top2[i].x : 0xFFFFFFFFu;
    best_idx[i] = e == 0xFFFFFFFFu ? -1 : (int)(e & 0xFFFF);
    best_dist[i] = e == 0xFFFFFFFFu ? 256 : (int)((e >> 16) & 0x1FF);
}

// ---- k_rig_resolve: modes 0 / 1 on a two-camera frame (Frame::Nleft != -1; R/src/ORBmatcher.cc:144-213, :2093-2160) ----
// Pair 0 = the left camera's keypoints [0, nL), pair 1 = the right camera's [nL, nL + nR), each with its own grid and candidate
// pool; map point i has a left query (pair 0, query i) and a right query (pair 1, query i).  The reference walks the points in
// order - left of i, right of i, left of i+1, ... - over ONE occupancy table, so one warp does the same: 32 lanes scan a
// query's candidate list, lane 0 applies the rule.  res / occ are indexed by the combined keypoint index.
__global__ void __launch_bounds__(32) k_rig_resolve(WinBufs W, int mode, float nnratio, int check_ori, int max_dist, int nL, int nR, int nq,
                                                   const int32_t* l2r, const int32_t* r2l, int32_t* res, int32_t* nmatches)
{
    extern __shared__ int s_mem[];
    const int lane = threadIdx.x;
    int* occ = s_mem;                          // [nL + nR] by combined keypoint
    int* claim_of = occ + nL + nR;             // [2][nq] by (side, query): the combined keypoint the query claimed, or -1
    uint8_t* bin_of = reinterpret_cast<uint8_t*>(claim_of + 2 * nq);   // [2][nq]
    __shared__ int hist[ORBX_HISTO_LENGTH];
    if (lane < ORBX_HISTO_LENGTH) hist[lane] = 0;
    for (int i = lane; i < nL + nR; i += 32) occ[i] = res[i] >= 0;
    for (int i = lane; i < 2 * nq; i += 32) { claim_of[i] = -1; bin_of[i] = 0xFF; }
    __syncwarp();
    int accepted = 0;
    for (int i = 0; i < nq; i++) {
        bool skip_right = false;
        for (int side = 0; side < 2; side++) {
            if (side == 1 && skip_right) break;
            const PairDesc P = W.pairs[side];
            const int cnt = W.q_cnt[(long long)side * W.K + i];
            if (cnt <= 0) {
                // mode 0: a left search that takes part but finds its window empty leaves the point (`continue` of :2033-2034)
                if (mode == 0 && side == 0 && (P.q[i].valid & 1)) skip_right = true;
                continue;
            }
            const int off = W.q_off[(long long)side * W.K + i];
            const uint32_t* pool = W.pool + (long long)side * W.POOL;
            const int kbase = side ? nL : 0;
            int d0 = 0x7fffffff, d1 = 0x7fffffff, k0 = 0x7fffffff, k1 = 0x7fffffff;
            uint32_t e0 = 0, e1 = 0;
            for (int k = lane; k < cnt; k += 32) {
                const uint32_t e = pool[off + k];
                const int i2 = e & 0xFFFF, d = (e >> 16) & 0x1FF;
                if (occ[kbase + i2]) continue;                                 // :89-91 / :159-161 / :2045-2047 / :2117-2119
                if (d < d0) { d1 = d0; k1 = k0; e1 = e0; d0 = d; k0 = k; e0 = e; }
                else if (d < d1) { d1 = d; k1 = k; e1 = e; }
            }
            warp_top2(d0, k0, e0, d1, e1, k1);
            __syncwarp();                                                      // every lane's reads of occ[] precede lane 0's writes below
            const int obs = !(P.q[i].valid & 2);                              // the claiming MapPoint has observations: it occupies what it takes
            if (mode == 1) {
                const int bd = d0 < 256 ? d0 : 256, bd2 = d1 < 256 ? d1 : 256;
                const int bl = d0 < 256 ? (int)(e0 >> 25) : -1, bl2 = d1 < 256 ? (int)(e1 >> 25) : -1;
                if (bd <= ORBX_TH_HIGH) {
                    if (bl == bl2 && (float)bd > nnratio * (float)bd2) { if (side == 0) skip_right = true; continue; }   // `continue` of :116-117 leaves the point
                    const int i2 = e0 & 0xFFFF;
                    if (lane == 0) {
                        const int32_t* partner = side ? r2l : l2r;
                        const int pi = partner ? partner[i2] : -1;
                        if (pi >= 0) { const int pk = side ? pi : nL + pi; res[pk] = i; occ[pk] = obs; accepted++; }      // :123-127 / :191-195
                        res[kbase + i2] = i; occ[kbase + i2] = obs; accepted++;
                    }
                }
            } else {
                // left: :2038-2090 (best <= TH_HIGH or the caller's bound); right: :2106-2155 (the best alone)
                if (d0 <= max_dist) {
                    const int i2 = e0 & 0xFFFF;
                    if (lane == 0) {
                        res[kbase + i2] = i; claim_of[side * nq + i] = kbase + i2; accepted++;
                        if (obs) occ[kbase + i2] = 1;
                        if (check_ori) { const int bin = rot_bin(P.q[i].angle, P.k2[i2].angle); bin_of[side * nq + i] = (uint8_t)bin; hist[bin]++; }
                    }
                }
            }
            __syncwarp();
        }
        __syncwarp();
    }
    __syncwarp();
    if (check_ori && mode == 0) {                                             // :2163-2183
        int ind1, ind2, ind3;
        three_maxima(hist, ind1, ind2, ind3);
        int culled = 0;
        for (int i = lane; i < 2 * nq; i += 32) {
            const int b = bin_of[i];
            if (b != 0xFF && b != ind1 && b != ind2 && b != ind3) { res[claim_of[i]] = -2; culled++; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) culled += __shfl_xor_sync(0xffffffffu, culled, o);
        accepted -= culled;
    }
    if (lane == 0) *nmatches = accepted;
}

static int run_window(orbx_matcher* m, const WinBufs& W, int npairs, int nq_max, int mode, float nnratio, int check_ori,
                      int32_t* d_out, int32_t* d_nm, float* d_prev, cudaStream_t s, int max_dist = ORBX_TH_HIGH)
{
    int npad = 1; while (npad < m->K) npad <<= 1;
    CKM(ORBX_OPTIN_SMEM(k_grid_build));
    orbx_launch_pdl(k_grid_build, dim3(npairs), dim3(GRID_NT), sizeof(int) * (NCELL + 2) + 2 * sizeof(unsigned short) * (size_t)m->K, s, W, m->K, mode == 2 ? 0 : -1); ORBX_COUNT_LAUNCH(1);   // also empties the pairs' candidate pools
    dim3 cg((nq_max + CAND_WARPS - 1) / CAND_WARPS, npairs);
    if (nq_max > 0) { orbx_launch_pdl(k_window_candidates, cg, dim3(CAND_WARPS * 32), 0, s, W); ORBX_COUNT_LAUNCH(1); }
    if (mode == 3) { CKM(cudaGetLastError()); return ORBX_OK; }
    CKM(ORBX_OPTIN_SMEM(k_window_resolve));
    int only_unresolved = 0;
    if (mode == 2 && m->K <= INIT_MAX_K && !getenv("ORBX_SEQ_RESOLVE")) {
        // parallel fixed-point resolve; pairs it cannot finish are marked and fall through to the sequential kernel below
        CKM(ORBX_OPTIN_SMEM(k_init_resolve));
        orbx_launch_pdl(k_init_resolve, dim3(npairs), dim3(npairs <= 32 ? INIT_NT_FEW : INIT_NT), (2 + INIT_CLAIMS) * m->K * sizeof(int), s, W, nnratio, check_ori, d_out, d_nm, d_prev); ORBX_COUNT_LAUNCH(1);
        only_unresolved = 1;
    } else if ((mode == 0 || mode == 1) && m->K <= PROJ_MAX_K && !getenv("ORBX_SEQ_RESOLVE")) {
        CKM(ORBX_OPTIN_SMEM(k_proj_resolve));
        orbx_launch_pdl(k_proj_resolve, dim3(npairs), dim3(npairs <= 32 ? INIT_NT_FEW : INIT_NT), 3 * m->K * sizeof(int), s, W, mode, nnratio, check_ori, max_dist, d_out, d_nm); ORBX_COUNT_LAUNCH(1);
        only_unresolved = 1;
    }
    orbx_launch_pdl(k_window_resolve, dim3(npairs), dim3(32), 2 * m->K * sizeof(int), s, W, mode, nnratio, check_ori, max_dist, d_out, d_nm, d_prev, only_unresolved); ORBX_COUNT_LAUNCH(1);
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

extern "C" int orbx_search_by_projection_ex(orbx_matcher* m, int mode, const orbx_proj_query* q, const uint8_t* qdesc, int nq,
                                            const orbx_keypoint* k2, const uint8_t* d2, const float* uright2, int n2,
                                            const float bounds[4], int32_t* assigned, float nnratio, int check_ori, int max_dist,
                                            const float* inv_level_sigma2, int nlevels, double chi2,
                                            int32_t* best_idx, int32_t* best_dist, int* nmatches)
{
    if (!bounds || (chi2 > 0 && (!inv_level_sigma2 || nlevels < 1 || nlevels > ORBX_MAX_LEVELS))) {
        orbx_set_error("%s%s", "orbx_search_by_projection: invalid arguments", "");
        return ORBX_E_INVALID;
    }
    orbx_proj_options o;
    memset(&o, 0, sizeof(o));
    for (int i = 0; i < 4; i++) o.bounds[i] = bounds[i];
    o.query_origin[0] = bounds[0]; o.query_origin[1] = bounds[2];
    o.nnratio = nnratio; o.check_ori = check_ori; o.max_dist = max_dist;
    o.nlevels = chi2 > 0 ? nlevels : 0;
    for (int l = 0; l < o.nlevels; l++) o.inv_level_sigma2[l] = inv_level_sigma2[l];
    o.chi2_mono = chi2; o.chi2_stereo = 0.0;
    return orbx_search_by_projection_opts(m, mode, q, qdesc, nq, k2, d2, uright2, n2, &o, assigned, best_idx, best_dist, nmatches);
}

extern "C" int orbx_search_by_projection_opts(orbx_matcher* m, int mode, const orbx_proj_query* q, const uint8_t* qdesc, int nq,
                                              const orbx_keypoint* k2, const uint8_t* d2, const float* uright2, int n2,
                                              const orbx_proj_options* opt, int32_t* assigned,
                                              int32_t* best_idx, int32_t* best_dist, int* nmatches)
{
    const bool indep = mode == 3;
    if (!m || !opt || (mode != 0 && mode != 1 && mode != 3) || nq < 0 || n2 < 0 || nq > m->K || n2 > m->K ||
        (nq > 0 && (!q || !qdesc)) || (n2 > 0 && (!k2 || !d2)) || (!indep && n2 > 0 && !assigned) || (indep && nq > 0 && (!best_idx || !best_dist)) ||
        (opt->chi2_mono > 0 && (opt->nlevels < 1 || opt->nlevels > ORBX_MAX_LEVELS)) || (opt->chi2_stereo > 0 && !(opt->chi2_mono > 0))) {
        orbx_set_error("%s%s", "orbx_search_by_projection: invalid arguments / more keypoints than max_keypoints", "");
        return ORBX_E_INVALID;
    }
    const float* bounds = opt->bounds;
    const float nnratio = opt->nnratio; const int check_ori = opt->check_ori, max_dist = opt->max_dist;
    const double chi2 = opt->chi2_mono; const int nlevels = opt->nlevels; const float* inv_level_sigma2 = opt->inv_level_sigma2;
    if (nmatches) *nmatches = 0;
    if (indep) for (int i = 0; i < nq; i++) { best_idx[i] = -1; best_dist[i] = 256; }
    if (nq == 0 || n2 == 0) return ORBX_OK;
    CKM(cudaSetDevice(m->p.device));
    cudaStream_t s = m->stream;
    set_bounds(m, bounds);
    m->W.qminX = opt->query_origin[0]; m->W.qminY = opt->query_origin[1];
    m->W.gate_chi2 = chi2 > 0 ? chi2 : 0.0;
    m->W.gate_chi2_stereo = opt->chi2_stereo > 0 ? opt->chi2_stereo : 0.0;
    for (int l = 0; l < ORBX_MAX_LEVELS; l++) m->W.inv_sigma2[l] = (chi2 > 0 && l < nlevels) ? inv_level_sigma2[l] : 0.f;
    // The caller's arrays are pageable (std::vector / cv::Mat in the class layer): seven separate copies cost ~8 us each.  They are
    // packed into ONE pinned block (host memcpy: ~120 KB, a few us) that travels with one copy; the pair descriptor, the occupancy
    // table (in / out) and the match count live in the same block, so the results come back with one copy too.
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t o_q = 0, o_qd = o_q + al(sizeof(orbx_proj_query) * nq), o_k2 = o_qd + al((size_t)32 * nq), o_d2 = o_k2 + al(sizeof(orbx_keypoint) * n2),
                 o_ur = o_d2 + al((size_t)32 * n2), o_pd = o_ur + al(sizeof(float) * n2), o_as = o_pd + al(sizeof(PairDesc)),
                 o_nm = o_as + al(sizeof(int32_t) * n2), total = o_nm + 256;
    if (total > m->pack_bytes) {
        if (m->d_pack) { CKM(cudaStreamSynchronize(s)); cudaFree(m->d_pack); cudaFreeHost(m->h_pack); m->d_pack = nullptr; m->h_pack = nullptr; m->pack_bytes = 0; }
        const size_t cap_bytes = total + total / 2;
        CKM(cudaMalloc((void**)&m->d_pack, cap_bytes));
        if (cudaMallocHost((void**)&m->h_pack, cap_bytes) != cudaSuccess) { cudaFree(m->d_pack); m->d_pack = nullptr; orbx_set_error("%s%s", "cudaMallocHost failed", ""); return ORBX_E_NOMEM; }
        m->pack_bytes = cap_bytes;
    }
    uint8_t* hp = m->h_pack; uint8_t* dp = m->d_pack;
    memcpy(hp + o_q, q, sizeof(orbx_proj_query) * nq);
    memcpy(hp + o_qd, qdesc, (size_t)32 * nq);
    memcpy(hp + o_k2, k2, sizeof(orbx_keypoint) * n2);
    memcpy(hp + o_d2, d2, (size_t)32 * n2);
    if (uright2) memcpy(hp + o_ur, uright2, sizeof(float) * n2);
    if (!indep) memcpy(hp + o_as, assigned, sizeof(int32_t) * n2);
    PairDesc pd{};
    pd.k1 = nullptr; pd.d1 = nullptr; pd.k2 = reinterpret_cast<const orbx_keypoint*>(dp + o_k2); pd.d2 = dp + o_d2;
    pd.uright2 = uright2 ? reinterpret_cast<const float*>(dp + o_ur) : nullptr;
    pd.q = reinterpret_cast<orbx_proj_query*>(dp + o_q); pd.qdesc = dp + o_qd; pd.n1 = nq; pd.n2 = n2; pd.nq = nq;
    memcpy(hp + o_pd, &pd, sizeof(pd));
    CKM(cudaMemcpyAsync(dp, hp, o_nm, cudaMemcpyHostToDevice, s));
    OrbxPdlScope pdl_scope(true);                      // one pair: a chain of small kernels (programmatic dependent launch, orbx_internal.h)
    WinBufs W = m->W;
    W.pairs = reinterpret_cast<PairDesc*>(dp + o_pd);
    int32_t* d_res = reinterpret_cast<int32_t*>(dp + o_as); int32_t* d_cnt = reinterpret_cast<int32_t*>(dp + o_nm);
    int rc = run_window(m, W, 1, nq, mode, nnratio, check_ori, d_res, d_cnt, nullptr, s, max_dist);
    m->W.gate_chi2 = 0.0; m->W.gate_chi2_stereo = 0.0;
    m->W.qminX = m->W.minX; m->W.qminY = m->W.minY;
    if (rc) return rc;
    int nm = 0;
    if (indep) {
        k_best_from_top2<<<(nq + 255) / 256, 256, 0, s>>>(W, nq, m->d_knn_idx, m->d_knn_dist); ORBX_COUNT_LAUNCH(1);
        CKM(cudaMemcpyAsync(best_idx, m->d_knn_idx, sizeof(int32_t) * nq, cudaMemcpyDeviceToHost, s));
        CKM(cudaMemcpyAsync(best_dist, m->d_knn_dist, sizeof(int32_t) * nq, cudaMemcpyDeviceToHost, s));
        rc = m_check_err(m, s);
        if (rc) return rc;
        for (int i = 0; i < nq; i++) nm += best_idx[i] >= 0;
    } else {
        CKM(cudaMemcpyAsync(hp + o_as, dp + o_as, o_nm + sizeof(int32_t) - o_as, cudaMemcpyDeviceToHost, s));
        rc = m_check_err(m, s);
        if (rc) return rc;
        memcpy(assigned, hp + o_as, sizeof(int32_t) * n2);
        nm = *reinterpret_cast<const int32_t*>(hp + o_nm);
    }
    if (nmatches) *nmatches = nm;
    return ORBX_OK;
}

extern "C" int orbx_search_by_projection_rig(orbx_matcher* m, int mode, const orbx_proj_query* ql, const orbx_proj_query* qr,
                                             const uint8_t* qdesc, int nq, const orbx_keypoint* k2, const uint8_t* d2, int n_left, int n_right,
                                             const int32_t* l2r, const int32_t* r2l, const orbx_proj_options* opt, int32_t* assigned,
                                             int* nmatches)
{
    const int n2 = n_left + n_right;
    if (!m || !opt || (mode != 0 && mode != 1) || nq < 0 || n_left < 0 || n_right < 0 || nq > m->K || n2 > m->K || m->P < 2 ||
        (nq > 0 && (!ql || !qr || !qdesc)) || (n2 > 0 && (!k2 || !d2 || !assigned)) || (size_t)(n2 + 3 * nq) * sizeof(int) > 200 * 1024) {
        orbx_set_error("%s%s", "orbx_search_by_projection_rig: invalid arguments (needs max_batch >= 2, n_left + n_right <= max_keypoints)", "");
        return ORBX_E_INVALID;
    }
    if (nmatches) *nmatches = 0;
    if (nq == 0 || n2 == 0) return ORBX_OK;
    CKM(cudaSetDevice(m->p.device));
    cudaStream_t s = m->stream;
    set_bounds(m, opt->bounds);
    m->W.qminX = opt->query_origin[0]; m->W.qminY = opt->query_origin[1];
    // combined keypoints / descriptors in the single-frame staging; pair 0 reads the left half, pair 1 the right half
    CKM(cudaMemcpyAsync(m->W.q, ql, sizeof(orbx_proj_query) * nq, cudaMemcpyHostToDevice, s));
    CKM(cudaMemcpyAsync(m->W.q + m->K, qr, sizeof(orbx_proj_query) * nq, cudaMemcpyHostToDevice, s));
    CKM(cudaMemcpyAsync(m->d_qdesc, qdesc, (size_t)32 * nq, cudaMemcpyHostToDevice, s));
    CKM(cudaMemcpyAsync(m->d_k2, k2, sizeof(orbx_keypoint) * n2, cudaMemcpyHostToDevice, s));
    CKM(cudaMemcpyAsync(m->d_d2, d2, (size_t)32 * n2, cudaMemcpyHostToDevice, s));
    CKM(cudaMemcpyAsync(m->d_out, assigned, sizeof(int32_t) * n2, cudaMemcpyHostToDevice, s));
    int32_t* d_partner = nullptr;
    if (l2r || r2l) {
        int rc = orbx_m_gen_scratch(m, sizeof(int32_t) * (size_t)(n2 > 0 ? n2 : 1));
        if (rc) return rc;
        d_partner = reinterpret_cast<int32_t*>(m->d_gen);
        if (l2r && n_left) CKM(cudaMemcpyAsync(d_partner, l2r, sizeof(int32_t) * n_left, cudaMemcpyHostToDevice, s));
        if (r2l && n_right) CKM(cudaMemcpyAsync(d_partner + n_left, r2l, sizeof(int32_t) * n_right, cudaMemcpyHostToDevice, s));
    }
    PairDesc pd[2] = {};
    for (int side = 0; side < 2; side++) {
        pd[side].k2 = m->d_k2 + (side ? n_left : 0); pd[side].d2 = m->d_d2 + (size_t)(side ? n_left : 0) * 32; pd[side].uright2 = nullptr;
        pd[side].q = m->W.q + (size_t)side * m->K; pd[side].qdesc = m->d_qdesc; pd[side].n1 = nq; pd[side].n2 = side ? n_right : n_left; pd[side].nq = nq;
    }
    CKM(cudaMemcpyAsync(m->W.pairs, pd, sizeof(pd), cudaMemcpyHostToDevice, s));
    int rc = run_window(m, m->W, 2, nq, 3, opt->nnratio, opt->check_ori, nullptr, nullptr, nullptr, s);      // grids + candidate pools of both halves
    m->W.qminX = m->W.minX; m->W.qminY = m->W.minY;
    if (rc) return rc;
    CKM(ORBX_OPTIN_SMEM(k_rig_resolve));
    const size_t smem = (size_t)(n2 + 2 * nq) * sizeof(int) + 2 * (size_t)nq;
    k_rig_resolve<<<1, 32, smem, s>>>(m->W, mode, opt->nnratio, opt->check_ori, opt->max_dist > 0 ? opt->max_dist : ORBX_TH_HIGH, n_left, n_right, nq,
                                      l2r ? d_partner : nullptr, r2l ? d_partner + n_left : nullptr, m->d_out, m->d_nm); ORBX_COUNT_LAUNCH(1);
    CKM(cudaGetLastError());
    int nm = 0;
    CKM(cudaMemcpyAsync(assigned, m->d_out, sizeof(int32_t) * n2, cudaMemcpyDeviceToHost, s));
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
    if (mode != 0 && mode != 1) return ORBX_E_INVALID;
    return orbx_search_by_projection_ex(m, mode, q, qdesc, nq, k2, d2, uright2, n2, bounds, assigned, nnratio, check_ori, ORBX_TH_HIGH,
                                        nullptr, 0, 0.0, nullptr, nullptr, nmatches);
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
    // mvKeysUn when the caller supplied undistorted keypoints (same [slot][cap] layout), mvKeys otherwise
    // The brute-force kNN-2 does not depend on the window search (both only read the slots).  With a few pairs each kernel is a
    // few CTAs and the two chains run side by side (same policy as the extractor's run_batch_dag: under stream capture, where the
    // fork / join are graph edges; ORBX_DAG=1 forces it, 0 disables it).  With a full batch either kernel fills the GPU and a
    // fork gains nothing (measured: 3.424 vs 3.420 ms per 512 frames).
    static const int dag_env = getenv("ORBX_DAG") ? atoi(getenv("ORBX_DAG")) : -1;
    OrbxPdlScope pdl_scope(npairs <= 2);
    bool fork = false;
    if (d_knn_idx && d_knn_dist && npairs <= 2 && dag_env != 0) {
        if (!m->s_bf) {
            CKM(cudaStreamCreateWithFlags(&m->s_bf, cudaStreamNonBlocking));
            CKM(cudaEventCreateWithFlags(&m->ev_bf_fork, cudaEventDisableTiming));
            CKM(cudaEventCreateWithFlags(&m->ev_bf_join, cudaEventDisableTiming));
        }
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (dag_env != 1) CKM(cudaStreamIsCapturing(s, &cs));
        fork = dag_env == 1 || cs == cudaStreamCaptureStatusActive;
    }
    BfArgs A{};
    A.q = nullptr; A.desc = desc; A.n = n; A.a = a; A.b = b; A.cap = cap;
    A.idx = d_knn_idx; A.dist = d_knn_dist; A.out_stride = m->K; A.idx_base = 0;
    if (fork) {
        CKM(cudaEventRecord(m->ev_bf_fork, s));
        CKM(cudaStreamWaitEvent(m->s_bf, m->ev_bf_fork, 0));
        rc = bf_launch(m, A, npairs, cap, cap, m->s_bf);
        if (rc) return rc;
        CKM(cudaEventRecord(m->ev_bf_join, m->s_bf));
    }
    orbx_launch_pdl(k_setup_slot_pairs, dim3(npairs), dim3(256), 0, s, W, m->d_kps_src ? m->d_kps_src : kps, desc, n, a, b, cap, (float)window); ORBX_COUNT_LAUNCH(1);
    rc = run_window(m, W, npairs, cap, 2, nnratio, check_ori, d_matches12, d_nmatches, nullptr, s);
    if (rc) return rc;
    if (fork) CKM(cudaStreamWaitEvent(s, m->ev_bf_join, 0));
    else if (d_knn_idx && d_knn_dist) {
        rc = bf_launch(m, A, npairs, cap, cap, s);
        if (rc) return rc;
    }
    return ORBX_OK;
}

// Camera of the stream pipelines (orbx_extract_match_batch*): with distortion coefficients set, every chunk's keypoints
// are undistorted on the device (Frame::UndistortKeyPoints) right after extraction and the matching runs on mvKeysUn, as
// the reference does.  K = NULL clears it.
extern "C" int orbx_matcher_set_camera(orbx_matcher* m, const float* K, const float* dist, int ndist, const float* P)
{
    if (!m) return ORBX_E_INVALID;
    if (!K) { m->cam_set = false; return ORBX_OK; }
    if (!dist || !P || ndist < 4 || ndist > 12) return ORBX_E_INVALID;
    for (int i = 0; i < 9; i++) { m->cam_K[i] = K[i]; m->cam_P[i] = P[i]; }
    for (int i = 0; i < 12; i++) m->cam_dist[i] = i < ndist ? dist[i] : 0.f;
    m->cam_ndist = ndist;
    m->cam_set = dist[0] != 0.0f;                       // R/src/Frame.cc:723-727: mvKeysUn = mvKeys without distortion
    return ORBX_OK;
}

// device view of mvKeysUn of the last pipeline call ([slots][orbx_extractor_max_keypoints], slot i + 1 = frame i); NULL when
// no camera is set
extern "C" int orbx_matcher_undistorted_device(orbx_matcher* m, orbx_keypoint** d_kps_un)
{
    if (!m || !d_kps_un) return ORBX_E_INVALID;
    *d_kps_un = m->cam_set ? m->d_kps_un : nullptr;
    return ORBX_OK;
}

// Undistorted keypoints for the slot-based searches: a DEVICE array laid out like the extractor's results
// ([slots][orbx_extractor_max_keypoints]), e.g. the output of orbx_undistort_slots_device; NULL = use the extractor's own.
extern "C" int orbx_matcher_set_slot_keypoints(orbx_matcher* m, const orbx_keypoint* d_kps_un)
{
    if (!m) return ORBX_E_INVALID;
    m->d_kps_src = d_kps_un;
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

// scratch block of the candidate-list searches (orbx_search.cu); all users synchronise their stream before returning
int orbx_m_gen_scratch(orbx_matcher* m, size_t bytes)
{
    if (bytes <= m->gen_bytes) return ORBX_OK;
    if (m->d_gen) cudaFree(m->d_gen);
    m->d_gen = nullptr; m->gen_bytes = 0;
    CKM(cudaMalloc((void**)&m->d_gen, bytes));
    m->gen_bytes = bytes;
    return ORBX_OK;
}

int orbx_m_ensure_pipeline(orbx_matcher* m)
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
static int extract_match_pipeline_impl(orbx_extractor* ex, orbx_matcher* m, bool host, const uint8_t* imgs, int batch, int width,
                                  int height, int stride, size_t frame_stride, int lap0, int lap1,
                                  const float bounds[4], int window, float nnratio, int check_ori,
                                  orbx_keypoint* kps, uint8_t* desc, int cap, int32_t* n, int32_t* mono_index,
                                  int32_t* matches12, int32_t* nmatches, int32_t* knn_idx, int32_t* knn_dist,
                                  int32_t* d_matches12, int32_t* d_nmatches, int32_t* d_knn_idx, int32_t* d_knn_dist,
                                  cudaStream_t s)
{
    int rc = orbx_ex_configure(ex, width, height);
    if (rc) return rc;
    CKM(cudaSetDevice(m->p.device));
    if ((rc = orbx_m_ensure_pipeline(m))) return rc;
    if (host && knn_idx && knn_dist) {
        // device landing zone of the BF kNN-2 tables (rows of stride K like every matcher output)
        const size_t need = (size_t)m->P * m->K * 2;
        if (need > m->pipe_knn_elems) {
            if (m->d_pipe_knn) { CKM(cudaDeviceSynchronize()); cudaFree(m->d_pipe_knn); m->d_pipe_knn = nullptr; m->pipe_knn_elems = 0; }
            CKM(cudaMalloc((void**)&m->d_pipe_knn, sizeof(int32_t) * need * 2));
            m->pipe_knn_elems = need;
        }
        d_knn_idx = m->d_pipe_knn; d_knn_dist = m->d_pipe_knn + need;
    }
    struct SrcGuard {                                  // the pipeline's own mvKeysUn is the source only for the duration of the call
        orbx_matcher* m; const orbx_keypoint* saved;
        ~SrcGuard() { m->d_kps_src = saved; }
    } guard{m, m->d_kps_src};
    if (m->cam_set) {
        const size_t need = (size_t)(m->P + 1) * orbx_ex_out_cap(ex);
        if (need > m->kps_un_elems) {
            // a fresh buffer has no predecessor in slot 0: the first frame of the first call matches against an empty slot anyway
            if (m->d_kps_un) { CKM(cudaDeviceSynchronize()); cudaFree(m->d_kps_un); m->d_kps_un = nullptr; m->kps_un_elems = 0; }
            CKM(cudaMalloc((void**)&m->d_kps_un, sizeof(orbx_keypoint) * need));
            CKM(cudaMemset(m->d_kps_un, 0, sizeof(orbx_keypoint) * need));
            m->kps_un_elems = need;
        }
        m->d_kps_src = m->d_kps_un;
    }
    const bool direct = host && orbx_ex_can_fetch_direct(ex, kps, desc, cap, n, mono_index);
    // host path: chunks hide the PCIe copies under the kernels.  The call is H2D-bound in the middle, so what is left is
    // the fill (first chunk's copy) and the drain (last chunk's kernels + D2H): the first and last chunk are small
    // (1/16 of the batch each), the rest is split evenly.  Device path: chunking only shrinks the kernels and makes the
    // two kernel streams fight for the SMs (measured: 1 chunk 5.7 ms, 4 chunks 6.2 ms per 512 frames): one chunk.
    // a batch whose frames were prefetched (orbx_extract_match_batch_prefetch) is already on the device: its chunks only overlap the
    // result copies with the kernels
    const uint8_t* d_pref = nullptr; cudaEvent_t pref_ready = nullptr;
    const bool prefetched = host && orbx_ex_take_prefetched(ex, imgs, batch, width, height, &d_pref, &pref_ready);
    int nchunks = host ? (batch >= 64 ? (prefetched ? 3 : 6) : 1) : 1;
    if (const char* e = getenv(host ? (prefetched ? "ORBX_PREFETCH_CHUNKS" : "ORBX_HOST_CHUNKS") : "ORBX_DEVICE_CHUNKS")) { const int v = atoi(e); if (v >= 1 && v <= ORBX_MAX_CHUNKS) nchunks = v; }
    if (nchunks > batch) nchunks = batch;
    int c_f0[ORBX_MAX_CHUNKS], c_cnt[ORBX_MAX_CHUNKS];
    {
        const bool taper = host && !prefetched && nchunks >= 4 && batch >= 16 * nchunks && !getenv("ORBX_UNIFORM_CHUNKS");
        int f = 0, k = 0;
        if (prefetched && nchunks >= 3 && batch >= 64 && !getenv("ORBX_UNIFORM_CHUNKS")) {
            // input already resident: only the result copies are left to hide, so the chunks shrink towards the end (the drain is
            // the last chunk's kernels + copies): 1/2, 1/4, ..., the last two share the rest
            int left = batch;
            for (int i = 0; i < nchunks; i++) {
                int cnt = (i + 1 < nchunks) ? left / 2 : left;
                if (cnt < 1) cnt = left;
                c_f0[k] = f; c_cnt[k++] = cnt; f += cnt; left -= cnt;
                if (left == 0) break;
            }
        } else {
        const int edge = taper ? batch / 16 : 0;
        const int mid = taper ? nchunks - 2 : nchunks, rest = batch - 2 * edge;
        if (taper) { c_f0[k] = 0; c_cnt[k++] = edge; f = edge; }
        for (int i = 0; i < mid; i++) {
            const int cnt = rest / mid + (i < rest % mid ? 1 : 0);
            c_f0[k] = f; c_cnt[k++] = cnt; f += cnt;
        }
        if (taper) { c_f0[k] = f; c_cnt[k++] = edge; }
        }
        nchunks = k;
    }
    int32_t* dm12 = host ? m->d_out : d_matches12;
    int32_t* dnm = host ? m->d_nm : d_nmatches;
    if (host && nchunks == 1 && !prefetched) {
        // One chunk (small batches, the single-frame latency path).  Order of the call:
        //   s     : H2D | extraction | ev_ex | matching + mailbox kernel | D2H matches12, kNN | ev_done | slot carry | ev_carry
        //   s_d2h :                    ev_ex -> D2H keypoints, descriptors (beside the matcher) | ev_exd2h
        // The host waits for ev_done and ev_exd2h: the carry of the last frame into slot 0 (needed by the NEXT call only) is off
        // the latency path, and the five small results (n, monoIndex, nmatches, the two error words) arrive through ONE kernel
        // that stores them into mapped pinned memory instead of five copies.
        if (!m->h_mail) {
            CKM(cudaHostAlloc((void**)&m->h_mail, sizeof(int32_t) * (2 + 3 * (size_t)m->P), cudaHostAllocMapped));
            CKM(cudaHostGetDevicePointer((void**)&m->d_mail, m->h_mail, 0));
            CKM(cudaEventCreateWithFlags(&m->ev_ex, cudaEventDisableTiming));
            CKM(cudaEventCreateWithFlags(&m->ev_done, cudaEventDisableTiming));
            CKM(cudaEventCreateWithFlags(&m->ev_exd2h, cudaEventDisableTiming));
            CKM(cudaEventCreateWithFlags(&m->ev_carry, cudaEventDisableTiming));
        }
        rc = orbx_ex_stage_input(ex, imgs, 0, batch, width, height, stride, frame_stride, s);
        if (rc) return rc;
        // The kernel launches of the step go through two CUDA graphs (extraction: in the level-parallel order of run_batch_dag;
        // matching: window search beside the brute-force search, then the mailbox kernel): the first call with a set of arguments
        // runs them directly (lazy allocations happen there), the second captures the same sequences, later calls replay them with
        // one launch each.  Copies in and out stay outside (their host pointers change from call to call).
        auto issue_part = [&](int part) -> int {
            if (part == 0) {
                int r = orbx_ex_run_staged(ex, 0, batch, lap0, lap1, 1, s);
                if (r) return r;
                if (m->cam_set) {
                    r = orbx_undistort_slots_device(ex, 1, batch, m->cam_K, m->cam_dist, m->cam_ndist, m->cam_P, m->d_kps_un + (size_t)orbx_ex_out_cap(ex), s);
                    if (r) return r;
                }
                return ORBX_OK;
            }
            int r = match_slots_impl(m, ex, m->d_pair_a, m->d_pair_b, batch, 0, bounds, window, nnratio, check_ori, dm12, dnm, d_knn_idx, d_knn_dist, s);
            if (r) return r;
            int32_t* dn = nullptr; int32_t* dmono = nullptr;
            if ((r = orbx_extractor_results_device(ex, nullptr, nullptr, &dn, &dmono, nullptr, nullptr))) return r;
            orbx_launch_pdl(k_mailbox, dim3((batch + 127) / 128), dim3(128), 0, s, (const int*)(dn + 1), (const int*)(dmono + 1), (const int*)dnm,
                            (const unsigned*)orbx_ex_err_device(ex), (const unsigned*)m->W.err, batch, m->d_mail);
            ORBX_COUNT_LAUNCH(1);
            CKM(cudaGetLastError());
            return ORBX_OK;
        };
        static const bool no_graph = getenv("ORBX_NO_GRAPH") != nullptr;
        orbx_matcher::LatGraph& G = m->lg;
        const bool same = G.seen > 0 && G.ex == ex && G.batch == batch && G.width == width && G.height == height && G.lap0 == lap0 && G.lap1 == lap1 &&
                          G.window == window && G.check_ori == check_ori && G.knn == (d_knn_idx != nullptr) && G.cam == (int)m->cam_set &&
                          G.nnratio == nnratio && memcmp(G.bounds, bounds, sizeof(G.bounds)) == 0 && G.d_knn == d_knn_idx &&
                          G.geom_gen == orbx_ex_geom_gen(ex) && G.d_kps_un == m->d_kps_un;      // the extractor's / matcher's buffers are still the captured ones
        auto drop_graphs = [&]() {
            for (int k = 0; k < 2; k++) {
                if (G.exec[k]) { cudaGraphExecDestroy(G.exec[k]); G.exec[k] = nullptr; }
                if (G.graph[k]) { cudaGraphDestroy(G.graph[k]); G.graph[k] = nullptr; }
            }
        };
        int mode = 0;                                    // 0: direct launches, 1: capture now, 2: replay
        if (no_graph || orbx_ex_profiling(ex)) mode = 0;
        else if (same && G.seen == 2 && G.exec[0] && G.exec[1]) mode = 2;
        else if (same && G.seen == 1) { drop_graphs(); mode = 1; }
        else if (!same) {
            drop_graphs();
            G.ex = ex; G.batch = batch; G.width = width; G.height = height; G.lap0 = lap0; G.lap1 = lap1; G.window = window; G.check_ori = check_ori;
            G.knn = d_knn_idx != nullptr; G.cam = (int)m->cam_set; G.nnratio = nnratio; memcpy(G.bounds, bounds, sizeof(G.bounds)); G.d_knn = d_knn_idx;
            G.geom_gen = orbx_ex_geom_gen(ex); G.d_kps_un = m->d_kps_un;
            G.seen = 1;
        }
        auto run_part = [&](int part) -> int {
            if (mode == 2) { CKM(cudaGraphLaunch(G.exec[part], s)); ORBX_COUNT_LAUNCH(G.nkernels[part]); return ORBX_OK; }
            if (mode == 1) {
                const unsigned long long before = g_orbx_launches.load(std::memory_order_relaxed);
                CKM(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
                int r = issue_part(part);
                cudaGraph_t graph = nullptr;
                const cudaError_t ce = cudaStreamEndCapture(s, &graph);
                const int captured = (int)(g_orbx_launches.load(std::memory_order_relaxed) - before);
                g_orbx_launches.fetch_sub(captured, std::memory_order_relaxed);        // nothing ran yet
                if (!r && ce == cudaSuccess && graph && cudaGraphInstantiate(&G.exec[part], graph, 0) == cudaSuccess) {
                    G.graph[part] = graph; G.nkernels[part] = captured;
                    CKM(cudaGraphLaunch(G.exec[part], s)); ORBX_COUNT_LAUNCH(captured);
                    if (part == 1) G.seen = 2;
                    return ORBX_OK;
                }
                cudaGetLastError();
                if (graph) cudaGraphDestroy(graph);
                G.exec[part] = nullptr;
                G.seen = -1; mode = 0;                                                 // not capturable here: stay on direct launches
                if (r) return r;
            }
            return issue_part(part);
        };
        if ((rc = run_part(0))) return rc;
        CKM(cudaEventRecord(m->ev_ex, s));
        CKM(cudaStreamWaitEvent(m->s_d2h, m->ev_ex, 0));
        rc = orbx_ex_fetch_async(ex, 1, batch, 0, kps, desc, cap, n, mono_index, m->s_d2h, direct, false);
        if (rc) return rc;
        CKM(cudaEventRecord(m->ev_exd2h, m->s_d2h));
        if ((rc = run_part(1))) return rc;
        if (matches12) CKM(cudaMemcpy2DAsync(matches12, sizeof(int32_t) * cap, dm12, sizeof(int32_t) * m->K, sizeof(int32_t) * (cap < m->K ? cap : m->K), batch,
                                             cudaMemcpyDeviceToHost, s));
        if (knn_idx && knn_dist) {
            const size_t wbytes = sizeof(int32_t) * 2 * (cap < m->K ? cap : m->K);
            CKM(cudaMemcpy2DAsync(knn_idx, sizeof(int32_t) * 2 * cap, d_knn_idx, sizeof(int32_t) * 2 * m->K, wbytes, batch, cudaMemcpyDeviceToHost, s));
            CKM(cudaMemcpy2DAsync(knn_dist, sizeof(int32_t) * 2 * cap, d_knn_dist, sizeof(int32_t) * 2 * m->K, wbytes, batch, cudaMemcpyDeviceToHost, s));
        }
        CKM(cudaEventRecord(m->ev_done, s));
        // the carry: after everything the caller waits for
        rc = orbx_extractor_copy_slot(ex, batch, 0, s);
        if (rc) return rc;
        if (m->cam_set) {
            const size_t capx = orbx_ex_out_cap(ex);
            CKM(cudaMemcpyAsync(m->d_kps_un, m->d_kps_un + (size_t)batch * capx, sizeof(orbx_keypoint) * capx, cudaMemcpyDeviceToDevice, s));
        }
        CKM(cudaEventRecord(m->ev_carry, s));
        CKM(cudaEventSynchronize(m->ev_done));
        CKM(cudaEventSynchronize(m->ev_exd2h));
        const int32_t* mail = m->h_mail;
        orbx_ex_set_fetched(ex, (unsigned)mail[0], mail + 2, mail + 2 + batch, batch, n, mono_index, direct);
        if (nmatches) memcpy(nmatches, mail + 2 + 2 * batch, sizeof(int32_t) * batch);
        if (mail[1]) {
            char buf[32]; snprintf(buf, sizeof(buf), "0x%x", (unsigned)mail[1]);
            orbx_set_error("matcher device capacity error flags %s%s", buf, " (raise max_candidates)");
            cudaMemsetAsync(m->W.err, 0, sizeof(unsigned), s);
            return ORBX_E_CAPACITY;
        }
        rc = orbx_ex_fetch_finish(ex, batch, kps, desc, cap, n, mono_index, direct, true);
        // the call still returns with nothing in flight (other entry points may read slot 0 on other streams): by now the carry,
        // a few microseconds of device time, has run beside the host work above
        CKM(cudaEventSynchronize(m->ev_carry));
        return rc;
    }
    // the side streams must not run ahead of work already queued on the kernel stream (previous call's carry)
    CKM(cudaEventRecord(m->ev_start, s));
    CKM(cudaStreamWaitEvent(m->s_match, m->ev_start, 0));
    if (host) {
        if (!prefetched) CKM(cudaStreamWaitEvent(m->s_h2d, m->ev_start, 0));      // (the copy stream may already carry the NEXT batch's prefetch)
        CKM(cudaStreamWaitEvent(m->s_d2h, m->ev_start, 0));
        if (prefetched) CKM(cudaStreamWaitEvent(s, pref_ready, 0));
        for (int c = 0; c < nchunks && !prefetched; c++) {
            const int f0 = c_f0[c], cnt = c_cnt[c];
            rc = orbx_ex_stage_input(ex, imgs, f0, cnt, width, height, stride, frame_stride, m->s_h2d);
            if (rc) return rc;
            CKM(cudaEventRecord(m->ev[c], m->s_h2d));
        }
    }
    for (int c = 0; c < nchunks; c++) {
        const int f0 = c_f0[c], cnt = c_cnt[c];
        // experiment (ORBX_DEVICE_STREAMS=2): the chunks of a device-resident batch alternate between two extraction streams
        static const bool two_streams = getenv("ORBX_DEVICE_STREAMS") && atoi(getenv("ORBX_DEVICE_STREAMS")) == 2;
        cudaStream_t sc = (!host && two_streams && (c & 1)) ? m->s_h2d : s;
        if (sc != s && c == 1) CKM(cudaStreamWaitEvent(sc, m->ev_start, 0));
        if (prefetched) {
            rc = orbx_ex_run_device(ex, d_pref, orbx_ex_pitch0(ex), orbx_ex_stride0(ex), f0, cnt, lap0, lap1, 1 + f0, s);
        } else if (host) {
            CKM(cudaStreamWaitEvent(s, m->ev[c], 0));
            rc = orbx_ex_run_staged(ex, f0, cnt, lap0, lap1, 1 + f0, s);
        } else {
            rc = orbx_ex_run_device(ex, imgs, stride, (long long)frame_stride, f0, cnt, lap0, lap1, 1 + f0, sc);
        }
        if (rc) return rc;
        if (m->cam_set) {                                // mvKeysUn of this chunk (result slots 1 + f0 ..)
            rc = orbx_undistort_slots_device(ex, 1 + f0, cnt, m->cam_K, m->cam_dist, m->cam_ndist, m->cam_P,
                                             m->d_kps_un + (size_t)(1 + f0) * orbx_ex_out_cap(ex), s);
            if (rc) return rc;
        }
        CKM(cudaEventRecord(m->ev_ext[c], sc));
        if (prefetched && c == nchunks - 1 && (rc = orbx_ex_prefetch_mark_read(ex, d_pref, s))) return rc;
        if (!host && c == nchunks - 1 && (rc = orbx_ex_prefetch_mark_read(ex, imgs, sc))) return rc;      // streaming form: imgs IS a prefetch buffer
        // pairs (slot f0+i, slot f0+i+1) are matched on a second kernel stream, concurrently with the extraction of the
        // next chunk; each chunk owns its slice of the pair scratch
        CKM(cudaStreamWaitEvent(m->s_match, m->ev_ext[c], 0));
        if (sc != s || (two_streams && !host && c > 0)) CKM(cudaStreamWaitEvent(m->s_match, m->ev_ext[c - 1], 0));   // the pair's other slot
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
            if (knn_idx && knn_dist) {
                const size_t wbytes = sizeof(int32_t) * 2 * (cap < m->K ? cap : m->K);
                CKM(cudaMemcpy2DAsync(knn_idx + (size_t)f0 * cap * 2, sizeof(int32_t) * 2 * cap, d_knn_idx + (size_t)f0 * m->K * 2,
                                      sizeof(int32_t) * 2 * m->K, wbytes, cnt, cudaMemcpyDeviceToHost, m->s_d2h));
                CKM(cudaMemcpy2DAsync(knn_dist + (size_t)f0 * cap * 2, sizeof(int32_t) * 2 * cap, d_knn_dist + (size_t)f0 * m->K * 2,
                                      sizeof(int32_t) * 2 * m->K, wbytes, cnt, cudaMemcpyDeviceToHost, m->s_d2h));
            }
        }
    }
    // slot 0 is read by the first chunk's matcher: carry the last frame over only after every matcher finished; the
    // caller's stream `s` thereby also waits for the matcher stream
    CKM(cudaStreamWaitEvent(s, m->ev[ORBX_MAX_CHUNKS + nchunks - 1], 0));
    rc = orbx_extractor_copy_slot(ex, batch, 0, s);
    if (rc) return rc;
    if (m->cam_set) {
        const size_t capx = orbx_ex_out_cap(ex);
        CKM(cudaMemcpyAsync(m->d_kps_un, m->d_kps_un + (size_t)batch * capx, sizeof(orbx_keypoint) * capx, cudaMemcpyDeviceToDevice, s));
    }
    if (!host) return ORBX_OK;                       // asynchronous: the caller synchronises `s`
    CKM(cudaStreamSynchronize(m->s_d2h));
    CKM(cudaStreamSynchronize(m->s_match));
    rc = m_check_err(m, s);      // synchronises the kernel stream
    if (rc) return rc;
    return orbx_ex_fetch_finish(ex, batch, kps, desc, cap, n, mono_index, direct);
}

// A failure in the middle of the pipeline leaves copies / kernels queued on the side streams that read the caller's buffers:
// drain them before the error is returned.
static int extract_match_pipeline(orbx_extractor* ex, orbx_matcher* m, bool host, const uint8_t* imgs, int batch, int width,
                                  int height, int stride, size_t frame_stride, int lap0, int lap1,
                                  const float bounds[4], int window, float nnratio, int check_ori,
                                  orbx_keypoint* kps, uint8_t* desc, int cap, int32_t* n, int32_t* mono_index,
                                  int32_t* matches12, int32_t* nmatches, int32_t* knn_idx, int32_t* knn_dist,
                                  int32_t* d_matches12, int32_t* d_nmatches, int32_t* d_knn_idx, int32_t* d_knn_dist,
                                  cudaStream_t s)
{
    const int rc = extract_match_pipeline_impl(ex, m, host, imgs, batch, width, height, stride, frame_stride, lap0, lap1, bounds, window,
                                               nnratio, check_ori, kps, desc, cap, n, mono_index, matches12, nmatches, knn_idx, knn_dist,
                                               d_matches12, d_nmatches, d_knn_idx, d_knn_dist, s);
    if (rc != ORBX_OK && m->s_h2d) {
        cudaStreamSynchronize(m->s_h2d); cudaStreamSynchronize(m->s_match); cudaStreamSynchronize(m->s_d2h); cudaStreamSynchronize(s);
        cudaGetLastError();
    }
    return rc;
}

// argument checks shared by the two entry points: the batch must fit the matcher's pair scratch AND the extractor's result slots
// 1..batch / per-frame scratch, both handles must live on the same device
static int pipeline_args_ok(orbx_extractor* ex, orbx_matcher* m, const void* imgs, int batch, int width, int stride, const float* bounds)
{
    if (!ex || !m || !imgs || !bounds || batch < 1) { orbx_set_error("%s%s", "orbx_extract_match_batch: null argument or batch < 1", ""); return ORBX_E_INVALID; }
    if (batch > m->P) { orbx_set_error("%s%s", "orbx_extract_match_batch: batch larger than the matcher's max_batch", ""); return ORBX_E_INVALID; }
    if (batch > orbx_ex_max_batch(ex)) { orbx_set_error("%s%s", "orbx_extract_match_batch: batch larger than the extractor's max_batch", ""); return ORBX_E_INVALID; }
    if (orbx_ex_device(ex) != m->p.device) { orbx_set_error("%s%s", "orbx_extract_match_batch: extractor and matcher are on different devices", ""); return ORBX_E_INVALID; }
    if (width > 0 && stride < width) { orbx_set_error("%s%s", "orbx_extract_match_batch: stride smaller than width", ""); return ORBX_E_INVALID; }
    return ORBX_OK;
}

// ---- streaming form: submit queues one batch and returns at once, wait blocks until that batch's results are on the host ----
// Per batch: the input copy (the prefetch path: own stream, one of two staging buffers), the kernels of the device-resident step
// (one chunk), a device-to-device copy of the results into one of two staging sets (~45 MB, microseconds), and the result copy to
// the caller's pinned buffers on the D2H stream.  With two batches in flight the three run at the same time for batches k+1, k
// and k-1.
static int stream_init(orbx_matcher* m, orbx_extractor* ex)
{
    const int cap = orbx_ex_out_cap(ex);
    if (m->st_init && m->st_cap == cap) return ORBX_OK;
    if (m->st_init) { orbx_set_error("%s%s", "orbx_stream_submit: the extractor's result capacity changed", ""); return ORBX_E_INVALID; }
    const size_t rows = (size_t)m->P;
    for (int i = 0; i < 2; i++) {
        orbx_matcher::StreamSet& S = m->st[i];
        CKM(cudaMalloc((void**)&S.kps, sizeof(orbx_keypoint) * rows * cap)); m->allocs.push_back(S.kps);
        CKM(cudaMalloc((void**)&S.desc, (size_t)32 * rows * cap)); m->allocs.push_back(S.desc);
        CKM(cudaMalloc((void**)&S.n, sizeof(int32_t) * rows * 3)); m->allocs.push_back(S.n);
        S.mono = S.n + rows; S.nm = S.mono + rows;
        CKM(cudaMalloc((void**)&S.m12, sizeof(int32_t) * rows * m->K * 5)); m->allocs.push_back(S.m12);
        S.knn_idx = S.m12 + rows * m->K; S.knn_dist = S.knn_idx + rows * m->K * 2;
        CKM(cudaEventCreateWithFlags(&S.ev_kernels, cudaEventDisableTiming));
        CKM(cudaEventCreateWithFlags(&S.ev_host, cudaEventDisableTiming));
        CKM(cudaMallocHost((void**)&S.h_err, 2 * sizeof(unsigned)));
        S.busy = false; S.batch = 0;
    }
    m->st_rows = rows; m->st_cap = cap; m->st_ticket = 0; m->st_init = true;
    return ORBX_OK;
}

extern "C" int orbx_stream_submit(orbx_extractor* ex, orbx_matcher* m, const uint8_t* imgs, int batch, int width, int height, int stride,
                                  size_t frame_stride, int lap0, int lap1, const float bounds[4], int window, float nnratio, int check_ori,
                                  orbx_keypoint* kps, uint8_t* desc, int cap, int32_t* n, int32_t* mono_index,
                                  int32_t* matches12, int32_t* nmatches, int32_t* knn_idx, int32_t* knn_dist, long long* ticket)
{
    int rc = pipeline_args_ok(ex, m, imgs, batch, width, stride, bounds);
    if (rc) return rc;
    if (!ticket || !kps || !desc || !n || !mono_index || !matches12 || !nmatches || (knn_idx == nullptr) != (knn_dist == nullptr)) return ORBX_E_INVALID;
    if (width <= 0 || height <= 0) return ORBX_E_EMPTY;
    CKM(cudaSetDevice(m->p.device));
    if ((rc = orbx_ex_configure(ex, width, height))) return rc;
    if ((rc = orbx_m_ensure_pipeline(m))) return rc;
    if ((rc = stream_init(m, ex))) return rc;
    if (cap != orbx_ex_out_cap(ex) || !orbx_host_pinned(kps) || !orbx_host_pinned(desc) || !orbx_host_pinned(n) || !orbx_host_pinned(mono_index) ||
        !orbx_host_pinned(matches12) || !orbx_host_pinned(nmatches) || (knn_idx && (!orbx_host_pinned(knn_idx) || !orbx_host_pinned(knn_dist)))) {
        orbx_set_error("%s%s", "orbx_stream_submit: result buffers must be pinned host memory with cap = orbx_extractor_max_keypoints", "");
        return ORBX_E_INVALID;
    }
    orbx_matcher::StreamSet& S = m->st[m->st_ticket & 1];
    if (S.busy) { orbx_set_error("%s%s", "orbx_stream_submit: two batches are in flight, wait for the older one first", ""); return ORBX_E_CAPACITY; }
    cudaStream_t s = orbx_ex_stream(ex);
    // input: already prefetched, or copied now on the copy stream (after the kernels that last read that staging buffer)
    const uint8_t* d_in = nullptr; cudaEvent_t ready = nullptr;
    if (!orbx_ex_take_prefetched(ex, imgs, batch, width, height, &d_in, &ready)) {
        rc = orbx_ex_prefetch(ex, imgs, batch, width, height, stride, frame_stride, m->s_h2d);
        if (rc) return rc;
        if (!orbx_ex_take_prefetched(ex, imgs, batch, width, height, &d_in, &ready)) return ORBX_E_INVALID;
    }
    CKM(cudaStreamWaitEvent(s, ready, 0));
    // kernels: the device-resident step (one chunk; matcher kernels on the second stream, joined into s at the end)
    if (knn_idx) {
        const size_t need = (size_t)m->P * m->K * 2;
        if (need > m->pipe_knn_elems) {
            if (m->d_pipe_knn) { CKM(cudaDeviceSynchronize()); cudaFree(m->d_pipe_knn); m->d_pipe_knn = nullptr; m->pipe_knn_elems = 0; }
            CKM(cudaMalloc((void**)&m->d_pipe_knn, sizeof(int32_t) * need * 2));
            m->pipe_knn_elems = need;
        }
    }
    int32_t* d_ki = knn_idx ? m->d_pipe_knn : nullptr; int32_t* d_kd = knn_idx ? m->d_pipe_knn + (size_t)m->P * m->K * 2 : nullptr;
    rc = extract_match_pipeline(ex, m, false, d_in, batch, width, height, orbx_ex_pitch0(ex), (size_t)orbx_ex_stride0(ex), lap0, lap1, bounds, window,
                                nnratio, check_ori, nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                m->d_out, m->d_nm, d_ki, d_kd, s);
    if (rc) return rc;
    // results -> staging set (device to device, on the kernel stream), flags included
    orbx_keypoint* dk; uint8_t* dd; int32_t* dn; int32_t* dmono; int ocap, slots;
    if ((rc = orbx_extractor_results_device(ex, &dk, &dd, &dn, &dmono, &ocap, &slots))) return rc;
    const size_t B = (size_t)batch;
    CKM(cudaMemcpyAsync(S.kps, dk + (size_t)ocap, sizeof(orbx_keypoint) * ocap * B, cudaMemcpyDeviceToDevice, s));          // slots 1 .. batch
    CKM(cudaMemcpyAsync(S.desc, dd + (size_t)ocap * 32, (size_t)32 * ocap * B, cudaMemcpyDeviceToDevice, s));
    CKM(cudaMemcpyAsync(S.n, dn + 1, sizeof(int32_t) * B, cudaMemcpyDeviceToDevice, s));
    CKM(cudaMemcpyAsync(S.mono, dmono + 1, sizeof(int32_t) * B, cudaMemcpyDeviceToDevice, s));
    CKM(cudaMemcpyAsync(S.nm, m->d_nm, sizeof(int32_t) * B, cudaMemcpyDeviceToDevice, s));
    CKM(cudaMemcpyAsync(S.m12, m->d_out, sizeof(int32_t) * B * m->K, cudaMemcpyDeviceToDevice, s));
    if (knn_idx) {
        CKM(cudaMemcpyAsync(S.knn_idx, d_ki, sizeof(int32_t) * B * m->K * 2, cudaMemcpyDeviceToDevice, s));
        CKM(cudaMemcpyAsync(S.knn_dist, d_kd, sizeof(int32_t) * B * m->K * 2, cudaMemcpyDeviceToDevice, s));
    }
    CKM(cudaEventRecord(S.ev_kernels, s));
    // results -> the caller's pinned buffers on the D2H stream; device rows have stride K, host rows stride cap
    cudaStream_t sd = m->s_d2h;
    CKM(cudaStreamWaitEvent(sd, S.ev_kernels, 0));
    CKM(cudaMemcpyAsync(S.h_err, orbx_ex_err_device(ex), sizeof(unsigned), cudaMemcpyDeviceToHost, sd));
    CKM(cudaMemcpyAsync(S.h_err + 1, m->W.err, sizeof(unsigned), cudaMemcpyDeviceToHost, sd));
    CKM(cudaMemcpyAsync(n, S.n, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, sd));
    CKM(cudaMemcpyAsync(mono_index, S.mono, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, sd));
    CKM(cudaMemcpyAsync(nmatches, S.nm, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, sd));
    CKM(cudaMemcpyAsync(kps, S.kps, sizeof(orbx_keypoint) * ocap * B, cudaMemcpyDeviceToHost, sd));
    CKM(cudaMemcpyAsync(desc, S.desc, (size_t)32 * ocap * B, cudaMemcpyDeviceToHost, sd));
    const size_t wcols = (size_t)(cap < m->K ? cap : m->K);
    CKM(cudaMemcpy2DAsync(matches12, sizeof(int32_t) * cap, S.m12, sizeof(int32_t) * m->K, sizeof(int32_t) * wcols, B, cudaMemcpyDeviceToHost, sd));
    if (knn_idx) {
        CKM(cudaMemcpy2DAsync(knn_idx, sizeof(int32_t) * 2 * cap, S.knn_idx, sizeof(int32_t) * 2 * m->K, sizeof(int32_t) * 2 * wcols, B, cudaMemcpyDeviceToHost, sd));
        CKM(cudaMemcpy2DAsync(knn_dist, sizeof(int32_t) * 2 * cap, S.knn_dist, sizeof(int32_t) * 2 * m->K, sizeof(int32_t) * 2 * wcols, B, cudaMemcpyDeviceToHost, sd));
    }
    CKM(cudaEventRecord(S.ev_host, sd));
    // (the staging set and the input buffer are reused by the batch after next, which cannot be submitted before this one was waited for)
    S.busy = true; S.batch = batch;
    *ticket = m->st_ticket++;
    return ORBX_OK;
}

extern "C" int orbx_stream_wait(orbx_extractor* ex, orbx_matcher* m, long long ticket)
{
    if (!ex || !m || !m->st_init || ticket < 0 || ticket >= m->st_ticket) return ORBX_E_INVALID;
    orbx_matcher::StreamSet& S = m->st[ticket & 1];
    if (!S.busy || ticket + 2 < m->st_ticket) return ORBX_OK;             // already waited for
    CKM(cudaSetDevice(m->p.device));
    CKM(cudaEventSynchronize(S.ev_host));
    S.busy = false;
    if (S.h_err[0] || S.h_err[1]) {
        char buf[48]; snprintf(buf, sizeof(buf), "extractor 0x%x matcher 0x%x", S.h_err[0], S.h_err[1]);
        orbx_set_error("orbx_stream_wait: device capacity error flags %s%s", buf, " (raise max_candidates / capacities)");
        cudaMemsetAsync(orbx_ex_err_device(ex), 0, sizeof(unsigned), orbx_ex_stream(ex));
        cudaMemsetAsync(m->W.err, 0, sizeof(unsigned), orbx_ex_stream(ex));
        return ORBX_E_CAPACITY;
    }
    return ORBX_OK;
}

// Starts the host-to-device copy of the frames of the NEXT orbx_extract_match_batch call (same arguments) and returns at once.
extern "C" int orbx_extract_match_batch_prefetch(orbx_extractor* ex, orbx_matcher* m, const uint8_t* imgs, int batch, int width,
                                                 int height, int stride, size_t frame_stride)
{
    if (!ex || !m || !imgs || batch < 1 || width <= 0 || height <= 0 || stride < width) return ORBX_E_INVALID;
    if (orbx_ex_device(ex) != m->p.device) { orbx_set_error("%s%s", "orbx_extract_match_batch_prefetch: extractor and matcher live on different devices", ""); return ORBX_E_INVALID; }
    CKM(cudaSetDevice(m->p.device));
    int rc = orbx_m_ensure_pipeline(m);
    if (rc) return rc;
    return orbx_ex_prefetch(ex, imgs, batch, width, height, stride, frame_stride, m->s_h2d);
}

extern "C" int orbx_extract_match_batch(orbx_extractor* ex, orbx_matcher* m, const uint8_t* imgs, int batch, int width,
                                        int height, int stride, size_t frame_stride, int lap0, int lap1,
                                        const float bounds[4], int window, float nnratio, int check_ori,
                                        orbx_keypoint* kps, uint8_t* desc, int cap, int32_t* n, int32_t* mono_index,
                                        int32_t* matches12, int32_t* nmatches, int32_t* knn_idx, int32_t* knn_dist)
{
    int rc = pipeline_args_ok(ex, m, imgs, batch, width, stride, bounds);
    if (rc) return rc;
    if ((knn_idx == nullptr) != (knn_dist == nullptr)) return ORBX_E_INVALID;
    if (width <= 0 || height <= 0) return ORBX_E_EMPTY;
    return extract_match_pipeline(ex, m, true, imgs, batch, width, height, stride, frame_stride, lap0, lap1, bounds, window, nnratio,
                                  check_ori, kps, desc, cap, n, mono_index, matches12, nmatches, knn_idx, knn_dist,
                                  nullptr, nullptr, nullptr, nullptr, orbx_ex_stream(ex));
}

extern "C" int orbx_extract_match_batch_device(orbx_extractor* ex, orbx_matcher* m, const uint8_t* d_imgs, int batch, int width,
                                               int height, int stride, size_t frame_stride, int lap0, int lap1,
                                               const float bounds[4], int window, float nnratio, int check_ori,
                                               int32_t* d_matches12, int32_t* d_nmatches, int32_t* d_knn_idx, int32_t* d_knn_dist,
                                               void* stream)
{
    int rc = pipeline_args_ok(ex, m, d_imgs, batch, width, stride, bounds);
    if (rc) return rc;
    if (!d_matches12 || !d_nmatches) return ORBX_E_INVALID;
    if (width <= 0 || height <= 0) return ORBX_E_EMPTY;
    return extract_match_pipeline(ex, m, false, d_imgs, batch, width, height, stride, frame_stride, lap0, lap1, bounds, window, nnratio,
                                  check_ori, nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, d_matches12, d_nmatches, d_knn_idx,
                                  d_knn_dist, stream ? (cudaStream_t)stream : orbx_ex_stream(ex));
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
