// orbx_octree.cu - DistributeOctTree on the GPU, one CTA per (frame, level).
//
// Replaces ORBextractor::DistributeOctTree / ExtractorNode::DivideNode (R/src/ORBextractor.cc:479-761).
//
// The reference walks a std::list of nodes and copies key vectors on every split.  Two observations make
// it parallel without changing one result bit:
//  (1) The boxes of the tree depend only on the level geometry (halfX = ceil(w/2) recursion), so every
//      candidate can compute its own root and its whole root-to-leaf path (2 bits per depth) on its own.
//      Sorting the candidates by (root, path) makes EVERY possible node a contiguous range of the sorted
//      array; a split is three binary searches, no point is ever moved again.
//  (2) Within one sweep of the reference's loops all splits are independent; the list order after a sweep
//      is "children in reverse creation order, then the untouched nodes in their old order" (push_front
//      at :621-656 / :691-724, erase at :660 / :726).  The largest-first phase (:671-736) processes nodes
//      by (size, creation) descending and breaks at the first node that lifts the list to >= N nodes,
//      which is a prefix-sum + first-index search.
// The sort at :682 orders equal-sized nodes by heap address in the reference (non-deterministic);
// the canonical rule used by the oracle and here is: equal size -> later-created node first.
// Winner per node (:742-758): greatest response, first in candidate order on ties.
#include "orbx_internal.h"

namespace {

constexpr int NT = 512;
constexpr int DMAX = 13;                 // path depth: 26 bits; root index: 6 bits
constexpr int MAX_IPT = 8;               // node items per thread -> up to 4096 nodes
static_assert(MAX_IPT * NT == ORBX_OCTREE_MAX_NODES, "orbx_extractor_create validates level quotas against this bound");
typedef unsigned long long u64;

__device__ __forceinline__ u64 block_scan_excl(u64 v, u64* total, u64* s_warp)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { u64 t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    __syncthreads();                      // protect s_warp reuse
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    // every warp scans the NT/32 warp totals with shuffles (lane = warp index)
    u64 ws = lane < NT / 32 ? s_warp[lane] : 0ull, wi = ws;
#pragma unroll
    for (int o = 1; o < NT / 32; o <<= 1) { u64 t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
    *total = __shfl_sync(0xffffffffu, wi, NT / 32 - 1);
    const u64 base = __shfl_sync(0xffffffffu, wi - ws, warp);
    return base + inc - v;
}

// npad <= NT: one key per thread in a register.  Exchanges at distance < 32 are shuffles; only the distances >= 32 go through
// shared memory, alternating between two halves of a scratch so that one barrier per such step is enough (npad = 512: 10 barriers
// instead of 45).  The small pyramid levels and the largest-first rounds spend their time in these barriers, not in the compares.
template <typename T, bool DESC>
__device__ void bitonic_sort_small(T* a, T* scratch, int npad)
{
    const int t = threadIdx.x;
    const bool act = t < npad;
    T x = act ? a[t] : T(0);
    int flip = 0;
    for (int k = 2; k <= npad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            T y;
            if (j >= 32) {
                T* sbuf = scratch + flip * NT; flip ^= 1;
                if (act) sbuf[t] = x;
                __syncthreads();
                y = act ? sbuf[t ^ j] : T(0);
            } else {
                if (sizeof(T) == 8) {
                    const unsigned lo = __shfl_xor_sync(0xffffffffu, (unsigned)((unsigned long long)x & 0xffffffffull), j);
                    const unsigned hi = __shfl_xor_sync(0xffffffffu, (unsigned)((unsigned long long)x >> 32), j);
                    y = (T)(((unsigned long long)hi << 32) | lo);
                } else {
                    y = (T)__shfl_xor_sync(0xffffffffu, (unsigned)x, j);
                }
            }
            const bool up = ((t & k) == 0) != DESC;          // this block of k sorts ascending
            const bool lower = (t & j) == 0;                 // this thread holds the lower position of the pair
            const bool take_min = lower == up;
            if ((y < x) == take_min && y != x) x = y;        // min or max of the pair
        }
    }
    __syncthreads();                                          // (the scratch may alias `a`)
    if (act) a[t] = x;
    __syncthreads();
}

template <typename T, bool DESC>
__device__ void bitonic_sort(T* a, int npad)
{
    for (int k = 2; k <= npad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (npad >> 1); t += NT) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const bool up = ((i & k) == 0) != DESC;
                T x = a[i], y = a[l];
                if ((x > y) == up) { a[i] = y; a[l] = x; }
            }
            __syncthreads();
        }
    }
}

// Stable LSD radix sort of the indices 0 .. npad-1 by their 32-bit path key, 4 bits per pass, for npad a multiple of 512 (NT = 16
// warps, warp w owns the npad / 16 consecutive positions w * C ..).  A pass: every warp walks its positions in order, 32 per step;
// match.any groups the lanes of a digit, the lowest lane of a group advances the warp's private counter of that digit, a lane's
// rank is counter + its place in the group (so equal digits keep their order); the 16 x 16 counters are scanned digit-major and
// the indices scattered.  8 passes x 4 barriers against the 66 - 78 barriers (and ~2.5 x the instructions) of the bitonic network
// on 64-bit keys.  keys[] is never moved: a pass reads the key through the index.  Result in ia.
__device__ __noinline__ void radix_sort_indices(const uint32_t* keys, unsigned short* ia, unsigned short* ib, unsigned short* hist, int* s_part, int npad)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = npad / (NT / 32), IT = C >> 5;                          // IT <= 8 for npad <= 4096
    unsigned short* in = ia; unsigned short* out = ib;
    for (int pass = 0; pass < 8; pass++) {
        const int shift = 4 * pass;
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        int rk[8]; unsigned short id[8];
#pragma unroll
        for (int it = 0; it < 8; it++) {
            if (it < IT) {
                const int i = warp * C + it * 32 + lane;
                const unsigned short x = pass == 0 ? (unsigned short)i : in[i];
                const unsigned d = (keys[x] >> shift) & 15u;
                const unsigned peers = __match_any_sync(0xffffffffu, d);
                const int leader = __ffs(peers) - 1;
                int old = 0;
                if (lane == leader) { old = hist[warp * 16 + d]; hist[warp * 16 + d] = (unsigned short)(old + __popc(peers)); }
                old = __shfl_sync(0xffffffffu, old, leader);
                rk[it] = (old + __popc(peers & ((1u << lane) - 1u))) | (int)(d << 16);
                id[it] = x;
                __syncwarp();
            }
        }
        __syncthreads();
        // exclusive scan of the counters in (digit, warp) order by the first 256 threads
        int cnt = 0, inc = 0;
        if (tid < 256) {
            cnt = hist[(tid & 15) * 16 + (tid >> 4)];
            inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            if (lane == 31) s_part[warp] = inc;
        }
        __syncthreads();
        if (tid < 256) {
            int base = inc - cnt;
            for (int w2 = 0; w2 < warp; w2++) base += s_part[w2];
            hist[(tid & 15) * 16 + (tid >> 4)] = (unsigned short)base;
        }
        __syncthreads();
#pragma unroll
        for (int it = 0; it < 8; it++)
            if (it < IT) out[hist[warp * 16 + (rk[it] >> 16)] + (rk[it] & 0xFFFF)] = id[it];
        __syncthreads();
        unsigned short* t = in; in = out; out = t;
    }
}

__device__ __forceinline__ int lower_bound_key(const u64* buf, int lo, int hi, unsigned key)
{
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((unsigned)(buf[mid] >> 32) < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// node record: lo | hi << 20 | depth << 40
__device__ __forceinline__ u64 mk_node(int lo, int hi, int depth) { return (u64)lo | ((u64)hi << 20) | ((u64)depth << 40); }
__device__ __forceinline__ int nd_lo(u64 n) { return (int)(n & 0xFFFFF); }
__device__ __forceinline__ int nd_hi(u64 n) { return (int)((n >> 20) & 0xFFFFF); }
__device__ __forceinline__ int nd_depth(u64 n) { return (int)(n >> 40); }

// split of a multi-point node: child boundaries b[0..4] (b0=lo, b4=hi); returns false if the key depth is exhausted
__device__ __forceinline__ bool split_node(const u64* buf, u64 node, int (&bd)[5])
{
    const int lo = nd_lo(node), hi = nd_hi(node), d = nd_depth(node) + 1;
    bd[0] = lo; bd[4] = hi;
    if (d > DMAX) { bd[1] = bd[2] = bd[3] = hi; return false; }
    const int shift = 2 * (DMAX - d);
    const unsigned P = (unsigned)(buf[lo] >> 32) >> (shift + 2);
    bd[1] = lower_bound_key(buf, lo, hi, ((P << 2) | 1u) << shift);
    bd[2] = lower_bound_key(buf, bd[1], hi, ((P << 2) | 2u) << shift);
    bd[3] = lower_bound_key(buf, bd[2], hi, ((P << 2) | 3u) << shift);
    return true;
}

struct OctShared {
    u64 warp_sums[NT / 32];
    int n_pts, n_nodes, n_vec, jstar, flag_overflow;
    int tot_c, tot_e;
    int radix_part[8];
    int row_prefix[ORBX_MAX_UNITS + 1];
};

__global__ void __launch_bounds__(NT, 4) k_octree(OrbxGeom g, OrbxBuffers b, int smem_pts, int ncap, int level_base, int allow_radix)
{
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ OctShared sh;
    const int tid = threadIdx.x;
    const int l = level_base + blockIdx.x, f = blockIdx.y;
    const OrbxLevel L = g.lv[l];
    int* out_n = b.lvl_n + (long long)f * g.nlevels + l;
    uint32_t* out_kp = b.lvl_kp + (long long)f * g.kp_total_cap + L.kp_base;

    // ---- smem carve-up ----
    u64* s_sort = reinterpret_cast<u64*>(smem);                          // [smem_pts]
    uint32_t* s_pts = reinterpret_cast<uint32_t*>(s_sort + smem_pts);    // [smem_pts]
    u64* La = reinterpret_cast<u64*>(s_pts + smem_pts);                  // [ncap]
    u64* Lb = La + ncap;                                                 // [ncap]
    unsigned* Va = reinterpret_cast<unsigned*>(Lb + ncap);               // [ncap] multi-point nodes (list positions), creation order
    unsigned* Vb = Va + ncap;                                            // [ncap]
    unsigned* Vs = Vb + ncap;                                            // [pow2(ncap)] sort keys
    uint8_t* proc = reinterpret_cast<uint8_t*>(Vs + ncap * 2);           // [ncap]

    orbx_pdl_prologue();
    // ---- gather the level's candidates in reference order (cell rows top to bottom) ----
    if (L.nRows <= 0 || L.nCols <= 0) { if (tid == 0) *out_n = 0; return; }
    const int* rc = b.row_count + (long long)f * g.total_rows + L.row_base;
    const int nUnits = L.nRows * L.nSeg;               // FAST segments of the level, reference order
    if (tid < 32) {
        // exclusive prefix of the segment counts, 32 segments per step
        int run = 0;
        for (int r0 = 0; r0 < nUnits; r0 += 32) {
            const int r = r0 + tid;
            const int c = r < nUnits ? rc[r] : 0;
            int inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (tid >= o) inc += t; }
            if (r < nUnits) sh.row_prefix[r] = run + inc - c;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (tid == 0) {
            sh.row_prefix[nUnits] = run;
            if (run > L.cand_cap) { atomicOr(b.err, ORBX_DEVERR_CAND_OVERFLOW); run = L.cand_cap; }
            sh.n_pts = run;
        }
    }
    __syncthreads();
    const int n = sh.n_pts;
    if (n == 0) { if (tid == 0) *out_n = 0; return; }
    int npad = 1; while (npad < n) npad <<= 1;
    u64* buf; uint32_t* pts;
    if (npad <= smem_pts) { buf = s_sort; pts = s_pts; }
    else {
        u64* scratch = b.sort_scratch + (long long)f * b.sort_scratch_stride + b.sort_off[l];
        buf = scratch; pts = reinterpret_cast<uint32_t*>(scratch + npad);
    }
    const uint32_t* cand = b.row_cand + (long long)f * b.row_cand_stride;
    for (int r = tid >> 5; r < nUnits; r += NT / 32) {
        const int base = sh.row_prefix[r], cnt = min(sh.row_prefix[r + 1], n) - base;
        const uint32_t* src = cand + b.row_off[L.row_base + r];
        for (int k = tid & 31; k < cnt; k += 32) pts[base + k] = src[k];
    }
    __syncthreads();

    // ---- per-candidate root + path key (DivideNode geometry, :479-507; root assignment :553-567) ----
    const int Hbox = L.maxBY - ORBX_BORDER;
    // shared-memory sort buffers of 1024 .. 4096 entries take the radix path: keys[npad] u32 | ia[npad] u16 | ib[npad] u16 = the
    // 8 * npad bytes of buf; the 16 x 16 counters live in the (still unused) node list
    const bool radix = allow_radix && buf == s_sort && npad >= 1024 && npad <= 4096;
    uint32_t* rkeys = reinterpret_cast<uint32_t*>(buf);
    for (int i = tid; i < npad; i += NT) {
        u64 e = ~0ull;
        if (i < n) {
            const uint32_t p = pts[i];
            const int px = p & 0xFFF, py = (p >> 12) & 0xFFF;
            int r = (int)((float)px / L.hX);
            if (r > L.nIni - 1) r = L.nIni - 1;
            int x0 = (int)(L.hX * (float)r), x1 = (int)(L.hX * (float)(r + 1)), y0 = 0, y1 = Hbox;
            unsigned key = (unsigned)r;
#pragma unroll 1
            for (int d = 0; d < DMAX; d++) {
                const int mx = x0 + ((x1 - x0 + 1) >> 1), my = y0 + ((y1 - y0 + 1) >> 1);
                const unsigned cx = px < mx ? 0u : 1u, cy = py < my ? 0u : 1u;
                key = (key << 2) | (cy << 1) | cx;
                if (cx) x0 = mx; else x1 = mx;
                if (cy) y0 = my; else y1 = my;
            }
            e = ((u64)key << 32) | ((u64)(p >> 24) << 24) | (u64)(0xFFFFFF - i);
        }
        if (radix) rkeys[i] = (uint32_t)(e >> 32); else buf[i] = e;
    }
    __syncthreads();
    if (radix) {
        unsigned short* ia = reinterpret_cast<unsigned short*>(rkeys + npad);
        radix_sort_indices(rkeys, ia, ia + npad, reinterpret_cast<unsigned short*>(La), sh.radix_part, npad);
        // the sorted records (key | score | 0xFFFFFF - index) replace keys and indices: read everything, then write
        u64 ev[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int i = tid + k * NT;
            if (i < npad) {
                const int x = ia[i];
                ev[k] = x < n ? ((u64)rkeys[x] << 32) | ((u64)(pts[x] >> 24) << 24) | (u64)(0xFFFFFF - x) : ~0ull;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; k++) { const int i = tid + k * NT; if (i < npad) buf[i] = ev[k]; }
        __syncthreads();
    } else {
        // (buf holds smem_pts >= 4 * NT entries in shared memory, or the global scratch: entries npad .. are free)
        if (npad <= NT && buf == s_sort) bitonic_sort_small<u64, false>(buf, buf + NT, npad);
        else bitonic_sort<u64, false>(buf, npad);
    }

    // ---- initial list: non-empty roots in order (:550-583) ----
    if (tid == 0) {
        int cnt = 0, lo = 0;
        for (int r = 0; r < L.nIni; r++) {
            const int hi = (r == L.nIni - 1) ? n : lower_bound_key(buf, lo, n, (unsigned)(r + 1) << (2 * DMAX));
            if (hi > lo) La[cnt++] = mk_node(lo, hi, 0);
            lo = hi;
        }
        sh.n_nodes = cnt;
        sh.flag_overflow = 0;
    }
    __syncthreads();

    const int N = L.quota;
    const int ipt = (ncap + NT - 1) / NT;
    u64* Lcur = La; u64* Lnew = Lb;
    unsigned* Vcur = Va; unsigned* Vnew = Vb;
    int nn = sh.n_nodes;
    bool finish = false;
    bool sorted_phase = false;
    int nvec = 0;

    while (!finish) {
        if (!sorted_phase) {
            // ================= full sweep over the list (:604-663) =================
            int bd[MAX_IPT][5];
            int kind[MAX_IPT];               // 0 none, 1 unsplit, 2 split
            u64 mine = 0;                    // packed counts: children | unsplit << 20 | multi << 40
            for (int it = 0; it < ipt; it++) {
                const int i = tid * ipt + it;
                kind[it] = 0;
                if (i < nn) {
                    const u64 nd = Lcur[i];
                    if (nd_hi(nd) - nd_lo(nd) > 1 && split_node(buf, nd, bd[it])) {
                        kind[it] = 2;
                        int c = 0, e = 0;
#pragma unroll
                        for (int k = 0; k < 4; k++) { const int s = bd[it][k + 1] - bd[it][k]; c += s > 0; e += s > 1; }
                        mine += (u64)c | ((u64)e << 40);
                    } else {
                        if (nd_hi(nd) - nd_lo(nd) > 1) atomicOr(b.err, ORBX_DEVERR_OCTREE_DEPTH);
                        kind[it] = 1;
                        mine += 1ull << 20;
                    }
                }
            }
            u64 tot;
            u64 pre = block_scan_excl(mine, &tot, sh.warp_sums);
            const int C = (int)(tot & 0xFFFFF), U = (int)((tot >> 20) & 0xFFFFF), E = (int)(tot >> 40);
            int pc = (int)(pre & 0xFFFFF), pu = (int)((pre >> 20) & 0xFFFFF), pe = (int)(pre >> 40);
            const int nnew = C + U;
            if (nnew > ncap) {               // cannot happen for a quota-consistent ncap; never write out of bounds
                if (tid == 0) atomicOr(b.err, ORBX_DEVERR_NODE_OVERFLOW);
                break;
            }
            for (int it = 0; it < ipt; it++) {
                const int i = tid * ipt + it;
                if (kind[it] == 2) {
                    const int d = nd_depth(Lcur[i]) + 1;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int s = bd[it][k + 1] - bd[it][k];
                        if (s > 0) {
                            const int pos = C - 1 - pc;
                            Lnew[pos] = mk_node(bd[it][k], bd[it][k + 1], d);
                            if (s > 1) Vnew[pe++] = (unsigned)pos;
                            pc++;
                        }
                    }
                } else if (kind[it] == 1) {
                    Lnew[C + pu] = Lcur[i];
                    pu++;
                }
            }
            __syncthreads();
            { u64* t = Lcur; Lcur = Lnew; Lnew = t; }
            { unsigned* t = Vcur; Vcur = Vnew; Vnew = t; }
            const int prev = nn;
            nn = nnew; nvec = E;
            if (nn >= N || nn == prev) finish = true;                 // :667
            else if (nn + 3 * nvec > N) sorted_phase = true;          // :671
        } else {
            // ================= largest-first round (:674-735) =================
            const int prev = nn;
            int mpad = 1; while (mpad < nvec) mpad <<= 1;
            for (int j = tid; j < mpad; j += NT) {
                unsigned key = 0;                                      // pads sort to the end (descending)
                if (j < nvec) {
                    const u64 nd = Lcur[Vcur[j]];
                    key = ((unsigned)(nd_hi(nd) - nd_lo(nd)) << 12) | (unsigned)j;   // (size, creation seq)
                }
                Vs[j] = key;
            }
            for (int i = tid; i < nn; i += NT) proc[i] = 0;
            if (tid == 0) sh.jstar = nvec - 1;
            __syncthreads();
            if (mpad > 1) {
                // scratch: the segment-prefix table of the gather phase (ORBX_MAX_UNITS + 1 >= 2 * NT words), dead by now
                static_assert(ORBX_MAX_UNITS + 1 >= 2 * NT, "row_prefix doubles as the small sort's scratch");
                if (mpad <= NT) bitonic_sort_small<unsigned, true>(Vs, reinterpret_cast<unsigned*>(sh.row_prefix), mpad);
                else bitonic_sort<unsigned, true>(Vs, mpad);
            }
            int bd[MAX_IPT][5];
            int lpos[MAX_IPT];
            u64 mine = 0;                    // children | multi << 40   (growth = children - 1)
            for (int it = 0; it < ipt; it++) {
                const int j = tid * ipt + it;
                lpos[it] = -1;
                if (j < nvec) {
                    const int vp = Vs[j] & 0xFFF;
                    lpos[it] = (int)Vcur[vp];
                    const u64 nd = Lcur[lpos[it]];
                    int c = 0, e = 0;
                    if (split_node(buf, nd, bd[it])) {
#pragma unroll
                        for (int k = 0; k < 4; k++) { const int s = bd[it][k + 1] - bd[it][k]; c += s > 0; e += s > 1; }
                    } else {
                        atomicOr(b.err, ORBX_DEVERR_OCTREE_DEPTH);
                        bd[it][1] = bd[it][2] = bd[it][3] = bd[it][4]; c = 1; e = 1;   // degenerate: node reproduces itself
                    }
                    mine += (u64)c | ((u64)e << 40);
                }
            }
            u64 tot;
            u64 pre = block_scan_excl(mine, &tot, sh.warp_sums);
            // break index: first j (processing order) with  nn + sum_{<=j}(c-1) >= N   (:728-729)
            {
                int pc = (int)(pre & 0xFFFFF);
                for (int it = 0; it < ipt; it++) {
                    const int j = tid * ipt + it;
                    if (j < nvec) {
                        int c = 0;
#pragma unroll
                        for (int k = 0; k < 4; k++) c += (bd[it][k + 1] - bd[it][k]) > 0;
                        pc += c;
                        if (nn + pc - (j + 1) >= N) { atomicMin(&sh.jstar, j); break; }
                    }
                }
            }
            __syncthreads();
            const int jstar = sh.jstar;
            // totals over the processed prefix
            {
                int pc = (int)(pre & 0xFFFFF), pe = (int)(pre >> 40);
                for (int it = 0; it < ipt; it++) {
                    const int j = tid * ipt + it;
                    if (j < nvec && j <= jstar) {
                        int c = 0, e = 0;
#pragma unroll
                        for (int k = 0; k < 4; k++) { const int s = bd[it][k + 1] - bd[it][k]; c += s > 0; e += s > 1; }
                        pc += c; pe += e;
                        proc[lpos[it]] = 1;
                        if (j == jstar) { sh.tot_c = pc; sh.tot_e = pe; }
                    }
                }
            }
            if (nvec == 0 && tid == 0) { sh.tot_c = 0; sh.tot_e = 0; }
            __syncthreads();
            const int Cp = sh.tot_c, Ep = sh.tot_e;
            const int nproc = jstar + 1;
            const int nnew = nn - nproc + Cp;
            if (nnew > ncap) { if (tid == 0) atomicOr(b.err, ORBX_DEVERR_NODE_OVERFLOW); break; }
            {
                int pc = (int)(pre & 0xFFFFF), pe = (int)(pre >> 40);
                for (int it = 0; it < ipt; it++) {
                    const int j = tid * ipt + it;
                    if (j < nvec && j <= jstar) {
                        const int d = nd_depth(Lcur[lpos[it]]) + 1;
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const int s = bd[it][k + 1] - bd[it][k];
                            if (s > 0) {
                                const int pos = Cp - 1 - pc;
                                Lnew[pos] = mk_node(bd[it][k], bd[it][k + 1], d);
                                if (s > 1) Vnew[pe++] = (unsigned)pos;
                                pc++;
                            }
                        }
                    }
                }
            }
            // untouched nodes keep their order behind the new children
            {
                u64 keep = 0;
                for (int it = 0; it < ipt; it++) { const int i = tid * ipt + it; if (i < nn && !proc[i]) keep++; }
                u64 tk;
                int pk = (int)block_scan_excl(keep, &tk, sh.warp_sums);
                for (int it = 0; it < ipt; it++) {
                    const int i = tid * ipt + it;
                    if (i < nn && !proc[i]) Lnew[Cp + pk++] = Lcur[i];
                }
            }
            __syncthreads();
            { u64* t = Lcur; Lcur = Lnew; Lnew = t; }
            { unsigned* t = Vcur; Vcur = Vnew; Vnew = t; }
            nn = nnew; nvec = Ep;
            if (nn >= N || nn == prev) finish = true;                 // :732
        }
    }

    // ---- best point per node, list order (:742-758) ----
    if (nn > L.kp_cap) { if (tid == 0) atomicOr(b.err, ORBX_DEVERR_KP_OVERFLOW); nn = L.kp_cap; }
    for (int i = tid; i < nn; i += NT) {
        const u64 nd = Lcur[i];
        unsigned best = 0;
        for (int k = nd_lo(nd); k < nd_hi(nd); k++) { const unsigned v = (unsigned)buf[k]; best = max(best, v); }
        out_kp[i] = pts[0xFFFFFF - (best & 0xFFFFFF)];
    }
    if (tid == 0) *out_n = nn;
}

struct OctCfg { int smem_pts, ncap; size_t smem; };

OctCfg octree_cfg(const OrbxGeom& g)
{
    OctCfg c;
    int ncap = 64;
    for (int l = 0; l < g.nlevels; l++) {
        const int need = (g.lv[l].quota > 4 * g.lv[l].nIni ? g.lv[l].quota : 4 * g.lv[l].nIni) + 8;
        if (need > ncap) ncap = need;
    }
    c.ncap = ncap;
    c.smem_pts = 4096;
    c.smem = (size_t)c.smem_pts * 12 + (size_t)ncap * (8 + 8 + 4 + 4 + 8 + 1) + 64;
    return c;
}

}  // namespace

int orbx_octree_smem_bytes(const OrbxGeom& g) { return (int)octree_cfg(g).smem; }

void orbx_launch_octree(const OrbxGeom& g, const OrbxBuffers& b, int batch, cudaStream_t s, int l_begin, int l_end)
{
    const OctCfg c = octree_cfg(g);
    if (l_end < 0 || l_end > g.nlevels) l_end = g.nlevels;
    if (l_end <= l_begin) return;
    dim3 grid(l_end - l_begin, batch);
    ORBX_OPTIN_SMEM(k_octree);
    static const int allow_radix = getenv("ORBX_OCTREE_BITONIC") ? 0 : 1;      // comparison runs: the bitonic network everywhere
    orbx_launch_pdl(k_octree, grid, dim3(NT), (size_t)c.smem, s, g, b, c.smem_pts, c.ncap, l_begin, allow_radix);
    ORBX_COUNT_LAUNCH(1);
}

void orbx_octree_configure(const OrbxGeom& g)
{
    (void)g;
    ORBX_OPTIN_SMEM(k_octree);
}
