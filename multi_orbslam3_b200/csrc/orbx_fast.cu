// orbx_fast.cu - grid FAST-9/16 with the iniThFAST / minThFAST fallback.
//
// Replaces the cell loop of ORBextractor::ComputeKeyPointsOctTree (R/src/ORBextractor.cc:787-854) and the
// cv::FAST(cell, th, nonmax=true) calls inside it (:808, :827).
//
// One CTA owns one SEGMENT (a run of whole cells, at most ORBX_FAST_TP staged columns) of one row of cells of one
// level of one frame and never writes a score map to global memory.  Segments bound the shared-memory footprint
// independently of the image width (44 KB: 5 CTAs per SM) and every phase runs once over the whole tile:
//   1. rows [iniY, maxY) x columns [xa0, xa1) are staged in shared memory by bulk copies (x index = column - xa0);
//   2. SWAR screen, two 4-pixel groups per lane-step: |v - ring| per byte (VABSDIFF4.U8) on the 4 compass/diagonal
//      opposite pairs; a 9-arc contains one pixel of every opposite pair, so a corner at threshold t needs
//      max(|d_k|, |d_k+8|) > t for every pair.  Groups with a survivor go to a group queue and are expanded into a
//      dense pixel queue (~13 % of the pixels of the bench frames);
//   3. queue drain, one pixel per thread: signed test on all 8 pairs, then the exact arc measure
//      m = max over the 16 arcs of 9 of max(min d, min -d) on packed 16x2 lanes (VIMNMX.U16x2).
//      corner <=> m > t, cv::FAST response = m - 1, independent of t.  m is kept in a shared u8 map;
//   4. non-max suppression inside the cell: the reference's NMS only sees scores of the same FAST call (= the
//      same cell) and non-corners score 0, so a pixel survives at threshold t <=> m > t and m is a strict
//      maximum among its in-cell neighbours' m: ONE suppression pass serves both thresholds.  Survivors are
//      recorded in two bitmaps (minThFAST / iniThFAST);
//   5. cells with no iniThFAST survivor fall back to their minThFAST survivors (:825-828); every survivor
//      computes its own output slot from popcounts (cell by cell, row-major inside a cell = reference order).
#include <cmath>
#include <cstdlib>
#include "orbx_internal.h"

namespace {

constexpr int NT = 256;
constexpr int NW = NT / 32;
constexpr int TP = ORBX_FAST_TP;     // shared-memory pitch of a segment tile (compile time: ring offsets become immediates)
constexpr int MAX_SEG_CELLS = 16;    // cells per segment (a cell is at least 30 columns wide: TP / 30 rounded up)
constexpr int CLCAP = 1024;          // corners (m > minTh) of the whole tile; overflow -> map scan (still exact)
constexpr int QMIN = 1024;           // smallest pixel queue the launch is sized for (overflow is scored inline, never dropped)
constexpr int SMEM_TARGET = 44 * 1024;   // 5 CTAs per SM

// ---- bulk asynchronous copy (TMA engine, 1-D form: SASS UBLKCP) + mbarrier ----
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

// shared-memory atomic add as ONE instruction (the CUDA intrinsic expands to a warp-aggregation sequence)
__device__ __forceinline__ int atoms_add(int* p, int v)
{
    int old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
    return old;
}

// ring offsets (dx,dy), OpenCV order
__host__ __device__ constexpr int ring_off(int k)
{
    constexpr int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
    constexpr int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
    return dy[k] * TP + dx[k];
}

// Arc measure on packed 16-bit lanes.  p[k] holds both polarities of ring pixel k, biased by 256 so that
// both lanes stay positive: low = v - r_k + 256 (= d_k + 256), high = r_k - v + 256 (= -d_k + 256).
// One unsigned 16x2 min network (windows of 9 by doubling: 2, 4, 8, +1) then serves "9 darker" and
// "9 brighter" at once: m = max over arcs of max(min d, min -d).
// NOTE: the scalar formulation max(min(...), -max(...)) is MIScompiled by ptxas 12.9 / the 580 driver JIT for
// sm_100 (verified on B200: -Xptxas -O0 and -G give the right answer, -O1..-O3 return max(d)); see
// profiles/ptxas_minmax_bug/ and DESIGN.md.
__device__ __forceinline__ int arc_measure(const unsigned (&p)[16])
{
    unsigned l2[16], l4[16];
#pragma unroll
    for (int k = 0; k < 16; k++) l2[k] = __vminu2(p[k], p[(k + 1) & 15]);
#pragma unroll
    for (int k = 0; k < 16; k++) l4[k] = __vminu2(l2[k], l2[(k + 2) & 15]);
    unsigned best = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) best = __vmaxu2(best, __vminu2(__vminu2(l4[k], l4[(k + 4) & 15]), p[(k + 8) & 15]));
    const int a = (int)(best & 0xFFFF), b = (int)(best >> 16);
    return (a > b ? a : b) - 256;
}

// packed ring differences of one pixel: pk[q] = (v + 256 - r_q) | (256 - v + r_q) << 16
__device__ __forceinline__ void load_ring(const uint8_t* p, unsigned (&pk)[16])
{
    const int v = p[0];
    const unsigned cv = (unsigned)(v + 256) | ((unsigned)(256 - v) << 16);
#pragma unroll
    for (int q = 0; q < 16; q++) pk[q] = cv + (unsigned)p[ring_off(q)] * 0xFFFFu;   // immediate offsets
}

// OpenCV's high-speed test on the 8 opposite pairs, both polarities at once and branch-free:
// "every pair has a member darker than v - t" <=> min_k max(d_k, d_k+8) > t (low lanes); brighter: high lanes.
__device__ __forceinline__ bool pair_test(const unsigned (&pk)[16], int minTh)
{
    unsigned mn = __vmaxu2(pk[0], pk[8]);
#pragma unroll
    for (int k = 1; k < 8; k++) mn = __vminu2(mn, __vmaxu2(pk[k], pk[k + 8]));
    const int lim = 256 + minTh;
    return (int)(mn & 0xFFFF) > lim || (int)(mn >> 16) > lim;
}

// exact per-pixel path used when a queue overflows: pair test, then the arc measure
__device__ __forceinline__ void score_pixel(const uint8_t* T, uint8_t* M, int x, int yt, int minTh,
                                            int* cl_count)
{
    unsigned pk[16];
    load_ring(T + yt * TP + x, pk);
    if (!pair_test(pk, minTh)) return;
    const int m = arc_measure(pk);
    if (m > minTh) { M[(yt - 3) * TP + x] = (uint8_t)m; atoms_add(cl_count, CLCAP + 1); }   // forces the exact map-scan NMS path
}

// in-cell non-max suppression of one corner (tile column x, scored row r): neighbours outside the cell count as 0
// (they belong to another FAST call in the reference)
__device__ __forceinline__ void nms_pixel(const uint8_t* M, unsigned* Bmin, unsigned* Bini, int x, int r, int X0, int X1, int hs,
                                          int wCell, unsigned wrcp, int iniTh)
{
    const uint8_t* qm = M + r * TP + x;
    const int sc = qm[0];
    const int j = (int)(((unsigned)(x - X0) * wrcp) >> 16);
    const int c0 = X0 + j * wCell, c1 = min(c0 + wCell, X1);     // cell interior [c0, c1)
    const bool hl = x - 1 >= c0, hr = x + 1 < c1, vu = r > 0, vd = r + 1 < hs;
    const int l0 = hl ? qm[-1] : 0, r0 = hr ? qm[1] : 0;
    const int u0 = vu ? qm[-TP] : 0, ul = (vu && hl) ? qm[-TP - 1] : 0, ur = (vu && hr) ? qm[-TP + 1] : 0;
    const int d0 = vd ? qm[TP] : 0, dl = (vd && hl) ? qm[TP - 1] : 0, dr = (vd && hr) ? qm[TP + 1] : 0;
    const int mx = max(max(max(l0, r0), max(u0, ul)), max(max(ur, d0), max(dl, dr)));
    if (sc > mx) {
        constexpr int bw = TP / 32;
        atomicOr(&Bmin[r * bw + (x >> 5)], 1u << (x & 31));
        if (sc > iniTh) atomicOr(&Bini[r * bw + (x >> 5)], 1u << (x & 31));
    }
}

// per-byte flag (bit 7) of "a > t" for t <= 126: bytes < 128 carry into bit 7 when a + 127 - t >= 128
__device__ __forceinline__ unsigned gt_flags(unsigned a, unsigned c127mt) { return ((a & 0x7f7f7f7fu) + c127mt) | a; }

// number of set bits in columns [c0, c1) of a bitmap row, 0 <= c1 - c0 < 64 (a cell is narrower than 60 columns);
// the word after the range's first word is read unconditionally: the bitmaps are padded by one word
__device__ __forceinline__ int popc_range(const unsigned* row, int c0, int c1)
{
    const int w = c0 >> 5;
    const unsigned long long v = ((unsigned long long)row[w + 1] << 32 | row[w]) >> (c0 & 31);
    return __popcll(v & ((1ull << (c1 - c0)) - 1ull));
}

struct FastSmem {
    int qg_count, q_count, q2_count, cl_count;
    int cell_off[MAX_SEG_CELLS + 1];
    unsigned char use_ini[MAX_SEG_CELLS];
};

// group-queue capacity: every pair of groups of the tile (cannot overflow)
__host__ __device__ inline int qg_cap(int hs, int ng) { return (hs * (ng + 2) + 3) & ~3; }      // words: one 64-bit entry per pair of groups

// words of the two survivor bitmaps [hs][TP/32] + one padding word for popc_range, rounded to 16 bytes (zeroed together with M)
__host__ __device__ inline int bitmap_words(int hs) { return (2 * hs * (TP / 32) + 1 + 3) & ~3; }

// shared-memory bytes of one segment tile, without the pixel queue
__host__ __device__ inline int tile_bytes(int nrow, int hs, int ng, int ncell)
{
    return nrow * TP + hs * TP + bitmap_words(hs) * 4 + qg_cap(hs, ng) * 4 + CLCAP * 4 + ((ncell * hs * 2 + 15) & ~15);
}

__global__ void __launch_bounds__(NT, 5) k_fast_seg(OrbxGeom g, OrbxBuffers b, const uint8_t* level0, int pitch0,
                                                    long long stride0, int smem_total, int unit0)
{
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ FastSmem sh;
    __shared__ __align__(8) unsigned long long s_bar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int f = blockIdx.y;
    // per-segment record precomputed by the host (orbx_fast_units)
    const int unit = unit0 + blockIdx.x;                              // segment index inside the frame (a launch may cover one level only)
    const int4 u0 = __ldg(b.unit_tab + 4 * unit), u1 = __ldg(b.unit_tab + 4 * unit + 1), u2 = __ldg(b.unit_tab + 4 * unit + 2);
    const int l = u0.x, iniY = u0.y, nrow = u0.z, hs = u0.w;          // level, first staged row, staged rows, scored rows (tile rows 3 .. nrow-4)
    const int xa0 = u1.x, sw = u1.y, X0 = u1.z, X1 = u1.w;            // staged columns [xa0, xa0+sw), scored tile columns [X0, X1)
    const int ncell = u2.x, tbytes = u2.y;
    const unsigned wrcp = (unsigned)u2.z, prcp = (unsigned)u2.w;      // (n * wrcp) >> 16 == n / wCell; (n * prcp) >> 20 == n / (pairs of groups per row)
    orbx_pdl_prologue();                                              // (the segment table above is immutable after configure)
    int* row_count = b.row_count + (long long)f * g.total_rows + unit;
    if (hs <= 0) { if (tid == 0) *row_count = 0; return; }
    const int wCell = g.lv[l].wCell, row_cap = g.lv[l].row_cap;
    const int g0 = X0 >> 2, g1 = (X1 - 1) >> 2, ng = g1 - g0 + 1;      // 4-pixel groups holding scored columns
    constexpr int bw = TP / 32;                                        // bitmap words per row

    // ---- smem carve-up (per CTA: tiles of different levels have different shapes) ----
    uint8_t* T = smem;                                              // [nrow][TP] pixels
    uint8_t* M = T + nrow * TP;                                     // [hs][TP]   arc measure (0 = not a corner at minTh)
    unsigned* Bmin = reinterpret_cast<unsigned*>(M + hs * TP);      // [hs][bw] survivors at minTh
    unsigned* Bini = Bmin + hs * bw;                                // [hs][bw] (+ padding) survivors at iniTh
    unsigned* QG = Bmin + bitmap_words(hs);                         // [qgcap] pairs of 4-pixel groups with screen survivors (64-bit entries)
    unsigned* Q2 = QG;                                              // survivors of the signed pair test (QG is dead by then)
    const int qgcap = qg_cap(hs, ng), q2cap = qgcap;
    unsigned* CL = QG + qgcap;                                      // [CLCAP] corners: x | scored row << 16
    unsigned short* cnt_ini = reinterpret_cast<unsigned short*>(CL + CLCAP);   // [ncell][hs] exclusive row prefix of the chosen counts inside a cell
    unsigned* Q = reinterpret_cast<unsigned*>(smem + tbytes);       // [qcap] candidate queue: x | tile row << 16
    const int qcap = (smem_total - tbytes) >> 2;

    const uint8_t* img; int pitch;
    if (l == 0) { img = level0 + (long long)f * stride0; pitch = pitch0; }
    else { img = b.pyr[l] + (long long)f * g.lv[l].frame_stride; pitch = g.lv[l].pitch; }

    // ---- 1. stage rows ----
    const bool vec = ((pitch & 15) == 0) && (pitch >= xa0 + sw) && ((reinterpret_cast<uintptr_t>(img) & 15) == 0);
    if (vec) {
        // one bulk copy per row, issued by the lanes of warp 0; completion is counted in bytes on an mbarrier, and the
        // zero-fill of the score map / bitmaps below overlaps the copies
        if (tid == 0) mbar_init(&s_bar, 1);
        __syncthreads();
        if (warp == 0) {
            if (lane == 0) mbar_expect_tx(&s_bar, (unsigned)(nrow * sw));
            __syncwarp();
            for (int r = lane; r < nrow; r += 32) bulk_g2s(T + r * TP, img + (long long)(iniY + r) * pitch + xa0, (unsigned)sw, &s_bar);
        }
    } else {
        for (int r = warp; r < nrow; r += NW)
            for (int c = lane; c < sw; c += 32) T[r * TP + c] = xa0 + c < g.lv[l].w ? __ldg(img + (long long)(iniY + r) * pitch + xa0 + c) : 0;
    }
    {
        uint4* z = reinterpret_cast<uint4*>(M);                         // score map and both bitmaps are contiguous
        const int nz = (hs * TP + bitmap_words(hs) * 4) >> 4;
        for (int k = tid; k < nz; k += NT) z[k] = make_uint4(0, 0, 0, 0);
        if (tid == 0) { sh.cl_count = 0; sh.qg_count = 0; sh.q_count = 0; sh.q2_count = 0; }
    }
    const int minTh = g.min_th, iniTh = g.ini_th;
    if (vec) mbar_wait(&s_bar, 0);
    __syncthreads();

    // ---- 2. screen: items = (scored row, pair of adjacent 4-pixel groups), dealt linearly to the threads (no idle lanes at
    //      narrow levels).  Stage 1 (compass pairs) runs here on every item; an item with a survivor is queued as one 64-bit
    //      entry by a per-lane shared-memory atomic (the order of the queue is irrelevant: the output order comes from the
    //      bitmaps).  Stage 2 (diagonal pairs) and the segment-edge masks run in the expansion below, where every lane holds
    //      a queued item: dense lanes instead of the 12-of-32 of a branch inside this loop. ----
    const unsigned c127 = (unsigned)(127 - (minTh < 126 ? minTh : 126)) * 0x01010101u;
    const bool screen_ok = minTh <= 126;
    constexpr int W = TP / 4;
    const int k0 = g0 >> 1, npairs = (g1 >> 1) - k0 + 1;          // pairs of groups (8-byte aligned)
    uint2* QG2 = reinterpret_cast<uint2*>(QG);
    {
        const int nitems = hs * npairs;
        const unsigned* Tw = reinterpret_cast<const unsigned*>(T + 3 * TP);
        for (int i = tid; i < nitems; i += NT) {
            const int ry = (int)(((unsigned)i * prcp) >> 20), kp = k0 + i - ry * npairs;
            const unsigned* rc = Tw + ry * W + 2 * kp;
            unsigned ca = 0x80808080u, cb = 0x80808080u;
            if (screen_ok) {
                // |d_k| | |d_k+8| >= max(|d_k|, |d_k+8|): one threshold test per pair, still only a necessary
                // condition (exact when t = 2^n - 1, e.g. the reference's minThFAST = 7)
                const uint2 V = *reinterpret_cast<const uint2*>(rc);
                const uint2 P3 = *reinterpret_cast<const uint2*>(rc + 3 * W), M3 = *reinterpret_cast<const uint2*>(rc - 3 * W);
                const unsigned Lw = rc[-1], Rw = rc[2];
                // pair (0, 8): (0,+3) / (0,-3);  pair (4, 12): (+3,0) / (-3,0)
                ca = gt_flags(__vabsdiffu4(V.x, P3.x) | __vabsdiffu4(V.x, M3.x), c127);
                cb = gt_flags(__vabsdiffu4(V.y, P3.y) | __vabsdiffu4(V.y, M3.y), c127);
                ca &= gt_flags(__vabsdiffu4(V.x, __funnelshift_r(V.x, V.y, 24)) | __vabsdiffu4(V.x, __funnelshift_r(Lw, V.x, 8)), c127);
                cb &= gt_flags(__vabsdiffu4(V.y, __funnelshift_r(V.y, Rw, 24)) | __vabsdiffu4(V.y, __funnelshift_r(V.x, V.y, 8)), c127);
                ca &= 0x80808080u; cb &= 0x80808080u;
            }
            if (ca | cb) {
                // entry.x = flags of the first group (bits 7, 15, 23, 31) | pair index (bits 0-6) | tile row (bits 8-14); entry.y = flags of the second
                const int o = atoms_add(&sh.qg_count, 1);                 // <= hs * npairs entries by construction
                QG2[o] = make_uint2(ca | (unsigned)kp | ((unsigned)(ry + 3) << 8), cb);
            }
        }
    }
    __syncthreads();
    // ---- stage 2 of the screen on the queued items, then expansion into pixel entries (dense queue for phase 3a) ----
    {
        const int ngq = sh.qg_count;
        // scored columns only: the first and last group straddle the segment's edges
        const unsigned mask0 = 0x80808080u << (8 * (X0 & 3)), mask1 = 0x80808080u >> (8 * (3 - ((X1 - 1) & 3)));
        for (int gi = tid; gi < ngq; gi += NT) {
            unsigned ca, cb, e;
            {
                const uint2 ge = QG2[gi];
                const int kp = ge.x & 0x7F, yt = (ge.x >> 8) & 0x7F;
                ca = ge.x & 0x80808080u; cb = ge.y;
                if (screen_ok) {
                    const unsigned* rc = reinterpret_cast<const unsigned*>(T + yt * TP) + 2 * kp;
                    const uint2 V = *reinterpret_cast<const uint2*>(rc);
                    // pair (2, 10): (+2,+2) / (-2,-2);  pair (6, 14): (+2,-2) / (-2,+2)
                    const uint2 P2 = *reinterpret_cast<const uint2*>(rc + 2 * W), M2 = *reinterpret_cast<const uint2*>(rc - 2 * W);
                    const unsigned P2l = rc[2 * W - 1], P2r = rc[2 * W + 2], M2l = rc[-2 * W - 1], M2r = rc[-2 * W + 2];
                    ca &= gt_flags(__vabsdiffu4(V.x, __funnelshift_r(P2.x, P2.y, 16)) | __vabsdiffu4(V.x, __funnelshift_r(M2l, M2.x, 16)), c127);
                    ca &= gt_flags(__vabsdiffu4(V.x, __funnelshift_r(M2.x, M2.y, 16)) | __vabsdiffu4(V.x, __funnelshift_r(P2l, P2.x, 16)), c127);
                    cb &= gt_flags(__vabsdiffu4(V.y, __funnelshift_r(P2.y, P2r, 16)) | __vabsdiffu4(V.y, __funnelshift_r(M2.x, M2.y, 16)), c127);
                    cb &= gt_flags(__vabsdiffu4(V.y, __funnelshift_r(M2.y, M2r, 16)) | __vabsdiffu4(V.y, __funnelshift_r(P2.x, P2.y, 16)), c127);
                }
                const int ga = 2 * kp;
                if (ga < g0) ca = 0;
                if (ga == g0) ca &= mask0;
                if (ga == g1) ca &= mask1;
                if (ga + 1 > g1) cb = 0;
                if (ga + 1 == g0) cb &= mask0;
                if (ga + 1 == g1) cb &= mask1;
                e = ((unsigned)ga << 2) | ((unsigned)yt << 16);
            }
            const int n = __popc(ca) + __popc(cb);
            if (!n) continue;
            int o = atoms_add(&sh.q_count, n);                          // per-lane: the queue order is irrelevant
            if (o + n <= qcap) {
#pragma unroll
                for (int q = 0; q < 8; q++)
                    if ((q < 4 ? ca : cb) & (0x80u << (8 * (q & 3)))) Q[o++] = e + q;
            } else {
#pragma unroll 1
                for (int q = 0; q < 8; q++)
                    if ((q < 4 ? ca : cb) & (0x80u << (8 * (q & 3)))) {
                        if (o < qcap) Q[o] = e + q;
                        else score_pixel(T, M, (e + q) & 0xFFFF, (int)(e >> 16), minTh, &sh.cl_count);   // queue full: score inline
                        o++;
                    }
            }
        }
    }
    __syncthreads();
    // ---- 3a. signed pair test, dense and branch-free; survivors are compacted into Q2 ----
    {
        const int nq = min(sh.q_count, qcap);
        for (int e = tid; e < nq; e += NT) {
            const unsigned ent = Q[e];
            unsigned pk[16];
            load_ring(T + (ent >> 16) * TP + (ent & 0xFFFF), pk);
            if (pair_test(pk, minTh)) {
                const int o = atoms_add(&sh.q2_count, 1);
                if (o < q2cap) Q2[o] = ent;
                else score_pixel(T, M, ent & 0xFFFF, ent >> 16, minTh, &sh.cl_count);
            }
        }
    }
    __syncthreads();
    // ---- 3b. exact arc measure for the survivors ----
    {
        const int nq2 = min(sh.q2_count, q2cap);
        for (int e = tid; e < nq2; e += NT) {
            const unsigned ent = Q2[e];
            const int x = ent & 0xFFFF, yq = ent >> 16;
            unsigned pk[16];
            load_ring(T + yq * TP + x, pk);
            const int m = arc_measure(pk);
            if (m > minTh) {
                M[(yq - 3) * TP + x] = (uint8_t)m;
                const int o = atoms_add(&sh.cl_count, 1);
                if (o < CLCAP) CL[o] = (unsigned)x | ((unsigned)(yq - 3) << 16);
            }
        }
    }
    __syncthreads();

    // ---- 4. in-cell non-max suppression ----
    // Corners were listed by phase 3b, so the suppression runs one corner per thread instead of diverging over a
    // sparse map; a corner-dense tile (list overflow) falls back to scanning the map, still exact.
    if (sh.cl_count <= CLCAP) {
        const int ncl = sh.cl_count;
        for (int e = tid; e < ncl; e += NT) {
            const unsigned ent = CL[e];
            const int x = ent & 0xFFFF, r = ent >> 16;
            nms_pixel(M, Bmin, Bini, x, r, X0, X1, hs, wCell, wrcp, iniTh);
        }
    } else {
        for (int r = warp; r < hs; r += NW)
            for (int wx = lane; wx < TP / 4; wx += 32) {
                const unsigned word = reinterpret_cast<const unsigned*>(M + r * TP)[wx];
                if (!word) continue;
                for (int q = 0; q < 4; q++)
                    if ((word >> (8 * q)) & 0xFF) nms_pixel(M, Bmin, Bini, (wx << 2) + q, r, X0, X1, hs, wCell, wrcp, iniTh);
            }
    }
    __syncthreads();

    // ---- 5. counts per (cell, row), threshold choice per cell, offsets: one warp per cell, one lane per scored row ----
    for (int j = warp; j < ncell; j += NW) {
        const int c0 = X0 + j * wCell, c1 = min(c0 + wCell, X1);
        int run = 0, ti = 0;
        // first sweep: does the cell have an iniThFAST survivor at all?
        for (int r0 = 0; r0 < hs; r0 += 32) {
            const int r = r0 + lane;
            const int bq = (r < hs && c1 > c0) ? popc_range(Bini + r * bw, c0, c1) : 0;
            ti += __popc(__ballot_sync(0xffffffffu, bq > 0));
        }
        const bool ui = ti > 0;
        const unsigned* Bsel = ui ? Bini : Bmin;
        for (int r0 = 0; r0 < hs; r0 += 32) {
            const int r = r0 + lane;
            const int c = (r < hs && c1 > c0) ? popc_range(Bsel + r * bw, c0, c1) : 0;
            int inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            if (r < hs) cnt_ini[j * hs + r] = (unsigned short)(run + inc - c);      // exclusive prefix inside the cell
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) { sh.use_ini[j] = ui; sh.cell_off[j + 1] = run; }            // cell totals, scanned below
    }
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        sh.cell_off[0] = 0;
        for (int j = 0; j < ncell; j++) { run += sh.cell_off[j + 1]; sh.cell_off[j + 1] = run; }   // cell_off[j] = first slot of cell j
        *row_count = min(run, row_cap);
        if (run > row_cap) atomicOr(b.err, ORBX_DEVERR_CAND_OVERFLOW);
    }
    __syncthreads();

    // ---- ordered emission: every survivor computes its own slot ----
    uint32_t* out = b.row_cand + (long long)f * b.row_cand_stride + b.row_off[unit];
    const int yrel0 = iniY - ORBX_BORDER + 3;
    if (sh.cl_count <= CLCAP) {
        // one listed corner per thread: survivors of the NMS look up their own slot (dense lanes; the bitmap walk below
        // runs at ~5 threads per instruction)
        const int ncl = sh.cl_count;
        for (int e = tid; e < ncl; e += NT) {
            const unsigned ent = CL[e];
            const int x = ent & 0xFFFF, r = ent >> 16;
            const int j = (int)(((unsigned)(x - X0) * wrcp) >> 16);
            const unsigned* Bsel = (sh.use_ini[j] ? Bini : Bmin) + r * bw;
            if (!((Bsel[x >> 5] >> (x & 31)) & 1u)) continue;
            const int c0 = X0 + j * wCell;
            const int o = sh.cell_off[j] + cnt_ini[j * hs + r] + (x > c0 ? popc_range(Bsel, c0, x) : 0);
            if (o < row_cap)
                out[o] = (uint32_t)(x + xa0 - ORBX_BORDER) | ((uint32_t)(yrel0 + r) << 12) | ((uint32_t)(M[r * TP + x] - 1) << 24);
        }
        return;
    }
    for (int k = tid; k < hs * bw; k += NT) {
        const int r = k / bw, wi = k - r * bw;                 // bw is a compile-time constant
        unsigned bits = Bmin[k];
        const unsigned ibits = Bini[k];
        while (bits) {
            const int bpos = __ffs(bits) - 1;
            bits &= bits - 1;
            const int x = (wi << 5) + bpos;
            const int j = (int)(((unsigned)(x - X0) * wrcp) >> 16);
            const bool ui = sh.use_ini[j];
            if (ui && !((ibits >> bpos) & 1u)) continue;
            const int c0 = X0 + j * wCell;
            const int rank = x > c0 ? popc_range((ui ? Bini : Bmin) + r * bw, c0, x) : 0;
            const int o = sh.cell_off[j] + cnt_ini[j * hs + r] + rank;
            if (o < row_cap)
                out[o] = (uint32_t)(x + xa0 - ORBX_BORDER) | ((uint32_t)(yrel0 + r) << 12) | ((uint32_t)(M[r * TP + x] - 1) << 24);
        }
    }
}

// shared memory of the launch: the largest tile of any level plus a pixel queue, rounded up to the 4-CTA target
size_t fast_smem_bytes(const OrbxGeom& g)
{
    size_t need = 0;
    for (int l = 0; l < g.nlevels; l++) {
        const OrbxLevel& L = g.lv[l];
        if (L.nCols <= 0 || L.nRows <= 0) continue;
        const int hs = L.hCell, ng = TP / 4;
        const size_t n = (size_t)tile_bytes(hs + 6, hs, ng, L.segCells) + QMIN * 4;
        if (n > need) need = n;
    }
    static const int target = getenv("ORBX_FAST_SMEM_KB") ? atoi(getenv("ORBX_FAST_SMEM_KB")) * 1024 : SMEM_TARGET;
    return need > (size_t)target ? need : (size_t)target;
}

}  // namespace

// Segments of a cell row: the fewest equal runs of whole cells whose staged width (3-pixel halo, 16-byte aligned on
// both sides) fits the tile pitch.  Returns cells per segment.
int orbx_fast_plan(int w, int nCols, int wCell)
{
    if (nCols <= 0) return 0;
    for (int ns = 1; ns <= nCols; ns++) {
        const int sc = (nCols + ns - 1) / ns;
        bool ok = true;
        for (int j0 = 0; j0 < nCols && ok; j0 += sc) {
            const int j1 = j0 + sc < nCols ? j0 + sc : nCols;
            const int cs0 = ORBX_EDGE + j0 * wCell;
            int cs1 = ORBX_EDGE + j1 * wCell; if (cs1 > w - ORBX_EDGE) cs1 = w - ORBX_EDGE;
            if (cs1 <= cs0) continue;
            const int xa0 = (cs0 - 3) & ~15, xa1 = (cs1 + 3 + 15) & ~15;
            if (xa1 - xa0 > TP || sc > MAX_SEG_CELLS) ok = false;
        }
        if (ok) return sc;
    }
    return 1;
}

// Per-segment records read by k_fast_seg, 4 x int4 per segment in launch order (level, cell row, segment):
//   {level, first staged row, staged rows, scored rows (<= 0: empty segment)}
//   {xa0, staged width, X0, X1}     {cells, tile bytes, 2^16 / wCell, 2^20 / screen items per row}   {2^16 / scored rows, -, -, -}
void orbx_fast_units(const OrbxGeom& g, std::vector<int4>& tab)
{
    tab.clear();
    for (int l = 0; l < g.nlevels; l++) {
        const OrbxLevel& L = g.lv[l];
        for (int i = 0; i < L.nRows; i++)
            for (int sg = 0; sg < L.nSeg; sg++) {
                int4 a = make_int4(l, 0, 0, 0), bq = make_int4(0, 0, 0, 0), c = make_int4(0, 0, 0, 0);
                const int iniY = ORBX_BORDER + i * L.hCell;
                int maxY = iniY + L.hCell + 6; if (maxY > L.maxBY) maxY = L.maxBY;
                const int nrow = maxY - iniY, hs = nrow - 6;
                const int j0 = sg * L.segCells, ncell = L.segCells < L.nCols - j0 ? L.segCells : L.nCols - j0;
                const int xs1 = L.w - ORBX_EDGE;
                const int cs0 = ORBX_EDGE + j0 * L.wCell;
                int cs1 = cs0 + ncell * L.wCell; if (cs1 > xs1) cs1 = xs1;
                if (iniY < L.maxBY - 3 && hs > 0 && cs1 > cs0) {
                    const int xa0 = (cs0 - 3) & ~15, xa1 = (cs1 + 3 + 15) & ~15;
                    const int X0 = cs0 - xa0, X1 = cs1 - xa0;
                    const int ng = ((X1 - 1) >> 2) - (X0 >> 2) + 1;
                    const int npairs = (((X1 - 1) >> 2) >> 1) - ((X0 >> 2) >> 1) + 1;              // screen items per row: 2 groups each
                    a = make_int4(l, iniY, nrow, hs);
                    bq = make_int4(xa0, xa1 - xa0, X0, X1);
                    c = make_int4(ncell, tile_bytes(nrow, hs, ng, ncell), (65536 + L.wCell - 1) / L.wCell, ((1 << 20) + npairs - 1) / npairs);
                }
                tab.push_back(a); tab.push_back(bq); tab.push_back(c);
                tab.push_back(make_int4(hs > 0 ? (65536 + hs - 1) / hs : 0, 0, 0, 0));
            }
    }
}

// host-only introspection of the segment plan of one pyramid level (no device needed; used by the CPU test-suite)
extern "C" int orbx_fast_segment_plan(int level_width, int* n_cols, int* w_cell, int* seg_cells, int* n_seg, int* max_staged_width)
{
    if (level_width < 8 || !n_cols || !w_cell || !seg_cells || !n_seg || !max_staged_width) return ORBX_E_INVALID;
    const float fw = (float)(level_width - 2 * ORBX_BORDER);
    const int nc = fw > 0 ? (int)(fw / (float)ORBX_FAST_W) : 0;                       // R/src/ORBextractor.cc:779-785
    *n_cols = nc; *w_cell = 0; *seg_cells = 0; *n_seg = 0; *max_staged_width = 0;
    if (nc <= 0) return ORBX_OK;
    const int wc = (int)ceilf(fw / nc);
    const int sc = orbx_fast_plan(level_width, nc, wc);
    *w_cell = wc; *seg_cells = sc; *n_seg = (nc + sc - 1) / sc;
    for (int j0 = 0; j0 < nc; j0 += sc) {
        const int j1 = j0 + sc < nc ? j0 + sc : nc;
        const int cs0 = ORBX_EDGE + j0 * wc;
        int cs1 = ORBX_EDGE + j1 * wc; if (cs1 > level_width - ORBX_EDGE) cs1 = level_width - ORBX_EDGE;
        if (cs1 <= cs0) continue;
        const int wst = ((cs1 + 3 + 15) & ~15) - ((cs0 - 3) & ~15);
        if (wst > *max_staged_width) *max_staged_width = wst;
    }
    return ORBX_OK;
}

void orbx_fast_configure(const OrbxGeom& g)
{
    (void)g;
    ORBX_OPTIN_SMEM(k_fast_seg);
}

void orbx_launch_fast(const OrbxGeom& g, const OrbxBuffers& b, const uint8_t* level0, int pitch0,
                      long long stride0, int batch, cudaStream_t s, int l_begin, int l_end)
{
    if (l_end < 0 || l_end > g.nlevels) l_end = g.nlevels;
    // the segments of a level are contiguous in the table: [row_base, row_base + nRows * nSeg)
    const int unit0 = l_begin < g.nlevels ? g.lv[l_begin].row_base : g.total_rows;
    const int unit1 = l_end < g.nlevels ? g.lv[l_end].row_base : g.total_rows;
    if (unit1 <= unit0) return;
    dim3 grid(unit1 - unit0, batch);
    const size_t smem = fast_smem_bytes(g);
    ORBX_OPTIN_SMEM(k_fast_seg);
    orbx_launch_pdl(k_fast_seg, grid, dim3(NT), smem, s, g, b, level0, pitch0, stride0, (int)smem, unit0);
    ORBX_COUNT_LAUNCH(1);
}
