// orbx_fast.cu - grid FAST-9/16 with the iniThFAST / minThFAST fallback.
//
// Replaces the cell loop of ORBextractor::ComputeKeyPointsOctTree (R/src/ORBextractor.cc:787-854) and the
// cv::FAST(cell, th, nonmax=true) calls inside it (:808, :827).
//
// One CTA owns one row of cells of one level of one frame and never writes a score map to global memory:
//   1. rows [iniY, maxY) are staged in shared memory with 16-byte loads (x index = absolute column);
//   2. SWAR screen, 4 pixels per thread-step: |v - ring| per byte (VABSDIFF4.U8) on the 4 compass/diagonal
//      opposite pairs; a 9-arc contains one pixel of every opposite pair, so a corner at threshold t needs
//      max(|d_k|, |d_k+8|) > t for every pair.  Survivors (a few % of pixels) go to a shared-memory queue;
//   3. queue drain, one pixel per thread: signed test on all 8 pairs, then the exact arc measure
//      m = max over the 16 arcs of 9 of max(min d, min -d) on packed 16x2 lanes (VIMNMX.U16x2).
//      corner <=> m > t, cv::FAST response = m - 1, independent of t.  m is kept in a shared u8 map;
//   4. non-max suppression inside the cell: the reference's NMS only sees scores of the same FAST call (= the
//      same cell) and non-corners score 0, so a pixel survives at threshold t <=> m > t and m is a strict
//      maximum among its in-cell neighbours' m: ONE suppression pass serves both thresholds.  Survivors are
//      recorded in two bitmaps (minThFAST / iniThFAST);
//   5. cells with no iniThFAST survivor fall back to their minThFAST survivors (:825-828); every survivor
//      computes its own output slot from popcounts (cell by cell, row-major inside a cell = reference order).
#include "orbx_internal.h"

namespace {

constexpr int NT = 512;
constexpr int MAX_CELLS = 128;
constexpr int NW = NT / 32;
constexpr int QGCAP = 4096;          // 4-pixel groups with a screen survivor per band: 16 rows x <= 256 groups, cannot overflow
constexpr int QCAP = 4096;           // screen survivors (pixels) per band (overflow is scored inline, never dropped)
constexpr int Q2CAP = 2048;          // signed-test survivors per band (same overflow rule)
constexpr int CLCAP = 2048;          // corners (m > minTh) of the whole cell row; overflow -> map scan (still exact)
constexpr int BAND_ROWS = NW;        // one tile row per warp and band

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
__device__ __constant__ int8_t c_ring[16][2] = {
    {0, 3}, {1, 3}, {2, 2}, {3, 1}, {3, 0}, {3, -1}, {2, -2}, {1, -3},
    {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

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
__device__ __forceinline__ void load_ring(const uint8_t* p, const int (&roff)[16], unsigned (&pk)[16])
{
    const int v = p[0];
    const unsigned cv = (unsigned)(v + 256) | ((unsigned)(256 - v) << 16);
#pragma unroll
    for (int q = 0; q < 16; q++) pk[q] = cv + (unsigned)p[roff[q]] * 0xFFFFu;
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
__device__ __forceinline__ void score_pixel(const uint8_t* T, uint8_t* M, int tp, const int (&roff)[16], int x, int yt, int minTh,
                                            int* cl_count)
{
    unsigned pk[16];
    load_ring(T + yt * tp + x, roff, pk);
    if (!pair_test(pk, minTh)) return;
    const int m = arc_measure(pk);
    if (m > minTh) { M[(yt - 3) * tp + x] = (uint8_t)m; atoms_add(cl_count, CLCAP + 1); }   // forces the exact map-scan NMS path
}

// per-byte flag (bit 7) of "a > t" for t <= 126: bytes < 128 carry into bit 7 when a + 127 - t >= 128
__device__ __forceinline__ unsigned gt_flags(unsigned a, unsigned c127mt) { return ((a & 0x7f7f7f7fu) + c127mt) | a; }

// number of set bits in columns [c0, c1) of a bitmap row
__device__ __forceinline__ int popc_range(const unsigned* row, int c0, int c1)
{
    int n = 0;
    for (int w = c0 >> 5; w <= (c1 - 1) >> 5; w++) {
        unsigned v = row[w];
        const int lo = w << 5;
        if (c0 > lo) v &= 0xFFFFFFFFu << (c0 - lo);
        if (c1 < lo + 32) v &= 0xFFFFFFFFu >> (lo + 32 - c1);
        n += __popc(v);
    }
    return n;
}

struct FastSmem {
    int qg_count, q_count, q2_count, cl_count;
    int cell_off[MAX_CELLS + 1];
    unsigned char use_ini[MAX_CELLS];
};

__global__ void __launch_bounds__(NT) k_fast_rows(OrbxGeom g, OrbxBuffers b, const uint8_t* level0, int pitch0,
                                                  long long stride0)
{
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ FastSmem sh;

    const int tid = threadIdx.x;
    const int f = blockIdx.y;
    int l = 0;
    while (l + 1 < g.nlevels && (int)blockIdx.x >= g.lv[l + 1].row_base) l++;
    const OrbxLevel L = g.lv[l];
    const int i = blockIdx.x - L.row_base;
    int* row_count = b.row_count + (long long)f * g.total_rows + blockIdx.x;
    if (i >= L.nRows) return;
    const int iniY = ORBX_BORDER + i * L.hCell;
    int maxY = iniY + L.hCell + 6;
    if (iniY >= L.maxBY - 3) { if (tid == 0) *row_count = 0; return; }
    if (maxY > L.maxBY) maxY = L.maxBY;
    const int nrow = maxY - iniY;                      // staged rows
    const int hs = nrow - 6;                           // scored rows: tile rows 3 .. nrow-4
    const int xs0 = ORBX_EDGE, xs1 = L.w - ORBX_EDGE;  // scored columns [19, w-19)
    if (hs <= 0 || xs1 <= xs0) { if (tid == 0) *row_count = 0; return; }
    const int tp = (L.w + 15) & ~15;                   // smem pitch; x index = absolute column
    const int bw = (L.w + 31) >> 5;                    // bitmap words per row
    const int hmax = L.hCell;                          // scored rows of a full cell row of this level

    // ---- smem carve-up (sizes use the level's maxima so that every cell row has the same layout) ----
    uint8_t* T = smem;                                              // [hmax+6][tp] pixels
    uint8_t* M = T + (size_t)(hmax + 6) * tp;                       // [hmax][tp]   arc measure (0 = not a corner at minTh)
    unsigned* Q = reinterpret_cast<unsigned*>(M + (size_t)hmax * tp);   // [QCAP] candidate queue: x | tile row << 16
    unsigned* QG = Q + QCAP;                                        // [QGCAP] groups with survivors: gx | tile row << 12 | mask << 20
    unsigned* Q2 = QG + QGCAP;                                      // [Q2CAP] survivors of the signed pair test
    unsigned* CL = Q2 + Q2CAP;                                      // [CLCAP] corners: x | scored row << 16
    unsigned* Bmin = CL + CLCAP;                                    // [hmax][bw] survivors at minTh
    unsigned* Bini = Bmin + hmax * bw;                              // [hmax][bw] survivors at iniTh
    unsigned short* cnt_min = reinterpret_cast<unsigned short*>(Bini + hmax * bw);   // [nCols][hmax]
    unsigned short* cnt_ini = cnt_min + L.nCols * hmax;             // [nCols][hmax]; later: exclusive row prefix of the chosen counts

    const uint8_t* img; int pitch;
    if (l == 0) { img = level0 + (long long)f * stride0; pitch = pitch0; }
    else { img = b.pyr[l] + (long long)f * L.frame_stride; pitch = L.pitch; }

    // ---- 1. stage rows ----
    const bool vec = ((pitch & 15) == 0) && (pitch >= tp) && ((reinterpret_cast<uintptr_t>(img) & 15) == 0);
    const int lane = tid & 31, warp = tid >> 5;
    __shared__ __align__(8) unsigned long long s_bar;
    if (vec) {
        // one bulk copy per row, issued by the lanes of warp 0; completion is counted in bytes on an mbarrier, and the
        // zero-fill of the score map / bitmaps below overlaps the copies
        if (tid == 0) mbar_init(&s_bar, 1);
        __syncthreads();
        if (warp == 0) {
            if (lane == 0) mbar_expect_tx(&s_bar, (unsigned)(nrow * tp));
            __syncwarp();
            for (int r = lane; r < nrow; r += 32) bulk_g2s(T + r * tp, img + (long long)(iniY + r) * pitch, (unsigned)tp, &s_bar);
        }
    } else {
        for (int r = warp; r < nrow; r += NW)
            for (int c = lane; c < tp; c += 32) T[r * tp + c] = c < L.w ? __ldg(img + (long long)(iniY + r) * pitch + c) : 0;
    }
    {
        uint4* z = reinterpret_cast<uint4*>(M);
        const int nz = (hs * tp) >> 4;
        for (int k = tid; k < nz; k += NT) z[k] = make_uint4(0, 0, 0, 0);
        for (int k = tid; k < 2 * hmax * bw; k += NT) Bmin[k] = 0;
        if (tid == 0) sh.cl_count = 0;
    }
    const int minTh = g.min_th, iniTh = g.ini_th;
    int roff[16];
#pragma unroll
    for (int k = 0; k < 16; k++) roff[k] = c_ring[k][1] * tp + c_ring[k][0];
    if (vec) mbar_wait(&s_bar, 0);
    __syncthreads();

    // ---- 2 + 3. banded screen and drain ----
    const int gx0 = xs0 >> 2, gx1 = (xs1 - 1) >> 2;    // 4-pixel groups that contain scored columns
    const unsigned c127 = (unsigned)(127 - (minTh < 126 ? minTh : 126)) * 0x01010101u;
    const bool screen_ok = minTh <= 126;
    for (int y0 = 3; y0 < nrow - 3; y0 += BAND_ROWS) {
        if (tid == 0) { sh.qg_count = 0; sh.q_count = 0; sh.q2_count = 0; }
        __syncthreads();
        // -- screen: one tile row per warp, one 4-pixel group per lane-step --
        const int yt = y0 + warp;
        if (yt < nrow - 3) {
            const unsigned* rc = reinterpret_cast<const unsigned*>(T + yt * tp);
            const unsigned* rp3 = reinterpret_cast<const unsigned*>(T + (yt + 3) * tp);
            const unsigned* rm3 = reinterpret_cast<const unsigned*>(T + (yt - 3) * tp);
            const unsigned* rp2 = reinterpret_cast<const unsigned*>(T + (yt + 2) * tp);
            const unsigned* rm2 = reinterpret_cast<const unsigned*>(T + (yt - 2) * tp);
            for (int g0 = gx0; g0 <= gx1; g0 += 32) {      // warp-uniform trip count (ballots below)
                const int gx = g0 + lane;
                unsigned cand = 0;
                if (gx <= gx1) {
                    cand = 0x80808080u;
                    if (screen_ok) {
                        // |d_k| | |d_k+8| >= max(|d_k|, |d_k+8|): one threshold test per pair, still only a necessary
                        // condition (exact when t = 2^n - 1, e.g. the reference's minThFAST = 7)
                        const unsigned V = rc[gx];
                        // pair (0, 8): (0,+3) / (0,-3);  pair (4, 12): (+3,0) / (-3,0)
                        cand &= gt_flags(__vabsdiffu4(V, rp3[gx]) | __vabsdiffu4(V, rm3[gx]), c127);
                        cand &= gt_flags(__vabsdiffu4(V, __funnelshift_r(V, rc[gx + 1], 24)) |
                                         __vabsdiffu4(V, __funnelshift_r(rc[gx - 1], V, 8)), c127);
                        if (cand) {
                            // pair (2, 10): (+2,+2) / (-2,-2);  pair (6, 14): (+2,-2) / (-2,+2)
                            const unsigned p2c = rp2[gx], m2c = rm2[gx];
                            const unsigned a2 = __funnelshift_r(p2c, rp2[gx + 1], 16), a10 = __funnelshift_r(rm2[gx - 1], m2c, 16);
                            const unsigned a6 = __funnelshift_r(m2c, rm2[gx + 1], 16), a14 = __funnelshift_r(rp2[gx - 1], p2c, 16);
                            cand &= gt_flags(__vabsdiffu4(V, a2) | __vabsdiffu4(V, a10), c127);
                            cand &= gt_flags(__vabsdiffu4(V, a6) | __vabsdiffu4(V, a14), c127);
                        }
                    }
                    // keep scored columns only (only the first and last group straddle the border)
                    if (gx == gx0 || gx == gx1) {
                        const int xb = gx << 2;
#pragma unroll
                        for (int q = 0; q < 4; q++)
                            if (xb + q < xs0 || xb + q >= xs1) cand &= ~(0x80u << (8 * q));
                    }
                }
                // one queue entry per 4-pixel group that still has a candidate: one ballot, one atomic per warp-step
                const unsigned bal = __ballot_sync(0xffffffffu, cand != 0);
                if (bal) {
                    int base = 0;
                    if (lane == 0) base = atoms_add(&sh.qg_count, __popc(bal));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (cand) {
                        const unsigned m4 = ((cand >> 7) & 1u) | ((cand >> 14) & 2u) | ((cand >> 21) & 4u) | ((cand >> 28) & 8u);
                        const int o = base + __popc(bal & ((1u << lane) - 1));
                        if (o < QGCAP) QG[o] = (unsigned)gx | ((unsigned)yt << 12) | (m4 << 20);
                        else atomicOr(b.err, ORBX_DEVERR_CAND_OVERFLOW);      // unreachable for widths <= 4128
                    }
                }
            }
        }
        __syncthreads();
        // -- expand the group entries into pixel entries (dense queue for phase A) --
        {
            const int ng = min(sh.qg_count, QGCAP);
            for (int g0 = warp * 32; g0 < ng; g0 += NT) {
                const int gi = g0 + lane;
                const unsigned ge = gi < ng ? QG[gi] : 0u;
                const unsigned m4 = ge >> 20;
                const int n = __popc(m4);
                int inc = n;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
                const int total = __shfl_sync(0xffffffffu, inc, 31);
                int base = 0;
                if (lane == 0 && total) base = atoms_add(&sh.q_count, total);
                base = __shfl_sync(0xffffffffu, base, 0);
                int o = base + inc - n;
                const unsigned e = ((ge & 0xFFFu) << 2) | (((ge >> 12) & 0xFFu) << 16);
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (m4 & (1u << q)) {
                        if (o < QCAP) Q[o] = e + q;
                        else score_pixel(T, M, tp, roff, (e + q) & 0xFFFF, (int)((ge >> 12) & 0xFFu), minTh, &sh.cl_count);   // queue full: score inline
                        o++;
                    }
            }
        }
        __syncthreads();
        // -- phase A: signed pair test, dense and branch-free; survivors are compacted into Q2 --
        const int nq = min(sh.q_count, QCAP);
        for (int e0 = warp * 32; e0 < nq; e0 += NT) {
            const int e = e0 + lane;
            bool pass = false; unsigned ent = 0;
            if (e < nq) {
                ent = Q[e];
                unsigned pk[16];
                load_ring(T + (ent >> 16) * tp + (ent & 0xFFFF), roff, pk);
                pass = pair_test(pk, minTh);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, pass);
            if (bal) {
                int base = 0;
                if (lane == 0) base = atoms_add(&sh.q2_count, __popc(bal));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (pass) {
                    const int o = base + __popc(bal & ((1u << lane) - 1));
                    if (o < Q2CAP) Q2[o] = ent;
                    else score_pixel(T, M, tp, roff, ent & 0xFFFF, ent >> 16, minTh, &sh.cl_count);
                }
            }
        }
        __syncthreads();
        // -- phase B: exact arc measure for the survivors --
        const int nq2 = min(sh.q2_count, Q2CAP);
        for (int e = tid; e < nq2; e += NT) {
            const unsigned ent = Q2[e];
            const int x = ent & 0xFFFF, yq = ent >> 16;
            unsigned pk[16];
            load_ring(T + yq * tp + x, roff, pk);
            const int m = arc_measure(pk);
            if (m > minTh) {
                M[(yq - 3) * tp + x] = (uint8_t)m;
                const int o = atoms_add(&sh.cl_count, 1);
                if (o < CLCAP) CL[o] = (unsigned)x | ((unsigned)(yq - 3) << 16);
            }
        }
        __syncthreads();      // every thread has read the band's counters before thread 0 resets them
    }

    // ---- 4. in-cell non-max suppression ----
    // Corners were listed by phase B, so the suppression runs one corner per thread instead of diverging over a
    // sparse map; a corner-dense tile (list overflow) falls back to scanning the map, still exact.
    const int clcap = CLCAP;
    if (sh.cl_count <= clcap) {
        const int ncl = sh.cl_count;
        for (int e = tid; e < ncl; e += NT) {
            const unsigned ent = CL[e];
            const int x = ent & 0xFFFF, r = ent >> 16;
            const uint8_t* qm = M + r * tp + x;
            const int sc = qm[0];
            const int j = (x - xs0) / L.wCell;
            const int c0 = xs0 + j * L.wCell, c1 = min(c0 + L.wCell, xs1);     // cell interior [c0, c1)
            const bool hl = x - 1 >= c0, hr = x + 1 < c1, vu = r > 0, vd = r + 1 < hs;
            // neighbours outside the cell count as 0 (they belong to another FAST call in the reference)
            const int l0 = hl ? qm[-1] : 0, r0 = hr ? qm[1] : 0;
            const int u0 = vu ? qm[-tp] : 0, ul = (vu && hl) ? qm[-tp - 1] : 0, ur = (vu && hr) ? qm[-tp + 1] : 0;
            const int d0 = vd ? qm[tp] : 0, dl = (vd && hl) ? qm[tp - 1] : 0, dr = (vd && hr) ? qm[tp + 1] : 0;
            const int mx = max(max(max(l0, r0), max(u0, ul)), max(max(ur, d0), max(dl, dr)));
            if (sc > mx) {
                atomicOr(&Bmin[r * bw + (x >> 5)], 1u << (x & 31));
                if (sc > iniTh) atomicOr(&Bini[r * bw + (x >> 5)], 1u << (x & 31));
            }
        }
    } else {
        // corner-dense tile (more corners than the list holds): suppress in place over the map, still exact
        const int nw = tp >> 2;
        for (int r = warp; r < hs; r += NW)
            for (int wx = lane; wx < nw; wx += 32) {
                const unsigned word = reinterpret_cast<const unsigned*>(M + r * tp)[wx];
                if (!word) continue;
                for (int q = 0; q < 4; q++) {
                    const int sc = (word >> (8 * q)) & 0xFF;
                    if (!sc) continue;
                    const int x = (wx << 2) + q;
                    const uint8_t* qm = M + r * tp + x;
                    const int j = (x - xs0) / L.wCell;
                    const int c0 = xs0 + j * L.wCell, c1 = min(c0 + L.wCell, xs1);
                    const bool hl = x - 1 >= c0, hr = x + 1 < c1, vu = r > 0, vd = r + 1 < hs;
                    const int l0 = hl ? qm[-1] : 0, r0 = hr ? qm[1] : 0;
                    const int u0 = vu ? qm[-tp] : 0, ul = (vu && hl) ? qm[-tp - 1] : 0, ur = (vu && hr) ? qm[-tp + 1] : 0;
                    const int d0 = vd ? qm[tp] : 0, dl = (vd && hl) ? qm[tp - 1] : 0, dr = (vd && hr) ? qm[tp + 1] : 0;
                    const int mx = max(max(max(l0, r0), max(u0, ul)), max(max(ur, d0), max(dl, dr)));
                    if (sc > mx) {
                        atomicOr(&Bmin[r * bw + (x >> 5)], 1u << (x & 31));
                        if (sc > iniTh) atomicOr(&Bini[r * bw + (x >> 5)], 1u << (x & 31));
                    }
                }
            }
    }
    __syncthreads();

    // ---- 5. counts per (cell, row), threshold choice per cell, offsets ----
    for (int k = tid; k < L.nCols * hs; k += NT) {
        const int j = k / hs, r = k - j * hs;
        const int c0 = xs0 + j * L.wCell, c1 = min(c0 + L.wCell, xs1);
        int a = 0, bq = 0;
        if (c1 > c0) { a = popc_range(Bmin + r * bw, c0, c1); bq = popc_range(Bini + r * bw, c0, c1); }
        cnt_min[j * hmax + r] = (unsigned short)a;
        cnt_ini[j * hmax + r] = (unsigned short)bq;
    }
    __syncthreads();
    for (int j = tid; j < L.nCols; j += NT) {
        int ti = 0;
        for (int r = 0; r < hs; r++) ti += cnt_ini[j * hmax + r];
        const bool ui = ti > 0;
        int run = 0;
        for (int r = 0; r < hs; r++) {
            const int c = ui ? cnt_ini[j * hmax + r] : cnt_min[j * hmax + r];
            cnt_ini[j * hmax + r] = (unsigned short)run;        // exclusive prefix inside the cell
            run += c;
        }
        sh.use_ini[j] = ui;
        sh.cell_off[j + 1] = run;                               // cell totals, scanned below
    }
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        sh.cell_off[0] = 0;
        for (int j = 0; j < L.nCols; j++) { run += sh.cell_off[j + 1]; sh.cell_off[j + 1] = run; }   // cell_off[j] = first slot of cell j
        *row_count = min(run, L.row_cap);
        if (run > L.row_cap) atomicOr(b.err, ORBX_DEVERR_CAND_OVERFLOW);
    }
    __syncthreads();

    // ---- ordered emission: every survivor computes its own slot ----
    uint32_t* out = b.row_cand + (long long)f * b.row_cand_stride + b.row_off[blockIdx.x];
    const int yrel0 = iniY - ORBX_BORDER + 3;
    for (int k = tid; k < hs * bw; k += NT) {
        const int r = k / bw, wi = k - r * bw;
        unsigned bits = Bmin[r * bw + wi];
        const unsigned ibits = Bini[r * bw + wi];
        while (bits) {
            const int bpos = __ffs(bits) - 1;
            bits &= bits - 1;
            const int x = (wi << 5) + bpos;
            const int j = (x - xs0) / L.wCell;
            const bool ui = sh.use_ini[j];
            if (ui && !((ibits >> bpos) & 1u)) continue;
            const int c0 = xs0 + j * L.wCell;
            const int rank = x > c0 ? popc_range((ui ? Bini : Bmin) + r * bw, c0, x) : 0;
            const int o = sh.cell_off[j] + cnt_ini[j * hmax + r] + rank;
            if (o < L.row_cap)
                out[o] = (uint32_t)(x - ORBX_BORDER) | ((uint32_t)(yrel0 + r) << 12) | ((uint32_t)(M[r * tp + x] - 1) << 24);
        }
    }
}

size_t fast_smem_bytes(const OrbxGeom& g)
{
    size_t smem = 0;
    for (int l = 0; l < g.nlevels; l++) {
        const OrbxLevel& L = g.lv[l];
        if (L.nCols <= 0 || L.nRows <= 0) continue;
        const size_t tp = (L.w + 15) & ~15;
        const size_t bw = (L.w + 31) >> 5;
        const size_t need = (size_t)(L.hCell + 6) * tp + (size_t)L.hCell * tp + (QCAP + QGCAP + Q2CAP + CLCAP) * 4 + 2 * L.hCell * bw * 4 +
                            2 * (size_t)L.nCols * L.hCell * 2 + 16;
        if (need > smem) smem = need;
    }
    return smem;
}

}  // namespace

void orbx_fast_configure(const OrbxGeom& g)
{
    cudaFuncSetAttribute(k_fast_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast_smem_bytes(g));
}

void orbx_launch_fast(const OrbxGeom& g, const OrbxBuffers& b, const uint8_t* level0, int pitch0,
                      long long stride0, int batch, cudaStream_t s)
{
    if (g.total_rows == 0) return;
    dim3 grid(g.total_rows, batch);
    k_fast_rows<<<grid, NT, fast_smem_bytes(g), s>>>(g, b, level0, pitch0, stride0);
    ORBX_COUNT_LAUNCH(1);
}
