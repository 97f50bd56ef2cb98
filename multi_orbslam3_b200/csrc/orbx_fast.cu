// orbx_fast.cu - grid FAST-9/16 with the iniThFAST / minThFAST fallback.
//
// Replaces the cell loop of ORBextractor::ComputeKeyPointsOctTree (R/src/ORBextractor.cc:787-854) and the
// cv::FAST(cell, th, nonmax=true) calls inside it (:808, :827).
//
// One CTA owns one row of cells of one level of one frame.  It stages rows [iniY, maxY) in shared memory
// with 16-byte loads, computes the arc measure m (max over the 16 arcs of 9 ring pixels of
// max(min d, min -d); corner <=> m > t, cv::FAST response = m - 1, independent of t) once per pixel,
// and derives BOTH thresholds from it: because the reference's non-max suppression compares a corner
// only with scores inside the same FAST call (= the same cell) and non-corners score 0, a pixel survives
// at threshold t  <=>  m > t and m is a strict maximum among its in-cell neighbours' m.  Cells with no
// survivor at iniThFAST fall back to the minThFAST survivors (:825-828).  Survivors are emitted in the
// reference's order (cell by cell, row-major inside a cell) by warp-ballot compaction.  No score map is
// ever written to global memory.
#include "orbx_internal.h"

namespace {

constexpr int NT = 256;
constexpr int MAX_CELLS = 128;

// ring offsets (dx,dy), OpenCV order
__device__ __constant__ int8_t c_ring[16][2] = {
    {0, 3}, {1, 3}, {2, 2}, {3, 1}, {3, 0}, {3, -1}, {2, -2}, {1, -3},
    {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

// Arc measure on packed 16-bit lanes.  p[k] holds both polarities of ring pixel k, biased by 256 so that
// both lanes stay positive: low = v - r_k + 256 (= d_k + 256), high = r_k - v + 256 (= -d_k + 256).
// One unsigned 16x2 min network (windows of 9 by doubling: 2, 4, 8, +1) then serves "9 darker" and
// "9 brighter" at once: m = max over arcs of max(min d, min -d).
// NOTE: the scalar formulation max(min(...), -max(...)) is MIScompiled by ptxas 12.9 / the 580 driver JIT for
// sm_100 (verified on B200: -Xptxas -O0 and -G give the right answer, -O1..-O3 return max(d)); see DESIGN.md.
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

__global__ void __launch_bounds__(NT) k_fast_rows(OrbxGeom g, OrbxBuffers b, const uint8_t* level0, int pitch0,
                                                  long long stride0)
{
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ int s_cnt_ini[MAX_CELLS], s_cnt_min[MAX_CELLS], s_off[MAX_CELLS + 1];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int f = blockIdx.y;
    // locate (level, cell row)
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
    const int nrow = maxY - iniY;                  // staged rows
    const int ncol = L.maxBX - ORBX_BORDER;        // staged columns (x_rel = abs x - 16)
    const int tp = (ncol + 15) & ~15;              // smem pitch
    const int hs = nrow - 6, ws = ncol - 6;        // scored region (rel rows 3.., rel cols 3..)
    if (hs <= 0 || ws <= 0) { if (tid == 0) *row_count = 0; return; }
    uint8_t* T = smem;                             // [nrow][tp] pixels, later survivor flags
    uint8_t* M = smem + (size_t)(L.hCell + 6) * tp;   // [hs][tp] arc measure (0 when <= minTh)

    const uint8_t* img; int pitch;
    if (l == 0) { img = level0 + (long long)f * stride0; pitch = pitch0; }
    else { img = b.pyr[l] + (long long)f * L.frame_stride; pitch = L.pitch; }

    // ---- stage rows, 16 bytes per load when the layout allows ----
    const bool vec = ((pitch & 15) == 0) && ((reinterpret_cast<uintptr_t>(img) & 15) == 0);
    if (vec) {
        const int nv = tp >> 4;    // may read up to 15 bytes of row padding: pitch is a multiple of 16 >= w
        for (int k = tid; k < nrow * nv; k += NT) {
            int r = k / nv, c = k - r * nv;
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(img + (long long)(iniY + r) * pitch + ORBX_BORDER) + c);
            *reinterpret_cast<uint4*>(T + r * tp + c * 16) = v;
        }
    } else {
        for (int k = tid; k < nrow * ncol; k += NT) {
            int r = k / ncol, c = k - r * ncol;
            T[r * tp + c] = __ldg(img + (long long)(iniY + r) * pitch + ORBX_BORDER + c);
        }
    }
    for (int k = tid; k < L.nCols; k += NT) { s_cnt_ini[k] = 0; s_cnt_min[k] = 0; }
    __syncthreads();

    // ---- arc measure per scored pixel ----
    const int minTh = g.min_th, iniTh = g.ini_th;
    int roff[16];
#pragma unroll
    for (int k = 0; k < 16; k++) roff[k] = c_ring[k][1] * tp + c_ring[k][0];
    for (int k = tid; k < hs * ws; k += NT) {
        const int r = k / ws, c = k - r * ws;
        const uint8_t* p = T + (r + 3) * tp + c + 3;
        const int v = p[0];
        int m = 0;
        // high-speed pre-test at minTh on the opposite pairs (k, k+8)
        const int lo = v - minTh, hi = v + minTh;
#define CLS(q) ((p[roff[q]] < lo ? 1 : 0) | (p[roff[q]] > hi ? 2 : 0))
        int t = CLS(0) | CLS(8);
        if (t) {
            t &= CLS(4) | CLS(12);
            if (t) {
                t &= CLS(2) | CLS(10); t &= CLS(6) | CLS(14);
                if (t) {
                    t &= CLS(1) | CLS(9); t &= CLS(3) | CLS(11); t &= CLS(5) | CLS(13); t &= CLS(7) | CLS(15);
                    if (t) {
                        unsigned pk[16];
                        const unsigned cv = (unsigned)(v + 256) | ((unsigned)(256 - v) << 16);
#pragma unroll
                        for (int q = 0; q < 16; q++) pk[q] = cv + (unsigned)p[roff[q]] * 0xFFFFu;   // (v+256-r) | (256-v+r)<<16
                        m = arc_measure(pk);
                        if (m <= minTh) m = 0;
                    }
                }
            }
        }
#undef CLS
        M[r * tp + c] = (uint8_t)m;
    }
    __syncthreads();

    // ---- in-cell non-max suppression; flags overwrite the pixel tile: bit0 survivor(minTh), bit1 survivor(iniTh) ----
    for (int k = tid; k < hs * ws; k += NT) {
        const int r = k / ws, c = k - r * ws;
        const int s = M[r * tp + c];
        uint8_t flag = 0;
        if (s > 0) {
            const int j = c / L.wCell;
            const int c0 = j * L.wCell, c1 = min(c0 + L.wCell, ws);   // cell interior [c0, c1)
            const bool hl = c - 1 >= c0, hr = c + 1 < c1, vu = r > 0, vd = r + 1 < hs;
            const uint8_t* q = M + r * tp + c;
            bool keep = true;
            if (hl) keep &= s > q[-1];
            if (hr) keep &= s > q[1];
            if (vu) { keep &= s > q[-tp]; if (hl) keep &= s > q[-tp - 1]; if (hr) keep &= s > q[-tp + 1]; }
            if (vd) { keep &= s > q[tp]; if (hl) keep &= s > q[tp - 1]; if (hr) keep &= s > q[tp + 1]; }
            if (keep) {
                flag = 1;
                atomicAdd(&s_cnt_min[j], 1);
                if (s > iniTh) { flag = 3; atomicAdd(&s_cnt_ini[j], 1); }
            }
        }
        T[r * tp + c] = flag;
    }
    __syncthreads();

    // ---- per-cell threshold choice and offsets (cells are skipped like the reference does, :801) ----
    if (warp == 0) {
        int run = 0;
        for (int base = 0; base < L.nCols; base += 32) {
            const int j = base + lane;
            int cnt = 0;
            if (j < L.nCols) {
                const int iniX = ORBX_BORDER + j * L.wCell;
                if (iniX < L.maxBX - 6) cnt = s_cnt_ini[j] > 0 ? s_cnt_ini[j] : s_cnt_min[j];
            }
            int inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            if (j < L.nCols) s_off[j] = run + inc - cnt;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) {
            s_off[L.nCols] = run;
            *row_count = min(run, L.row_cap);
            if (run > L.row_cap) atomicOr(b.err, ORBX_DEVERR_CAND_OVERFLOW);
        }
    }
    __syncthreads();

    // ---- ordered emission: one warp per cell, row-major inside the cell ----
    uint32_t* out = b.row_cand + (long long)f * b.row_cand_stride + b.row_off[blockIdx.x];
    const int yrel0 = iniY - ORBX_BORDER + 3;
    for (int j = warp; j < L.nCols; j += NT / 32) {
        const int total = s_off[j + 1] - s_off[j];
        if (total == 0) continue;
        const uint8_t need = s_cnt_ini[j] > 0 ? 2 : 1;
        const int c0 = j * L.wCell, c1 = min(c0 + L.wCell, ws);
        int pos = s_off[j];
        for (int r = 0; r < hs; r++) {
            for (int cb = c0; cb < c1; cb += 32) {
                const int c = cb + lane;
                const bool on = c < c1 && (T[r * tp + c] & need);
                const unsigned bal = __ballot_sync(0xffffffffu, on);
                if (on) {
                    const int o = pos + __popc(bal & ((1u << lane) - 1));
                    if (o < L.row_cap)
                        out[o] = (uint32_t)(c + 3) | ((uint32_t)(yrel0 + r) << 12) | ((uint32_t)(M[r * tp + c] - 1) << 24);
                }
                pos += __popc(bal);
            }
        }
    }
}

}  // namespace

static size_t fast_smem_bytes(const OrbxGeom& g)
{
    size_t smem = 0;
    for (int l = 0; l < g.nlevels; l++) {
        const OrbxLevel& L = g.lv[l];
        if (L.nCols <= 0 || L.nRows <= 0) continue;
        const size_t tp = ((L.maxBX - ORBX_BORDER) + 15) & ~15;
        const size_t need = (size_t)(L.hCell + 6) * tp + (size_t)L.hCell * tp;
        if (need > smem) smem = need;
    }
    return smem;
}

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
