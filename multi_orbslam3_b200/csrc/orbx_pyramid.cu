// orbx_pyramid.cu - image pyramid + Gaussian blur, fused per level.
//
// Replaces ORBextractor::ComputePyramid (R/src/ORBextractor.cc:1152-1177: cv::resize INTER_LINEAR chain)
// and the per-level cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) of operator() (:1114-1115).
// One launch per level l: a CTA stages the source tile of level l-1 in shared memory, produces the
// 64x32 tile of level l plus a 3-px halo in shared memory (bit-exact OpenCV fixed-point bilinear),
// stores the tile, then runs the separable integer blur [18,34,48,56,48,34,18]/256 out of shared
// memory and stores the blurred tile.  Level l-1 is read once, level l and blur(l) are written once.
// The 19-px reflected border the reference materialises is never read downstream and is not built.
#include "orbx_internal.h"

namespace {

constexpr int TW = 64, TH = 32;          // dst tile
constexpr int RW = TW + 6, RH = TH + 6;  // tile + blur halo
constexpr int RP = 72;                   // smem pitch of the resized tile
constexpr int NT = 256;

struct PyrArgs {
    const uint8_t* src; int spitch; long long sstride; int sw, sh;   // level l-1 (or the frame for l == 0)
    uint8_t* dst; int dpitch; long long dstride; int w, h;           // level l (unused for l == 0)
    uint8_t* blur; int bpitch; long long bstride;                    // blur(l)
    const short4* xt; const short4* yt;                              // resize tables of level l
    int src_tw, src_th;                                              // smem source tile extents (max over tiles)
};

template <bool RESIZE>
__global__ void __launch_bounds__(NT) k_pyr_level(PyrArgs a)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* R = smem;                                        // [RH][RP]
    uint16_t* Hb = reinterpret_cast<uint16_t*>(smem + RH * RP);   // [RH][TW]
    uint8_t* S = smem + RH * RP + RH * TW * 2;                // [src_th][src_tp]
    const int tid = threadIdx.x;
    const int X0 = blockIdx.x * TW, Y0 = blockIdx.y * TH;
    const int f = blockIdx.z;
    const int w = a.w, h = a.h;
    const int ax = max(X0 - 3, 0), bx = min(X0 + TW + 3, w);
    const int ay = max(Y0 - 3, 0), by = min(Y0 + TH + 3, h);
    const int rw = bx - ax, rh = by - ay;
    const int tw = min(TW, w - X0), th = min(TH, h - Y0);

    if (RESIZE) {
        const uint8_t* src = a.src + (long long)f * a.sstride;
        const short4 xa = a.xt[ax], xb = a.xt[bx - 1];
        const short4 ya = a.yt[ay], yb = a.yt[by - 1];
        const int sx0 = xa.x & ~3, sx1 = xb.y;           // first column aligned down to 4
        const int sy0 = ya.x, sy1 = yb.y;
        const int sp = a.src_tw;                          // smem source pitch (multiple of 4)
        const int nwords = (sx1 - sx0 + 4) >> 2;
        const int nrows = sy1 - sy0 + 1;
        const bool aligned = ((a.spitch & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 3) == 0);
        if (aligned) {
            for (int i = tid; i < nrows * nwords; i += NT) {
                int r = i / nwords, c = i - r * nwords;
                // pitch is padded to a multiple of 64 for our own levels; for the caller's frame the
                // last word may run past the row end but stays inside the allocation's pitch
                const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(src + (long long)(sy0 + r) * a.spitch + sx0) + c);
                *reinterpret_cast<uint32_t*>(S + r * sp + c * 4) = v;
            }
        } else {
            const int ncols = sx1 - sx0 + 1;
            for (int i = tid; i < nrows * ncols; i += NT) {
                int r = i / ncols, c = i - r * ncols;
                S[r * sp + c] = __ldg(src + (long long)(sy0 + r) * a.spitch + sx0 + c);
            }
        }
        __syncthreads();
        for (int i = tid; i < rh * rw; i += NT) {
            int ry = i / rw, rx = i - ry * rw;
            const short4 xe = a.xt[ax + rx];
            const short4 ye = a.yt[ay + ry];
            const uint8_t* r0 = S + (ye.x - sy0) * sp - sx0;
            const uint8_t* r1 = S + (ye.y - sy0) * sp - sx0;
            const int H0 = r0[xe.x] * xe.z + r0[xe.y] * xe.w;
            const int H1 = r1[xe.x] * xe.z + r1[xe.y] * xe.w;
            const int v = (((ye.z * (H0 >> 4)) >> 16) + ((ye.w * (H1 >> 4)) >> 16) + 2) >> 2;
            R[ry * RP + rx] = (uint8_t)min(max(v, 0), 255);
        }
    } else {
        const uint8_t* src = a.src + (long long)f * a.sstride;
        for (int i = tid; i < rh * rw; i += NT) {
            int ry = i / rw, rx = i - ry * rw;
            R[ry * RP + rx] = __ldg(src + (long long)(ay + ry) * a.spitch + ax + rx);
        }
    }
    __syncthreads();

    if (RESIZE) {
        // store the level tile: 16 threads x 4 bytes per row
        uint8_t* dst = a.dst + (long long)f * a.dstride;
        const int ox = X0 - ax, oy = Y0 - ay;
        for (int i = tid; i < th * (TW / 4); i += NT) {
            int r = i / (TW / 4), c = (i - r * (TW / 4)) * 4;
            if (c < tw) {
                const uint8_t* p = R + (oy + r) * RP + ox + c;
                uint32_t v = p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24);
                *reinterpret_cast<uint32_t*>(dst + (long long)(Y0 + r) * a.dpitch + X0 + c) = v;   // pitch padding absorbs the tail
            }
        }
    }
    // horizontal blur pass -> Hb (fits 16 bit: 256*255)
    for (int i = tid; i < rh * TW; i += NT) {
        int ry = i / TW, j = i - ry * TW;
        if (j < tw) {
            const uint8_t* row = R + ry * RP - ax;
            const int x = X0 + j;
            int acc;
            if (x >= 3 && x + 3 < w) {
                acc = 18 * (row[x - 3] + row[x + 3]) + 34 * (row[x - 2] + row[x + 2]) +
                      48 * (row[x - 1] + row[x + 1]) + 56 * row[x];
            } else {
                acc = 18 * (row[orbx_reflect101(x - 3, w)] + row[orbx_reflect101(x + 3, w)]) +
                      34 * (row[orbx_reflect101(x - 2, w)] + row[orbx_reflect101(x + 2, w)]) +
                      48 * (row[orbx_reflect101(x - 1, w)] + row[orbx_reflect101(x + 1, w)]) + 56 * row[x];
            }
            Hb[ry * TW + j] = (uint16_t)acc;
        }
    }
    __syncthreads();
    // vertical pass, 4 pixels per thread
    uint8_t* bl = a.blur + (long long)f * a.bstride;
    for (int i = tid; i < th * (TW / 4); i += NT) {
        int r = i / (TW / 4), c = (i - r * (TW / 4)) * 4;
        if (c < tw) {
            const int y = Y0 + r;
            int yy[7];
#pragma unroll
            for (int k = 0; k < 7; k++) yy[k] = (orbx_reflect101(y + k - 3, h) - ay) * TW + c;
            uint32_t out = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                uint32_t acc = 18u * (Hb[yy[0] + q] + Hb[yy[6] + q]) + 34u * (Hb[yy[1] + q] + Hb[yy[5] + q]) +
                               48u * (Hb[yy[2] + q] + Hb[yy[4] + q]) + 56u * Hb[yy[3] + q];
                out |= ((acc + 32768u) >> 16) << (8 * q);
            }
            *reinterpret_cast<uint32_t*>(bl + (long long)y * a.bpitch + X0 + c) = out;
        }
    }
}

}  // namespace

void orbx_launch_pyramid(const OrbxGeom& g, const OrbxBuffers& b, const uint8_t* level0, int pitch0,
                         long long stride0, int batch, cudaStream_t s)
{
    for (int l = 0; l < g.nlevels; l++) {
        const OrbxLevel& L = g.lv[l];
        PyrArgs a{};
        a.w = L.w; a.h = L.h;
        a.blur = b.blur[l]; a.bpitch = L.pitch; a.bstride = L.frame_stride;
        dim3 grid((L.w + TW - 1) / TW, (L.h + TH - 1) / TH, batch);
        size_t smem = RH * RP + RH * TW * 2;
        if (l == 0) {
            a.src = level0; a.spitch = pitch0; a.sstride = stride0; a.sw = L.w; a.sh = L.h;
            k_pyr_level<false><<<grid, NT, smem, s>>>(a);
            ORBX_COUNT_LAUNCH(1);
        } else {
            const OrbxLevel& P = g.lv[l - 1];
            a.src = (l == 1) ? level0 : b.pyr[l - 1];
            a.spitch = (l == 1) ? pitch0 : P.pitch;
            a.sstride = (l == 1) ? stride0 : P.frame_stride;
            a.sw = P.w; a.sh = P.h;
            a.dst = b.pyr[l]; a.dpitch = L.pitch; a.dstride = L.frame_stride;
            a.xt = b.tabs + L.xtab_off; a.yt = b.tabs + L.ytab_off;
            // source tile extents for a (TW+6) x (TH+6) destination tile; +8 slack for 4-byte alignment
            const double sx = (double)P.w / L.w, sy = (double)P.h / L.h;
            a.src_tw = (((int)(RW * sx) + 12) + 3) & ~3;
            a.src_th = (int)(RH * sy) + 4;
            smem += (size_t)a.src_tw * a.src_th;
            k_pyr_level<true><<<grid, NT, smem, s>>>(a);
            ORBX_COUNT_LAUNCH(1);
        }
    }
}
