// orbx_pyramid.cu - image pyramid + Gaussian blur, fused per level.
//
// Replaces ORBextractor::ComputePyramid (R/src/ORBextractor.cc:1152-1177: cv::resize INTER_LINEAR chain)
// and the per-level cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) of operator() (:1114-1115).
// One launch per level l: a CTA stages the source tile of level l-1 in shared memory, produces the
// 64x32 tile of level l plus a 3-px halo in shared memory (bit-exact OpenCV fixed-point bilinear),
// stores the tile, then runs the separable integer blur [18,34,48,56,48,34,18]/256 out of shared
// memory and stores the blurred tile.  Level l-1 is read once, level l and blur(l) are written once.
// The 19-px reflected border the reference materialises is never read downstream and is not built.
#include <cuda.h>
#include <string.h>
#include "orbx_internal.h"

namespace generic {

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

}  // namespace generic

static void launch_pyramid_generic(const OrbxGeom& g, const OrbxBuffers& b, const uint8_t* level0, int pitch0,
                         long long stride0, int batch, cudaStream_t s, int l_begin, int l_end, int parts)
{
    using namespace generic;
    for (int l = l_begin; l < l_end; l++) {
        // (no split here: the resize part of a level runs the fused kernel, its blur part is then already done)
        if ((parts == ORBX_PYR_BLUR && l >= 1) || (parts == ORBX_PYR_RESIZE && l == 0)) continue;
        const OrbxLevel& L = g.lv[l];
        PyrArgs a{};
        a.w = L.w; a.h = L.h;
        a.blur = b.blur[l]; a.bpitch = L.pitch; a.bstride = L.frame_stride;
        dim3 grid((L.w + TW - 1) / TW, (L.h + TH - 1) / TH, batch);
        size_t smem = RH * RP + RH * TW * 2;
        if (l == 0) {
            a.src = level0; a.spitch = pitch0; a.sstride = stride0; a.sw = L.w; a.sh = L.h;
            ORBX_OPTIN_SMEM(k_pyr_level<false>);
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
            ORBX_OPTIN_SMEM(k_pyr_level<true>);
            k_pyr_level<true><<<grid, NT, smem, s>>>(a);
            ORBX_COUNT_LAUNCH(1);
        }
    }
}

// =====================================================================================================
// Fast path (scale factors up to 1.5): 128x64 destination tiles, 256 threads.
//   S  source tile of level l-1 (TMA 3-D tiled bulk load when the layout allows, else cooperative 16-byte loads)
//   Hs horizontally interpolated source rows, (S[x0]*c0 + S[x1]*c1) >> 4, 16 bit
//   R  resized tile + 3-px halo (halo outside the image is filled by BORDER_REFLECT_101 of the level itself)
//   Hb horizontal blur pass (exact integers <= 65280, 16 bit); blurred bytes are staged back into R
// The blur is exact in any pass order because only the final (sum + 32768) >> 16 rounds.
// =====================================================================================================
namespace fastp {

constexpr int TW = 128, TH = 64, NT = 256, NWARP = NT / 32;
constexpr int RW = TW + 6, RH = TH + 6;
constexpr int RO = 16;                 // column of R that holds dst x = X0 (16-byte aligned interior)
constexpr int RP = 160;                // R pitch
constexpr int HP = 144;                // Hs pitch (u16 elements): 36 groups of 4 columns

struct Args {
    const uint8_t* src; int spitch; long long sstride; int sw, sh;
    uint8_t* dst; int dpitch; long long dstride; int w, h;
    uint8_t* blur; int bpitch; long long bstride;
    const short4* xt; const short4* yt;
    int sp;        // smem source pitch (multiple of 16) = TMA box width
    int sh_max;    // smem source rows = TMA box height
    int use_tma;   // source tile arrives by cp.async.bulk.tensor (3-D tiled map: x, y, frame)
    int no_blur;   // resize and store the level only (the few-frame path runs the blur of the level on a branch stream)
};

// ---- TMA / mbarrier primitives (sm_90+ PTX; SASS: UTMALDG, SYNCS) ----
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
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int x, int y, int z, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(map), "r"(x), "r"(y), "r"(z),
                   "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

__device__ __forceinline__ void blur_and_store(uint8_t* R, uint16_t* Hb, uint8_t* blur_base, int bpitch, int X0, int Y0,
                                               int tw, int th)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // ---- horizontal pass: lane = 4-pixel group, warp = PAIR of rows; two dp4a per pixel.  The 16-bit results of rows
    //      (2k, 2k+1) are stored interleaved, one 32-bit word per column, so that the vertical pass can feed them to dp2a ----
    const unsigned K0 = 18u | (34u << 8) | (48u << 16) | (56u << 24), K1 = 48u | (34u << 8) | (18u << 16);
    unsigned* Hp = reinterpret_cast<unsigned*>(Hb);               // [(TH+6)/2][TW] words: H(2k, x) | H(2k+1, x) << 16
    for (int pr = warp; pr < (th + 7) / 2; pr += NWARP) {
        unsigned v[2][4];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const unsigned* rr = reinterpret_cast<const unsigned*>(R + (2 * pr + h) * RP) + (RO >> 2) + lane;
            const unsigned wm = rr[-1], w0 = rr[0], w1 = rr[1];
            v[h][0] = __dp4a(__funnelshift_r(w0, w1, 8), K1, __dp4a(__funnelshift_r(wm, w0, 8), K0, 0u));
            v[h][1] = __dp4a(__funnelshift_r(w0, w1, 16), K1, __dp4a(__funnelshift_r(wm, w0, 16), K0, 0u));
            v[h][2] = __dp4a(__funnelshift_r(w0, w1, 24), K1, __dp4a(__funnelshift_r(wm, w0, 24), K0, 0u));
            v[h][3] = __dp4a(w1, K1, __dp4a(w0, K0, 0u));
        }
        *reinterpret_cast<uint4*>(Hp + pr * TW + lane * 4) =
            make_uint4(v[0][0] | (v[1][0] << 16), v[0][1] | (v[1][1] << 16), v[0][2] | (v[1][2] << 16), v[0][3] | (v[1][3] << 16));
    }
    __syncthreads();
    // ---- vertical pass: 4 columns x 8 rows per thread, 4 dp2a per pixel (two taps each); blurred bytes go back into R ----
    {
        const int cg = tid & 31, r0 = (tid >> 5) * 8;             // NT == 32 * TH / 8
        if (r0 < th) {
            uint4 P[7];
#pragma unroll
            for (int k = 0; k < 7; k++) P[k] = *reinterpret_cast<const uint4*>(Hp + (r0 / 2 + k) * TW + cg * 4);
            // taps (18,34,48,56,48,34,18) over rows q..q+6: even q starts on a pair, odd q on the high half of one
            const unsigned WE0 = 18u | (34u << 8), WE1 = 48u | (56u << 8), WE2 = 48u | (34u << 8), WE3 = 18u;
            const unsigned WO0 = 18u << 8, WO1 = 34u | (48u << 8), WO2 = 56u | (48u << 8), WO3 = 34u | (18u << 8);
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int m = q >> 1;
                const unsigned w0 = (q & 1) ? WO0 : WE0, w1 = (q & 1) ? WO1 : WE1, w2 = (q & 1) ? WO2 : WE2, w3 = (q & 1) ? WO3 : WE3;
                const unsigned ax = __dp2a_lo(P[m + 3].x, w3, __dp2a_lo(P[m + 2].x, w2, __dp2a_lo(P[m + 1].x, w1, __dp2a_lo(P[m].x, w0, 32768u))));
                const unsigned ay = __dp2a_lo(P[m + 3].y, w3, __dp2a_lo(P[m + 2].y, w2, __dp2a_lo(P[m + 1].y, w1, __dp2a_lo(P[m].y, w0, 32768u))));
                const unsigned az = __dp2a_lo(P[m + 3].z, w3, __dp2a_lo(P[m + 2].z, w2, __dp2a_lo(P[m + 1].z, w1, __dp2a_lo(P[m].z, w0, 32768u))));
                const unsigned aw = __dp2a_lo(P[m + 3].w, w3, __dp2a_lo(P[m + 2].w, w2, __dp2a_lo(P[m + 1].w, w1, __dp2a_lo(P[m].w, w0, 32768u))));
                // byte 2 of each accumulator = (sum + 2^15) >> 16
                const unsigned o = __byte_perm(__byte_perm(ax, ay, 0x0062), __byte_perm(az, aw, 0x0062), 0x5410);
                *reinterpret_cast<unsigned*>(R + (3 + r0 + q) * RP + RO + cg * 4) = o;
            }
        }
    }
    __syncthreads();
    // ---- 16-byte stores of the blurred tile ----
    for (int it = tid; it < th * (TW / 16); it += NT) {
        const int r = it >> 3, c16 = it & 7;
        if (c16 * 16 < tw)
            *reinterpret_cast<uint4*>(blur_base + (long long)(Y0 + r) * bpitch + X0 + c16 * 16) =
                *reinterpret_cast<const uint4*>(R + (3 + r) * RP + RO + c16 * 16);
    }
}

// BORDER_REFLECT_101 fill of the halo that lies outside the level: columns first, then whole rows.
// we / he = tile-local index of the first column / row outside the image (>= TW+3 / TH+3 when none is).
__device__ __forceinline__ void reflect_halo(uint8_t* R, int X0, int Y0, int w, int h)
{
    const int tid = threadIdx.x;
    const int we = w - X0, he = h - Y0;
    const bool left = X0 == 0, right = we < TW + 3, top = Y0 == 0, bottom = he < TH + 3;
    if (left || right) {
        const int rlo = top ? 3 : 0, rhi = bottom ? 3 + he : RH;      // R rows that hold image pixels
        for (int it = tid; it < (rhi - rlo) * 3; it += NT) {
            const int row = rlo + it / 3, d = it % 3;
            uint8_t* rr = R + row * RP + RO;
            if (left) rr[-(d + 1)] = rr[d + 1];
            if (right && we + d < TW + 3) rr[we + d] = rr[we - 2 - d];
        }
    }
    __syncthreads();
    if (top || bottom) {
        for (int it = tid; it < 3 * (RP / 4); it += NT) {
            const int d = it / (RP / 4), c4 = it % (RP / 4);
            unsigned* base = reinterpret_cast<unsigned*>(R) + c4;
            if (top) base[(3 - (d + 1)) * (RP / 4)] = base[(3 + d + 1) * (RP / 4)];
            if (bottom && he + d < TH + 3) base[(3 + he + d) * (RP / 4)] = base[(3 + he - 2 - d) * (RP / 4)];
        }
    }
    __syncthreads();
}

template <bool RESIZE>
__global__ void __launch_bounds__(NT) k_pyr_fast(Args a, const __grid_constant__ CUtensorMap tmap)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) unsigned long long s_bar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int X0 = blockIdx.x * TW, Y0 = blockIdx.y * TH;
    const int f = blockIdx.z;
    const int w = a.w, h = a.h;
    const int tw = min(TW, w - X0), th = min(TH, h - Y0);
    const int ax = max(X0 - 3, 0), bx = min(X0 + TW + 3, w);
    const int ay = max(Y0 - 3, 0), by = min(Y0 + TH + 3, h);
    const int rw = bx - ax, rh = by - ay;
    const int rcol0 = RO - (X0 - ax), rrow0 = 3 - (Y0 - ay);      // R position of dst (ax, ay)

    // smem: S (source tile, first: TMA wants a 128-byte aligned destination) | R | Hb/Hs | sy
    const int s_bytes = RESIZE ? ((a.sh_max * a.sp + 127) & ~127) : 0;
    uint8_t* R = smem + s_bytes;                                   // [RH][RP]
    uint16_t* Hb = reinterpret_cast<uint16_t*>(R + RH * RP);       // [RH][TW]  (aliases Hs)
    const uint8_t* srcf = a.src + (long long)f * a.sstride;
    orbx_pdl_prologue();

    if (RESIZE) {
        uint16_t* Hs = Hb;                                         // [sh_max][HP]
        const int hs_bytes = max(a.sh_max * HP * 2, RH * TW * 2);
        int4* sy = reinterpret_cast<int4*>(reinterpret_cast<uint8_t*>(Hb) + hs_bytes);   // [RH] per-row (y0, y1, b0, b1)
        uint8_t* S = smem;                                                               // [sh_max][sp]
        const short4 xa = a.xt[ax], xb = a.xt[bx - 1];
        const short4 ya = a.yt[ay], yb = a.yt[by - 1];
        const int sx0 = xa.x & ~15, sx1 = xb.y;   // first source column aligned down to 16 bytes (TMA moves 16-byte granules)
        const int sy0 = ya.x, sy1 = yb.y;
        const int sp = a.sp;
        const int nrows = sy1 - sy0 + 1;
        // ---- source tile ----
        const bool vec = ((a.spitch & 15) == 0) && ((reinterpret_cast<uintptr_t>(srcf) & 15) == 0);
        if (a.use_tma) {
            // one elected thread arms the mbarrier with the box size and issues the 3-D tiled bulk load;
            // out-of-image parts of the box are zero-filled by the TMA unit and never indexed (tables clamp)
            if (tid == 0) mbar_init(&s_bar, 1);
            __syncthreads();
            if (tid == 0) {
                mbar_expect_tx(&s_bar, (unsigned)(a.sp * a.sh_max));
                tma_load_3d(S, &tmap, sx0, sy0, f, &s_bar);
            }
        } else if (vec) {
            const int nv = (sx1 - sx0 + 16) >> 4;                 // stays inside the row pitch (pitch is a multiple of 16 >= sw)
            for (int r = warp; r < nrows; r += NWARP) {
                const uint4* g = reinterpret_cast<const uint4*>(srcf + (long long)(sy0 + r) * a.spitch + sx0);
                uint4* d = reinterpret_cast<uint4*>(S + r * sp);
                for (int c = lane; c < nv; c += 32) d[c] = __ldg(g + c);
            }
        } else {
            const int ncols = sx1 - sx0 + 1;
            for (int r = warp; r < nrows; r += NWARP)
                for (int c = lane; c < ncols; c += 32) S[r * sp + c] = __ldg(srcf + (long long)(sy0 + r) * a.spitch + sx0 + c);
        }
        if (tid < rh) {
            const short4 ye = a.yt[ay + tid];
            sy[tid] = make_int4(ye.x - sy0, ye.y - sy0, ye.z, ye.w);
        }
        // ---- per-thread constants: one thread owns 4 consecutive R columns (one aligned word of R) ----
        // R column rho <-> dst x = ax + (rho - rcol0), clamped to the tile's valid range (clamped duplicates land in
        // cells that are either unused or rewritten by reflect_halo).
        // The ng (<= 36) column groups of THIS tile times RL = NT / ng row lanes: a narrow tile at the right image edge keeps
        // (almost) every thread busy instead of idling the lanes of its missing columns.
        const int g0 = rcol0 >> 2;
        const int ng = ((rcol0 + rw - 1) >> 2) - g0 + 1;
        const int RL = NT / ng;
        const int rl = tid / ng, gq = tid - rl * ng;
        const bool worker = rl < RL;
        unsigned coef[4], psel[4], wb = 0, wsh = 0;
        if (worker) {
            int o0[4], o1[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int c = min(max(4 * (g0 + gq) + i - rcol0, 0), rw - 1);
                const short4 xe = a.xt[ax + c];
                o0[i] = xe.x - sx0; o1[i] = xe.y - sx0;
                coef[i] = (unsigned)xe.z | ((unsigned)xe.w << 16);
            }
            // the 4 columns read source bytes [o0[0], o0[0] + 8) (scale <= 1.55: o1[3] - o0[0] <= 6): two funnel shifts bring that
            // window into a register pair, and one loop-invariant PRMT selector per column picks its (S0, S1)
            wb = (unsigned)(o0[0] >> 2); wsh = 8u * (unsigned)(o0[0] & 3);
#pragma unroll
            for (int i = 0; i < 4; i++) psel[i] = (unsigned)(o0[i] - o0[0]) | ((unsigned)(o1[i] - o0[0]) << 4);
        }
        if (a.use_tma) mbar_wait(&s_bar, 0);       // table loads above overlap the bulk copy
        __syncthreads();
        // ---- horizontal interpolation of every source row: 3 word loads, 2 funnel shifts, 4 x (PRMT + DP2A) per 4 columns ----
        if (worker) {
            for (int r = rl; r < nrows; r += RL) {
                const unsigned* sr = reinterpret_cast<const unsigned*>(S + r * sp) + wb;
                const unsigned A = sr[0], B = sr[1], C = sr[2];
                const unsigned W0 = __funnelshift_r(A, B, wsh), W1 = __funnelshift_r(B, C, wsh);
                unsigned hv[4];
#pragma unroll
                for (int i = 0; i < 4; i++) hv[i] = __dp2a_lo(coef[i], __byte_perm(W0, W1, psel[i]), 0u) >> 4;      // (S0*c0 + S1*c1) >> 4
                *reinterpret_cast<uint2*>(Hs + r * HP + 4 * gq) = make_uint2(hv[0] | (hv[1] << 16), hv[2] | (hv[3] << 16));
            }
        }
        __syncthreads();
        // ---- vertical interpolation -> R: ((b0*h0)>>16 + (b1*h1)>>16 + 2) >> 2, one aligned word of R per step ----
        if (worker) {
            for (int ry = rl; ry < rh; ry += RL) {
                const int4 ye = sy[ry];
                const uint2 h0 = *reinterpret_cast<const uint2*>(Hs + ye.x * HP + 4 * gq);
                const uint2 h1 = *reinterpret_cast<const uint2*>(Hs + ye.y * HP + 4 * gq);
                const unsigned b0 = (unsigned)ye.z << 16, b1 = (unsigned)ye.w << 16;
                const unsigned v0 = (__umulhi(b0, h0.x & 0xFFFFu) + __umulhi(b1, h1.x & 0xFFFFu) + 2u) >> 2;
                const unsigned v1 = (__umulhi(b0, h0.x >> 16) + __umulhi(b1, h1.x >> 16) + 2u) >> 2;
                const unsigned v2 = (__umulhi(b0, h0.y & 0xFFFFu) + __umulhi(b1, h1.y & 0xFFFFu) + 2u) >> 2;
                const unsigned v3 = (__umulhi(b0, h0.y >> 16) + __umulhi(b1, h1.y >> 16) + 2u) >> 2;
                *reinterpret_cast<unsigned*>(R + (rrow0 + ry) * RP + 4 * (g0 + gq)) = v0 | (v1 << 8) | (v2 << 16) | (v3 << 24);   // <= 255 by construction
            }
        }
        __syncthreads();
    } else {
        // level 0: R comes straight from the frame; 16-byte chunks cover x in [X0-16, X0+TW+16)
        const bool vec = ((a.spitch & 15) == 0) && ((reinterpret_cast<uintptr_t>(srcf) & 15) == 0);
        const int rowsn = rh;
        if (vec) {
            for (int it = tid; it < rowsn * (RP / 16); it += NT) {
                const int ry = it / (RP / 16), c16 = it % (RP / 16);
                const int x = X0 - RO + c16 * 16;
                uint4 v = make_uint4(0, 0, 0, 0);
                if (x >= 0 && x < a.spitch) v = __ldg(reinterpret_cast<const uint4*>(srcf + (long long)(ay + ry) * a.spitch + x));
                *reinterpret_cast<uint4*>(R + (rrow0 + ry) * RP + c16 * 16) = v;
            }
        } else {
            for (int it = tid; it < rowsn * rw; it += NT) {
                const int ry = it / rw, c = it % rw;
                R[(rrow0 + ry) * RP + rcol0 + c] = __ldg(srcf + (long long)(ay + ry) * a.spitch + ax + c);
            }
        }
        __syncthreads();
    }

    // halo outside the image
    reflect_halo(R, X0, Y0, w, h);

    if (RESIZE) {
        // ---- store the level tile (16-byte stores; the row pitch absorbs the tail) ----
        uint8_t* dst = a.dst + (long long)f * a.dstride;
        for (int it = tid; it < th * (TW / 16); it += NT) {
            const int r = it >> 3, c16 = it & 7;
            if (c16 * 16 < tw)
                *reinterpret_cast<uint4*>(dst + (long long)(Y0 + r) * a.dpitch + X0 + c16 * 16) =
                    *reinterpret_cast<const uint4*>(R + (3 + r) * RP + RO + c16 * 16);
        }
    }
    if (!a.no_blur) blur_and_store(R, Hb, a.blur + (long long)f * a.bstride, a.bpitch, X0, Y0, tw, th);
}

}  // namespace fastp

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        else cudaGetLastError();
    }
    return fn;
}

// 3-D u8 tensor (x, y, frame) over a batch of pitched images; box = (bw, bh, 1).  Returns false when the layout
// is not TMA-legal (base 16-byte aligned, strides multiples of 16) so that the caller uses the plain-load path.
bool orbx_make_tensor_map_3d(void* map128, const uint8_t* base, int w, int h, int pitch, long long fstride, int frames, int bw, int bh);
static bool make_map(CUtensorMap* m, const uint8_t* base, int w, int h, int pitch, long long fstride, int frames, int bw, int bh)
{
    return orbx_make_tensor_map_3d(m, base, w, h, pitch, fstride, frames, bw, bh);
}
bool orbx_make_tensor_map_3d(void* map128, const uint8_t* base, int w, int h, int pitch, long long fstride, int frames, int bw, int bh)
{
    CUtensorMap* m = reinterpret_cast<CUtensorMap*>(map128);
    EncodeTiledFn enc = get_encode();
    if (!enc || (reinterpret_cast<uintptr_t>(base) & 15) || (pitch & 15) || (fstride & 15) || bw > 256 || bh > 256) return false;
    cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)frames};
    cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)fstride};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t es[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

void orbx_launch_pyramid(const OrbxGeom& g, const OrbxBuffers& b, const uint8_t* level0, int pitch0,
                         long long stride0, int batch, cudaStream_t s, int l_begin, int l_end, int parts)
{
    using namespace fastp;
    if (l_end < 0 || l_end > g.nlevels) l_end = g.nlevels;
    // the fast kernels need a scale factor <= 1.5 (source tile extents) and a blur pitch that is a multiple of 16
    bool ok = true;
    for (int l = 1; l < g.nlevels; l++) {
        const double sx = (double)g.lv[l - 1].w / g.lv[l].w, sy = (double)g.lv[l - 1].h / g.lv[l].h;
        if (sx > 1.55 || sy > 1.55) ok = false;
    }
    if (!ok) { launch_pyramid_generic(g, b, level0, pitch0, stride0, batch, s, l_begin, l_end, parts); return; }
    for (int l = l_begin; l < l_end; l++) {
        if (l == 0 && !(parts & ORBX_PYR_BLUR)) continue;                 // level 0 has no resize part
        const OrbxLevel& L = g.lv[l];
        Args a{};
        a.w = L.w; a.h = L.h;
        a.blur = b.blur[l]; a.bpitch = L.pitch; a.bstride = L.frame_stride;
        dim3 grid((L.w + TW - 1) / TW, (L.h + TH - 1) / TH, batch);
        if (l == 0 || parts == ORBX_PYR_BLUR) {
            // blur only: of the caller's frame (level 0) or of a level the resize part has already stored
            if (l == 0) { a.src = level0; a.spitch = pitch0; a.sstride = stride0; }
            else { a.src = b.pyr[l]; a.spitch = L.pitch; a.sstride = L.frame_stride; }
            a.sw = L.w; a.sh = L.h;
            const size_t smem = RH * RP + RH * TW * 2;
            ORBX_OPTIN_SMEM(k_pyr_fast<false>);
            CUtensorMap none; memset(&none, 0, sizeof(none));
            orbx_launch_pdl(k_pyr_fast<false>, grid, dim3(NT), smem, s, a, none);
        } else {
            const OrbxLevel& P = g.lv[l - 1];
            a.src = (l == 1) ? level0 : b.pyr[l - 1];
            a.spitch = (l == 1) ? pitch0 : P.pitch;
            a.sstride = (l == 1) ? stride0 : P.frame_stride;
            a.sw = P.w; a.sh = P.h;
            a.dst = b.pyr[l]; a.dpitch = L.pitch; a.dstride = L.frame_stride;
            a.xt = b.tabs + L.xtab_off; a.yt = b.tabs + L.ytab_off;
            const double sx = (double)P.w / L.w, sy = (double)P.h / L.h;
            a.sp = (((int)(RW * sx) + 2 + 15 + 16) + 15) & ~15;     // span + alignment slack, multiple of 16
            a.sh_max = (int)(RH * sy) + 4;
            a.no_blur = (parts & ORBX_PYR_BLUR) ? 0 : 1;
            const size_t hs_bytes = (size_t)a.sh_max * HP * 2 > (size_t)RH * TW * 2 ? (size_t)a.sh_max * HP * 2 : (size_t)RH * TW * 2;
            const size_t smem = (((size_t)a.sh_max * a.sp + 127) & ~(size_t)127) + RH * RP + hs_bytes + RH * sizeof(int4);
            ORBX_OPTIN_SMEM(k_pyr_fast<true>);
            CUtensorMap map; memset(&map, 0, sizeof(map));
            a.use_tma = make_map(&map, a.src, a.sw, a.sh, a.spitch, a.sstride, batch, a.sp, a.sh_max) ? 1 : 0;
            orbx_launch_pdl(k_pyr_fast<true>, grid, dim3(NT), smem, s, a, map);
        }
        ORBX_COUNT_LAUNCH(1);
    }
}
