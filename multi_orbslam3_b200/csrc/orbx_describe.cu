// orbx_describe.cu - output assembly, IC_Angle orientation and steered-BRIEF descriptors.
//
// Replaces: tail of ComputeKeyPointsOctTree (R/src/ORBextractor.cc:862-872: border offset, octave, size),
//           computeOrientation / IC_Angle (:75-102, :470-477),
//           computeOrbDescriptor / computeDescriptors (:105-145, :1059-1066),
//           the scale + two-ended mono/stereo placement of operator() (:1102-1149).
#include <string.h>
#include "orbx_internal.h"

namespace {

// 256 test pairs, (x0,y0,x1,y1) as int8 packed into one 32-bit word per pair
__device__ uint32_t d_pattern[256];
// IC_Angle item table: for each byte alignment sh (0..3) of the patch's first column and each of the 31 x 9 aligned
// words of the patch: {byte mask of the pixels inside the circular patch, signed byte weights u, row, word}.
__device__ uint4 d_ic_table[4 * 288];
// IC_Angle patch half-widths per |v| (umax, R/src/ORBextractor.cc:452-467)
__device__ __constant__ int c_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

constexpr int FIN_NT = 256;

// ---- k_finalize: one CTA per frame ------------------------------------------------------------
// Concatenates the per-level octree outputs (levels in order, empty levels skipped), applies the border
// offset, scales coordinates (fp32 multiply for level != 0), decides the lapping-area side of every
// keypoint and turns the sequential monoIndex++/stereoIndex-- of the reference into prefix sums.
__global__ void __launch_bounds__(FIN_NT) k_finalize(OrbxGeom g, OrbxBuffers b, int lap0, int lap1, int first_slot)
{
    __shared__ int s_base[ORBX_MAX_LEVELS + 1];
    __shared__ int s_warp[FIN_NT / 32];
    __shared__ int s_run;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int f = blockIdx.x, slot_f = first_slot + f;
    const int* ln = b.lvl_n + (long long)f * g.nlevels;
    if (tid == 0) {
        int run = 0;
        for (int l = 0; l < g.nlevels; l++) { s_base[l] = run; run += ln[l]; }
        s_base[g.nlevels] = run;
        s_run = 0;
    }
    __syncthreads();
    int total = s_base[g.nlevels];
    if (total > g.out_cap) { if (tid == 0) atomicOr(b.err, ORBX_DEVERR_KP_OVERFLOW); total = g.out_cap; }
    const uint32_t* lk = b.lvl_kp + (long long)f * g.kp_total_cap;
    orbx_keypoint* okp = b.kps + (long long)slot_f * g.out_cap;
    uint2* work = b.work + (long long)f * g.out_cap;

    // total number of keypoints in the lapping area is needed up front? No: stereo slots count down from
    // total-1, mono slots count up from 0; both only need the running counts.
    for (int base = 0; base < total; base += FIN_NT) {
        const int i = base + tid;
        int lvl = 0; uint32_t p = 0; float sx = 0.f, sy = 0.f; int inlap = 0;
        if (i < total) {
            while (i >= s_base[lvl + 1]) lvl++;
            p = lk[g.lv[lvl].kp_base + (i - s_base[lvl])];
            sx = (float)((int)(p & 0xFFF) + ORBX_BORDER);
            sy = (float)((int)((p >> 12) & 0xFFF) + ORBX_BORDER);
            if (lvl != 0) { sx = __fmul_rn(sx, g.lv[lvl].scale); sy = __fmul_rn(sy, g.lv[lvl].scale); }
            inlap = (sx >= (float)lap0 && sx <= (float)lap1) ? 1 : 0;
        }
        // block exclusive scan of inlap
        int inc = inlap;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        int wbase = 0, tot = 0;
#pragma unroll
        for (int wi = 0; wi < FIN_NT / 32; wi++) { int t = s_warp[wi]; if (wi < warp) wbase += t; tot += t; }
        const int lap_before = s_run + wbase + inc - inlap;
        if (i < total) {
            const int slot = inlap ? (total - 1 - lap_before) : (i - lap_before);
            orbx_keypoint kp;
            kp.x = sx; kp.y = sy; kp.size = g.lv[lvl].size; kp.angle = -1.f;
            kp.response = (float)(p >> 24); kp.octave = lvl; kp.class_id = -1;
            okp[slot] = kp;
            // work item: level coords (with border), level, output slot
            work[i] = make_uint2(((p & 0xFFF) + ORBX_BORDER) | ((((p >> 12) & 0xFFF) + ORBX_BORDER) << 12) | ((uint32_t)lvl << 24),
                                 (uint32_t)slot);
        }
        __syncthreads();
        if (tid == 0) s_run += tot;
        __syncthreads();
    }
    if (tid == 0) {
        b.n[slot_f] = total;
        b.mono[slot_f] = total - s_run;    // monoIndex after the loop (:1149)
    }
}

// cv::fastAtan2 (OpenCV mathfuncs_core atan_f32): fp32, no FMA contraction
__device__ __forceinline__ float fast_atan2_deg(float y, float x)
{
    const float scale = (float)(180.0 / 3.141592653589793238462643383279502884);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    const float ax = fabsf(x), ay = fabsf(y);
    const float eps = (float)2.2204460492503131e-16;
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

constexpr int DESC_WARPS = 8;

// ---- k_orient_describe: one warp per keypoint -------------------------------------------------
// IC_Angle: lane v' = lane-15 owns patch row v = lane-15 (31 rows), integer moments, shuffle reduce.
// Descriptor: lane j produces byte j = tests 8j..8j+7; the rotated sample offsets use separately
// rounded fp32 products (no FMA) and round-half-even, cos/sin in fp64 rounded to fp32.
__global__ void __launch_bounds__(DESC_WARPS * 32) k_orient_describe(OrbxGeom g, OrbxBuffers b, const uint8_t* level0,
                                                                   int pitch0, long long stride0, int first_slot)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = blockIdx.y;
    const int slot_f = first_slot + f;
    const int total = b.n[slot_f];
    const int i = blockIdx.x * DESC_WARPS + warp;
    if (i >= total) return;
    const uint2 wk = b.work[(long long)f * g.out_cap + i];
    const int cx = wk.x & 0xFFF, cy = (wk.x >> 12) & 0xFFF, lvl = wk.x >> 24;
    const int slot = (int)wk.y;
    const OrbxLevel& L = g.lv[lvl];
    const uint8_t* img; int pitch;
    if (lvl == 0) { img = level0 + (long long)f * stride0; pitch = pitch0; }
    else { img = b.pyr[lvl] + (long long)f * L.frame_stride; pitch = L.pitch; }

    // ---- orientation: integer moments over the 749-pixel circular patch (rows v = -15..15, |u| <= umax[|v|]) ----
    // The patch is read as aligned 32-bit words, 9 per row; the 31 x 9 words are dealt to the lanes in row-major
    // order (neighbouring lanes read neighbouring words), masked to the row's extent and reduced with dp4a:
    // m10 += sum(u * I) uses signed byte weights u, m01 += v * sum(I).
    int m10 = 0, m01 = 0;
    {
        const uint8_t* p0 = img + (long long)(cy - ORBX_HALF_PATCH) * pitch + (cx - ORBX_HALF_PATCH);
        const int sh = (int)(reinterpret_cast<uintptr_t>(p0) & 3);        // same for every row when pitch % 4 == 0
        if ((pitch & 3) == 0) {
            const uint32_t* w0 = reinterpret_cast<const uint32_t*>(p0 - sh);
            const int pw = pitch >> 2;
            const uint4* tab = d_ic_table + sh * 288;
#pragma unroll
            for (int it = 0; it < 9; it++) {
                const int t = lane + 32 * it;                              // item = (row, word), 279 items
                if (t < 31 * 9) {
                    const uint4 e = __ldg(tab + t);                        // mask, weights, row, word
                    const uint32_t mw = __ldg(w0 + (int)e.z * pw + (int)e.w) & e.x;
                    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(m10) : "r"(mw), "r"(e.y), "r"(m10));   // unsigned pixels x signed weights u
                    m01 += ((int)e.z - ORBX_HALF_PATCH) * (int)__dp4a(mw, 0x01010101u, 0u);
                }
            }
        } else if (lane < 31) {
            const int v = lane - ORBX_HALF_PATCH;
            const int d = c_umax[v < 0 ? -v : v];
            const uint8_t* row = img + (long long)(cy + v) * pitch + cx;
            int rs = 0;
            for (int u = -d; u <= d; ++u) { const int val = __ldg(row + u); m10 += u * val; rs += val; }
            m01 = v * rs;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m10 += __shfl_xor_sync(0xffffffffu, m10, o);
        m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    // ---- descriptor on the blurred level ----
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    const float ang = __fmul_rn(angle, factorPI);
    double sd, cd;
    sincos((double)ang, &sd, &cd);                  // same kernels as cos() / sin(), one range reduction
    const float ca = (float)cd, sa = (float)sd;
    const uint8_t* bl = b.blur[lvl] + (long long)f * L.frame_stride;
    const uint8_t* center = bl + (long long)cy * L.pitch + cx;
    const uint4 w0 = reinterpret_cast<const uint4*>(d_pattern)[lane * 2];
    const uint4 w1 = reinterpret_cast<const uint4*>(d_pattern)[lane * 2 + 1];
    const uint32_t pw[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    uint32_t val = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const float x0 = (float)(int8_t)(pw[k] & 0xFF), y0 = (float)(int8_t)((pw[k] >> 8) & 0xFF);
        const float x1 = (float)(int8_t)((pw[k] >> 16) & 0xFF), y1 = (float)(int8_t)(pw[k] >> 24);
        const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, sa), __fmul_rn(y0, ca)));
        const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, ca), __fmul_rn(y0, sa)));
        const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, sa), __fmul_rn(y1, ca)));
        const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, ca), __fmul_rn(y1, sa)));
        const int t0 = __ldg(center + r0 * L.pitch + c0);
        const int t1 = __ldg(center + r1 * L.pitch + c1);
        val |= (uint32_t)(t0 < t1) << k;
    }
    b.desc[((long long)slot_f * g.out_cap + slot) * 32 + lane] = (uint8_t)val;
    if (lane == 0) b.kps[(long long)slot_f * g.out_cap + slot].angle = angle;
}

}  // namespace

static const int32_t h_pattern[1024] = {
#include "orb_pattern.inc"
};

static void upload_ic_table()
{
    static const int umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
    static uint4 tab[4 * 288];
    memset(tab, 0, sizeof(tab));
    for (int sh = 0; sh < 4; sh++)
        for (int t = 0; t < 31 * 9; t++) {
            const int row = t / 9, w = t % 9, v = row - ORBX_HALF_PATCH, d = umax[v < 0 ? -v : v];
            uint32_t mask = 0, wu = 0;
            for (int j = 0; j < 4; j++) {
                const int u = 4 * w + j - sh - ORBX_HALF_PATCH;          // byte j of word w sits at u
                if (u >= -d && u <= d) { mask |= 0xFFu << (8 * j); wu |= (uint32_t)(uint8_t)(int8_t)u << (8 * j); }
            }
            tab[sh * 288 + t] = make_uint4(mask, wu, (uint32_t)row, (uint32_t)w);
        }
    cudaMemcpyToSymbol(d_ic_table, tab, sizeof(tab));
}

void orbx_upload_pattern()
{
    upload_ic_table();
    uint32_t packed[256];
    for (int i = 0; i < 256; i++) {
        const int32_t* p = h_pattern + 4 * i;
        packed[i] = (uint32_t)(uint8_t)(int8_t)p[0] | ((uint32_t)(uint8_t)(int8_t)p[1] << 8) |
                    ((uint32_t)(uint8_t)(int8_t)p[2] << 16) | ((uint32_t)(uint8_t)(int8_t)p[3] << 24);
    }
    cudaMemcpyToSymbol(d_pattern, packed, sizeof(packed));
}

void orbx_launch_describe(const OrbxGeom& g, const OrbxBuffers& b, const uint8_t* level0, int pitch0,
                          long long stride0, int batch, int lap0, int lap1, int first_slot, cudaStream_t s)
{
    k_finalize<<<batch, FIN_NT, 0, s>>>(g, b, lap0, lap1, first_slot);
    dim3 grid((g.out_cap + DESC_WARPS - 1) / DESC_WARPS, batch);
    k_orient_describe<<<grid, DESC_WARPS * 32, 0, s>>>(g, b, level0, pitch0, stride0, first_slot);
    ORBX_COUNT_LAUNCH(2);
}
