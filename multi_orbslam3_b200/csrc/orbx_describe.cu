// orbx_describe.cu - output assembly, IC_Angle orientation and steered-BRIEF descriptors.
//
// Replaces: tail of ComputeKeyPointsOctTree (R/src/ORBextractor.cc:862-872: border offset, octave, size),
//           computeOrientation / IC_Angle (:75-102, :470-477),
//           computeOrbDescriptor / computeDescriptors (:105-145, :1059-1066),
//           the scale + two-ended mono/stereo placement of operator() (:1102-1149).
#include <cuda.h>
#include <stdlib.h>
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
    orbx_pdl_prologue();
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
    uint4* work = b.work + (long long)f * g.out_cap;

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
            work[i] = make_uint4(((p & 0xFFF) + ORBX_BORDER) | ((((p >> 12) & 0xFFF) + ORBX_BORDER) << 12) | ((uint32_t)lvl << 24),
                                 (uint32_t)slot, 0u, 0u);
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

constexpr int ORI_WARPS = 8;
constexpr int ORI_GROUP = 4;          // keypoints per warp in k_orient, processed interleaved (independent load chains)
constexpr int DESC_WARPS = 8;
constexpr int DESC_GROUP = 8;         // keypoints per warp in k_describe, one after the other through two staged boxes
constexpr int PATCH_R = 18;           // |rotated pattern offset| <= round(18.385) = 18
constexpr int PATCH_W = 80;           // box width in bytes.  The box must START on a 16-byte boundary of the image row (an unaligned
                                      // inner coordinate faults with "illegal instruction": profiles/probes/tma_box_probe_b200.txt),
                                      // so it begins at (cx - 18) & ~15 and needs 15 + 37 = 52 columns.  80 rather than 64: the row pitch
                                      // in shared memory is the box width and the BRIEF samples cluster around the patch centre; with
                                      // 16 words per row only two row phases exist, with 20 the rows cycle through 8 (0.440 -> 0.434 ms)
constexpr int PATCH_H = 2 * PATCH_R + 1;
constexpr int PATCH_BYTES = PATCH_W * PATCH_H;               // 2368
constexpr int PATCH_SLOT = (PATCH_BYTES + 127) & ~127;       // 2432: every buffer starts 128-byte aligned
constexpr int DESC_SMEM = DESC_WARPS * 2 * PATCH_SLOT;

// the 256 test pairs as floats, laid out for conflict-free 16-byte shared-memory reads: entry [k][lane] = (x0, y0, x1, y1) of
// pair 8 * lane + k (lane j produces descriptor byte j = tests 8j .. 8j+7)
__device__ float4 d_pattern_f[8 * 32];

struct DescMaps { CUtensorMap blur[ORBX_MAX_LEVELS]; };

// ---- TMA / mbarrier primitives (SASS: UTMALDG, SYNCS) ----
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
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

// ---- k_orient: IC_Angle + the rotation of every keypoint ------------------------------------------
// One warp per group of ORI_GROUP keypoints.  The phase is bound by the latency of the patch loads (every 31 x 31 patch is fresh
// from L2 / HBM), so the kernel keeps no large shared-memory state and runs at full occupancy, and a warp keeps the loads of its
// four keypoints in flight together.
//   moments: integer m10, m01 over the 749-pixel circular patch of the un-blurred level, read as aligned words dealt row-major to
//     the lanes (coalesced: ~6 sectors per load) and reduced with dp4a; the item table (byte mask, byte weights, row, word per
//     alignment) lives in shared memory;
//   angle: lane j = keypoint j: cv::fastAtan2 and ONE fp64 sincos per keypoint (the one-warp-per-keypoint kernel of round 1
//     executed both on all 32 lanes: 15 % of its instructions); angle -> the keypoint record, (cos, sin) -> the work item.
__global__ void __launch_bounds__(ORI_WARPS * 32, 5) k_orient(OrbxGeom g, OrbxBuffers b, const uint8_t* level0, int pitch0, long long stride0,
                                                          int first_slot)
{
    constexpr int G = ORI_GROUP;
    __shared__ uint4 s_ic[4 * 288];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = blockIdx.y;
    const int slot_f = first_slot + f;
    orbx_pdl_prologue();
    const int total = b.n[slot_f];
    // a CTA walks groups blockIdx.x, blockIdx.x + gridDim.x, ...: the table copy is paid once per CTA, not once per 32 keypoints
    if (blockIdx.x * ORI_WARPS * G >= total) return;
    for (int k = threadIdx.x; k < 4 * 288; k += ORI_WARPS * 32) s_ic[k] = d_ic_table[k];
    __syncthreads();
    uint4* work = b.work + (long long)f * g.out_cap;
    for (int i0 = (blockIdx.x * ORI_WARPS + warp) * G; i0 < total; i0 += gridDim.x * ORI_WARPS * G) {
        const int cnt = min(G, total - i0);
        // lane j < cnt holds the work item of keypoint i0 + j; the lanes of missing keypoints repeat keypoint i0 so that every
        // address below stays valid (their results are dropped)
        const uint4 wk = work[i0 + (lane < cnt ? lane : 0)];
        int m10[G], m01[G];
        const uint32_t* w0[G]; int pw[G]; const uint4* tab[G];
        bool aligned = true;
#pragma unroll
        for (int j = 0; j < G; j++) {
            const uint32_t w = __shfl_sync(0xffffffffu, wk.x, j);
            const int cx = w & 0xFFF, cy = (w >> 12) & 0xFFF, lvl = w >> 24;
            const uint8_t* img; int pitch;
            if (lvl == 0) { img = level0 + (long long)f * stride0; pitch = pitch0; }
            else { img = b.pyr[lvl] + (long long)f * g.lv[lvl].frame_stride; pitch = g.lv[lvl].pitch; }
            const uint8_t* p0 = img + (long long)(cy - ORBX_HALF_PATCH) * pitch + (cx - ORBX_HALF_PATCH);
            const int sh = (int)(reinterpret_cast<uintptr_t>(p0) & 3);    // same for every row when pitch % 4 == 0
            w0[j] = reinterpret_cast<const uint32_t*>(p0 - sh); pw[j] = pitch >> 2; tab[j] = s_ic + sh * 288;
            aligned = aligned && (pitch & 3) == 0;
            m10[j] = 0; m01[j] = 0;
        }
        if (aligned) {
#pragma unroll
            for (int it = 0; it < 9; it++) {
                const int t = lane + 32 * it;                              // item = (row, word), 279 items
                if (t < 31 * 9) {
#pragma unroll
                    for (int j = 0; j < G; j++) {
                        const uint4 e = tab[j][t];                         // mask, weights, row, word
                        const uint32_t mw = __ldg(w0[j] + (int)e.z * pw[j] + (int)e.w) & e.x;
                        asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(m10[j]) : "r"(mw), "r"(e.y), "r"(m10[j]));   // unsigned pixels x signed weights u
                        m01[j] += ((int)e.z - ORBX_HALF_PATCH) * (int)__dp4a(mw, 0x01010101u, 0u);
                    }
                }
            }
        } else {
            // a caller's level-0 pitch that is not a multiple of 4: plain byte loads, lane = patch row
#pragma unroll
            for (int j = 0; j < G; j++) {
                const uint32_t w = __shfl_sync(0xffffffffu, wk.x, j);
                const int cx = w & 0xFFF, cy = (w >> 12) & 0xFFF, lvl = w >> 24;
                const uint8_t* img; int pitch;
                if (lvl == 0) { img = level0 + (long long)f * stride0; pitch = pitch0; }
                else { img = b.pyr[lvl] + (long long)f * g.lv[lvl].frame_stride; pitch = g.lv[lvl].pitch; }
                if (lane < 31) {
                    const int v = lane - ORBX_HALF_PATCH;
                    const int d = c_umax[v < 0 ? -v : v];
                    const uint8_t* row = img + (long long)(cy + v) * pitch + cx;
                    int rs = 0;
                    for (int u = -d; u <= d; ++u) { const int val = __ldg(row + u); m10[j] += u * val; rs += val; }
                    m01[j] = v * rs;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < G; j++) { m10[j] = __reduce_add_sync(0xffffffffu, m10[j]); m01[j] = __reduce_add_sync(0xffffffffu, m01[j]); }
        int mm10 = m10[0], mm01 = m01[0];
#pragma unroll
        for (int j = 1; j < G; j++) if (lane == j) { mm10 = m10[j]; mm01 = m01[j]; }
        if (lane < cnt) {
            const float angle = fast_atan2_deg((float)mm01, (float)mm10);
            const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
            const float ang = __fmul_rn(angle, factorPI);
            double sd, cd;
            sincos((double)ang, &sd, &cd);          // same kernels as cos() / sin(), one range reduction
            b.kps[(long long)slot_f * g.out_cap + (int)wk.y].angle = angle;
            work[i0 + lane] = make_uint4(wk.x, wk.y, __float_as_uint((float)cd), __float_as_uint((float)sd));
        }
    }
}

// ---- k_describe: steered BRIEF from TMA-staged patches --------------------------------------------
// One warp per group of DESC_GROUP keypoints.  The 37 x 37 window of the BLURRED level that the rotated pattern can reach is
// staged in shared memory, one TMA box (80 x 37 bytes, cp.async.bulk.tensor on a per-level 3-D map) per keypoint, issued by one
// lane, double-buffered per warp and completed on an mbarrier: the 512 scattered byte reads of a descriptor hit shared memory
// instead of ~17 different L1 sectors per load instruction (with __ldg the kernel sat at 86 % of the L1TEX pipe, ncu r2_desc_v3).
// Lane j produces descriptor byte j; the rotated sample offsets use separately rounded fp32 products (no FMA) and round-half-even,
// exactly the reference's arithmetic.  TMA == false: the same with plain loads (driver without cuTensorMapEncodeTiled).
template <bool TMA>
__global__ void __launch_bounds__(DESC_WARPS * 32, 4) k_describe(OrbxGeom g, OrbxBuffers b, int first_slot, const __grid_constant__ DescMaps maps)
{
    constexpr int G = DESC_GROUP;
    extern __shared__ __align__(128) uint8_t s_patch[];                   // [warp][2][PATCH_SLOT] (TMA only)
    __shared__ float4 s_pat[8 * 32];
    __shared__ __align__(8) unsigned long long s_bar[DESC_WARPS][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = blockIdx.y;
    const int slot_f = first_slot + f;
    orbx_pdl_prologue();
    const int total = b.n[slot_f];
    if (blockIdx.x * DESC_WARPS * G >= total) return;                     // whole CTA past the frame's keypoints
    for (int k = threadIdx.x; k < 8 * 32; k += DESC_WARPS * 32) s_pat[k] = d_pattern_f[k];
    if (TMA && threadIdx.x < DESC_WARPS * 2) mbar_init(&s_bar[threadIdx.x >> 1][threadIdx.x & 1], 1);
    if (TMA) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const int i0 = (blockIdx.x * DESC_WARPS + warp) * G;
    if (i0 >= total) return;
    const int cnt = min(G, total - i0);
    const uint4 wk = b.work[(long long)f * g.out_cap + i0 + (lane < cnt ? lane : 0)];
    uint8_t* mybuf = s_patch + warp * 2 * PATCH_SLOT;
    // (a bulk tensor copy is issued by ONE lane: the instruction takes warp-uniform operands)
    auto issue = [&](int j) {
        const uint32_t w = __shfl_sync(0xffffffffu, wk.x, j);
        if (lane == 0) {
            const int cx = w & 0xFFF, cy = (w >> 12) & 0xFFF, lvl = w >> 24;
            mbar_expect_tx(&s_bar[warp][j & 1], PATCH_BYTES);
            tma_load_3d(mybuf + (j & 1) * PATCH_SLOT, &maps.blur[lvl], (cx - PATCH_R) & ~15, cy - PATCH_R, f, &s_bar[warp][j & 1]);
        }
    };
    if (TMA) {
        issue(0);
        if (cnt > 1) issue(1);
    }
    float4 pt[8];
#pragma unroll
    for (int k = 0; k < 8; k++) pt[k] = s_pat[k * 32 + lane];
#pragma unroll 2
    for (int j = 0; j < cnt; j++) {
        const uint32_t w = __shfl_sync(0xffffffffu, wk.x, j);
        const int slot = (int)__shfl_sync(0xffffffffu, wk.y, j);
        const float c = __uint_as_float(__shfl_sync(0xffffffffu, wk.z, j)), s_ = __uint_as_float(__shfl_sync(0xffffffffu, wk.w, j));
        const int cx = w & 0xFFF, cy = (w >> 12) & 0xFFF, lvl = w >> 24;
        const uint8_t* center; int bp;
        if (TMA) {
            center = mybuf + (j & 1) * PATCH_SLOT + PATCH_R * PATCH_W + PATCH_R + ((cx - PATCH_R) & 15);
            bp = PATCH_W;
            mbar_wait(&s_bar[warp][j & 1], (j >> 1) & 1);
        } else {
            bp = g.lv[lvl].pitch;
            center = b.blur[lvl] + (long long)f * g.lv[lvl].frame_stride + (long long)cy * bp + cx;
        }
        uint32_t val = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(pt[k].x, s_), __fmul_rn(pt[k].y, c)));
            const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(pt[k].x, c), __fmul_rn(pt[k].y, s_)));
            const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(pt[k].z, s_), __fmul_rn(pt[k].w, c)));
            const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(pt[k].z, c), __fmul_rn(pt[k].w, s_)));
            const int t0 = TMA ? (int)center[r0 * PATCH_W + c0] : (int)__ldg(center + r0 * bp + c0);
            const int t1 = TMA ? (int)center[r1 * PATCH_W + c1] : (int)__ldg(center + r1 * bp + c1);
            val |= (uint32_t)(t0 < t1) << k;
        }
        b.desc[((long long)slot_f * g.out_cap + slot) * 32 + lane] = (uint8_t)val;
        if (TMA && j + 2 < cnt) {
            __syncwarp();                                                 // every lane has read its samples: the buffer is free again
            issue(j + 2);
        }
    }
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
    static float4 pf[8 * 32];
    for (int lane = 0; lane < 32; lane++)
        for (int k = 0; k < 8; k++) {
            const int32_t* p = h_pattern + 4 * (8 * lane + k);
            pf[k * 32 + lane] = make_float4((float)p[0], (float)p[1], (float)p[2], (float)p[3]);
        }
    cudaMemcpyToSymbol(d_pattern_f, pf, sizeof(pf));
}

void orbx_launch_describe(const OrbxGeom& g, const OrbxBuffers& b, const uint8_t* level0, int pitch0,
                          long long stride0, int batch, int lap0, int lap1, int first_slot, cudaStream_t s)
{
    orbx_launch_pdl(k_finalize, dim3(batch), dim3(FIN_NT), 0, s, g, b, lap0, lap1, first_slot);
    // orientation: CTAs walk the keypoint groups of a frame in strides (the item table is copied once per CTA)
    {
        const int groups = (g.out_cap + ORI_WARPS * ORI_GROUP - 1) / (ORI_WARPS * ORI_GROUP);
        // big batches: 8 CTAs per frame amortise the table copy; a single frame spreads over the whole GPU instead
        dim3 grid(batch >= 32 && groups > 8 ? 8 : groups, batch);
        orbx_launch_pdl(k_orient, grid, dim3(ORI_WARPS * 32), 0, s, g, b, level0, pitch0, stride0, first_slot);
    }
    dim3 grid((g.out_cap + DESC_WARPS * DESC_GROUP - 1) / (DESC_WARPS * DESC_GROUP), batch);
    // one 3-D tensor map (x, y, frame) per blurred level with an 80 x 37 box; the blurred levels are this library's own buffers
    // (64-byte pitches), so the layout is always TMA-legal; if the driver entry point is missing the plain-load kernel runs
    alignas(64) DescMaps maps;
    memset(&maps, 0, sizeof(maps));
    bool tma = getenv("ORBX_DESC_NO_TMA") == nullptr;
    for (int l = 0; l < g.nlevels && tma; l++)
        tma = orbx_make_tensor_map_3d(&maps.blur[l], b.blur[l], g.lv[l].w, g.lv[l].h, g.lv[l].pitch, g.lv[l].frame_stride, batch, PATCH_W, PATCH_H);
    if (tma) {
        ORBX_OPTIN_SMEM(k_describe<true>);
        orbx_launch_pdl(k_describe<true>, grid, dim3(DESC_WARPS * 32), (size_t)DESC_SMEM, s, g, b, first_slot, maps);
    } else {
        orbx_launch_pdl(k_describe<false>, grid, dim3(DESC_WARPS * 32), 0, s, g, b, first_slot, maps);
    }
    ORBX_COUNT_LAUNCH(3);
}
