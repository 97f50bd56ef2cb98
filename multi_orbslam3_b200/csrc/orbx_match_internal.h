// orbx_match_internal.h - shared by the matcher translation units (orbx_match.cu, orbx_stereo.cu, orbx_search.cu):
// the 256-bit Hamming primitives, the warp top-2 reduction, the rotation histogram helpers and the matcher handle.
#pragma once
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
    float qminX, qminY;           // origin of the query cell range (= minX, minY for a Frame; the int-truncated KeyFrame::mnMinX/Y for a KeyFrame)
    double gate_chi2;             // > 0: per-candidate reprojection gate (Fuse), with the per-level 1 / sigma^2 below
    double gate_chi2_stereo;      // > 0: keypoints with a right coordinate use (ex^2 + ey^2 + er^2) against this bound instead
    float inv_sigma2[ORBX_MAX_LEVELS];
    unsigned* err;
};

}  // namespace

struct orbx_matcher {
    orbx_matcher_params p;
    int K, P, POOL;
    cudaStream_t stream;
    WinBufs W;
    // device staging for the host-pointer APIs (two frames + queries)
    orbx_keypoint* d_k1; orbx_keypoint* d_k2; uint8_t* d_d1; uint8_t* d_d2; uint8_t* d_qdesc; float* d_uright;
    float* d_prev; int32_t* d_out; int32_t* d_out2; int32_t* d_nm; float* d_sf;
    int32_t* d_knn_idx; int32_t* d_knn_dist;
    int32_t* d_pipe_knn; size_t pipe_knn_elems;   // BF kNN-2 tables of the host pipeline: idx [P][K][2] then dist [P][K][2]
    int32_t* d_part_idx; int32_t* d_part_dist; size_t part_elems;
    uint8_t* d_bfq; uint8_t* d_bft; size_t bfq_bytes, bft_bytes;
    unsigned* h_err;
    int32_t* d_pair_a; int32_t* d_pair_b;
    const orbx_keypoint* d_kps_src;          // optional replacement of the extractor's keypoints in the slot-based searches (mvKeysUn)
    // camera of the stream pipelines (orbx_matcher_set_camera): when set, every chunk is undistorted into d_kps_un
    bool cam_set; float cam_K[9], cam_P[9], cam_dist[12]; int cam_ndist;
    orbx_keypoint* d_kps_un; size_t kps_un_elems;
    uint8_t* d_gen; size_t gen_bytes;
    uint8_t* h_gen; size_t h_gen_bytes;       // pinned mirror of d_gen for the host-array searches (one copy each way)
    uint8_t* d_st; size_t st_bytes;          // stereo scratch
    int32_t* h_mono2; int mono2_cap;         // pinned monoIndex landing zone of the stereo pipeline (2 x batch)
    cudaStream_t s_bf; cudaEvent_t ev_bf_fork, ev_bf_join;      // few-pair path: the brute-force search runs beside the window search
    cudaStream_t s_h2d, s_d2h, s_match; cudaEvent_t ev[2 * ORBX_MAX_CHUNKS]; cudaEvent_t ev_ext[ORBX_MAX_CHUNKS]; cudaEvent_t ev_start;
    cudaEvent_t ev_r[2 * ORBX_MAX_CHUNKS];      // right camera of the stereo pipeline: [c] copy done, [MAX + c] extraction done
    // streaming form (orbx_stream_submit / orbx_stream_wait): two result staging sets on the device, so that the results of batch k
    // travel to the host while batch k+1 computes
    struct StreamSet {
        orbx_keypoint* kps; uint8_t* desc; int32_t* n; int32_t* mono; int32_t* m12; int32_t* nm; int32_t* knn_idx; int32_t* knn_dist;
        cudaEvent_t ev_kernels, ev_host; unsigned* h_err;      // h_err: pinned {extractor flags, matcher flags}
        bool busy; int batch;
    } st[2];
    size_t st_rows; int st_cap; long long st_ticket; bool st_init;
    // CUDA graph of the kernels of the single-chunk host call (the batch-1 latency path): captured on the second call with the same
    // arguments, replayed afterwards
    struct LatGraph {
        cudaGraphExec_t exec[2]; cudaGraph_t graph[2]; int nkernels[2]; int seen;      // [0] extraction, [1] matching + mailbox
        const void* ex; int batch, width, height, lap0, lap1, window, check_ori, knn, cam, geom_gen; float bounds[4]; float nnratio; const void* d_knn; const void* d_kps_un;
    } lg;
    // single-chunk host path: the small results (n, monoIndex, nmatches, both error words) reach the host through ONE kernel
    // that stores them into mapped pinned memory; ev_ex: extraction done (the keypoint / descriptor copies start on s_d2h beside
    // the matcher), ev_done / ev_exd2h: what the host waits for (the slot carry is queued after ev_done)
    int32_t* h_mail; int32_t* d_mail; cudaEvent_t ev_ex, ev_done, ev_exd2h, ev_carry;
    uint8_t* h_pack; uint8_t* d_pack; size_t pack_bytes;      // one pinned + one device block for the arguments of a host-array search call
    std::vector<void*> allocs;
};

// helpers shared across the matcher files
int orbx_m_ensure_pipeline(orbx_matcher* m);                 // side streams / events of the chunked host pipelines
int orbx_m_gen_scratch(orbx_matcher* m, size_t bytes);       // grows m->d_gen to at least `bytes`
