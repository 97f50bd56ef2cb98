// orbx_internal.h - shared declarations of the sm_100a ORB front-end kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include "../../include/orbx.h"

#define ORBX_EDGE 19          // EDGE_THRESHOLD, R/src/ORBextractor.cc:72
#define ORBX_BORDER 16        // minBorderX = EDGE_THRESHOLD-3, R/src/ORBextractor.cc:771
#define ORBX_HALF_PATCH 15    // HALF_PATCH_SIZE, R/src/ORBextractor.cc:71
#define ORBX_FAST_TP 288      // staged width limit (bytes) of one FAST segment tile
#define ORBX_MAX_UNITS 2048   // FAST segments (cell row x segment) per level
#define ORBX_OCTREE_MAX_NODES 4096   // k_octree: MAX_IPT (8) node items per thread x 512 threads; creation index packed in 12 bits
#define ORBX_FAST_W 30        // cell size W, R/src/ORBextractor.cc:767

// device error flag bits (written by kernels, read back at sync/download)
#define ORBX_DEVERR_CAND_OVERFLOW  1u   // more FAST corners in a cell row / level than the buffers hold
#define ORBX_DEVERR_OCTREE_DEPTH   2u   // octree node could not be separated within the key depth
#define ORBX_DEVERR_NODE_OVERFLOW  4u   // octree node table overflow
#define ORBX_DEVERR_KP_OVERFLOW    8u   // more keypoints than the result capacity
#define ORBX_DEVERR_POOL_OVERFLOW 16u   // matcher candidate pool overflow

// Geometry of one pyramid level and of its FAST cell grid (R/src/ORBextractor.cc:769-804)
struct OrbxLevel {
    int w, h;            // level size
    int pitch;           // bytes per row
    long long frame_stride;   // bytes per frame in the batched level buffer
    int maxBX, maxBY;    // maxBorderX/Y = dim - 16
    int nCols, nRows;    // cells
    int wCell, hCell;
    int nSeg, segCells;  // FAST segments per cell row, cells per segment
    int quota;           // mnFeaturesPerLevel
    int nIni;            // octree root count
    float hX;            // octree root width
    float scale;         // mvScaleFactor
    float size;          // keypoint.size = int(31*scale)
    int row_base;        // index of this level's first segment (cell row x segment, row-major) in the flattened list
    int row_cap;         // candidate capacity of one segment
    int cand_cap;        // candidate capacity of the level (octree input)
    int kp_cap;          // kept-keypoint capacity of the level
    int kp_base;         // offset of this level in the per-frame kept-keypoint buffer
    int xtab_off, ytab_off;   // offsets (in short4 units) of the resize tables
};

struct OrbxGeom {
    int nlevels;
    int width, height;
    int total_rows;      // sum of nRows * nSeg over levels
    int kp_total_cap;    // per-frame kept-keypoint buffer length (sum of kp_cap)
    int out_cap;         // per-frame result capacity
    int ini_th, min_th;
    OrbxLevel lv[ORBX_MAX_LEVELS];
};

// Device buffers of one extractor handle.
struct OrbxBuffers {
    uint8_t* pyr[ORBX_MAX_LEVELS];     // level images; pyr[0] may alias the caller's frames
    uint8_t* blur[ORBX_MAX_LEVELS];    // blurred levels
    short4* tabs;                      // resize tables: (s0, s1, c0, c1) per dst column / row
    uint32_t* row_cand;                // [batch][total_rows][row_cap(level)] packed x | y<<12 | score<<24 ... flattened by row offsets
    int* row_count;                    // [batch][total_rows]
    long long row_cand_stride;         // elements per frame
    int4* unit_tab;                    // [total_rows][4] FAST segment records (orbx_fast_units)
    int* row_off;                      // [total_rows] element offset of each cell row inside a frame
    uint32_t* lvl_kp;                  // [batch][kp_total_cap] kept keypoints per level, packed
    int* lvl_n;                        // [batch][nlevels]
    unsigned long long* sort_scratch;  // global sort scratch for oversized levels [batch][nlevels][...]
    long long sort_scratch_stride;     // per (frame) elements
    int* sort_off;                     // [nlevels] offsets into a frame's scratch
    uint4* work;                       // [batch][out_cap] work items: level coords | level, output slot, cos, sin (k_orient fills the last two)
    orbx_keypoint* kps;                // [slots][out_cap]
    uint8_t* desc;                     // [slots][out_cap][32]
    int* n;                            // [slots]
    int* mono;                         // [slots]
    unsigned int* err;                 // device error flags
};

#ifdef __CUDACC__
__device__ __forceinline__ int orbx_reflect101(int p, int n)
{
    // BORDER_REFLECT_101 for |overshoot| < n (n >= 2); n == 1 -> 0
    if (n == 1) return 0;
    if (p < 0) p = -p;
    if (p >= n) p = 2 * (n - 1) - p;
    return p;
}
#endif

// ---- programmatic dependent launch (PDL) ----
// The kernels of the per-frame chain (pyramid .. kNN merge) can be launched with cudaLaunchAttributeProgrammaticStreamSerialization:
// every CTA lets the NEXT kernel of the stream be scheduled as soon as it has started (griddepcontrol.launch_dependents) and
// waits for the PREVIOUS kernel to have completed and flushed (griddepcontrol.wait) before it touches global memory, so the
// launch latency and the CTA start-up of kernel k+1 hide behind the tail of kernel k.  Measured on a B200: one frame per call
// 0.303 -> 0.262 ms with direct launches (a few us with the launch graphs); C1 (512 mono frames) 3.415 -> 3.392 ms per step; but
// C2 3.79 -> 3.86 ms and C3 (1241-wide stereo frames) 4.73 -> 5.35 ms: there the early-resident CTAs of the next kernel hold
// shared memory and registers while they wait.  Policy (orbx_pdl_enabled): on inside a few-frame step only; ORBX_PDL=1 / 0
// forces it on / off everywhere.  The prologue is a no-op in a kernel launched without the attribute.
// RULE: a kernel launched through orbx_launch_pdl MUST call orbx_pdl_prologue() before its first global access.
#ifdef __CUDACC__
__device__ __forceinline__ void orbx_pdl_prologue()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
#endif
bool orbx_pdl_enabled();
extern thread_local int g_orbx_pdl_scope;
struct OrbxPdlScope {                       // marks the launches of a few-frame step (see orbx_pdl_enabled)
    bool on;
    explicit OrbxPdlScope(bool few) : on(few) { if (on) g_orbx_pdl_scope++; }
    ~OrbxPdlScope() { if (on) g_orbx_pdl_scope--; }
};
template <typename... KArgs, typename... Args>
inline cudaError_t orbx_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = orbx_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// kernel launchers (defined in the .cu files)
// (l_begin, l_end): the levels a launch covers, default all; the single-frame path launches level by level on branch streams
// parts: the resize + store of a level and its blur can be launched apart (blur: from the stored level)
#define ORBX_PYR_RESIZE 1
#define ORBX_PYR_BLUR 2
void orbx_launch_pyramid(const OrbxGeom& g, const OrbxBuffers& b, const uint8_t* level0, int pitch0,
                         long long stride0, int batch, cudaStream_t s, int l_begin = 0, int l_end = -1, int parts = ORBX_PYR_RESIZE | ORBX_PYR_BLUR);
void orbx_launch_fast(const OrbxGeom& g, const OrbxBuffers& b, const uint8_t* level0, int pitch0,
                      long long stride0, int batch, cudaStream_t s, int l_begin = 0, int l_end = -1);
void orbx_launch_octree(const OrbxGeom& g, const OrbxBuffers& b, int batch, cudaStream_t s, int l_begin = 0, int l_end = -1);
void orbx_launch_describe(const OrbxGeom& g, const OrbxBuffers& b, const uint8_t* level0, int pitch0,
                          long long stride0, int batch, int lap0, int lap1, int first_slot, cudaStream_t s);
int  orbx_octree_smem_bytes(const OrbxGeom& g);
void orbx_octree_configure(const OrbxGeom& g);
void orbx_fast_configure(const OrbxGeom& g);
int orbx_fast_plan(int w, int nCols, int wCell);
void orbx_fast_units(const OrbxGeom& g, std::vector<int4>& tab);
void orbx_upload_pattern();
// 3-D u8 tensor map (x, y, frame) over a batch of pitched images, box = (bw, bh, 1); map128 points at a CUtensorMap (128 bytes,
// 64-byte aligned).  false when the layout is not TMA-legal (base / strides not multiples of 16) or the driver entry is missing.
bool orbx_make_tensor_map_3d(void* map128, const uint8_t* base, int w, int h, int pitch, long long fstride, int frames, int bw, int bh);
void orbx_set_error(const char* fmt, const char* a, const char* b);
// Raises a kernel's dynamic shared-memory limit to the opt-in maximum of the CURRENT device, once per (kernel, device).
// The attribute is per function and per device, shared by every handle of the process: it is never tied to one handle's
// configuration and never lowered, so handles with different sizes and handles on different devices cannot interfere.
cudaError_t orbx_optin_smem(const void* kernel);
#define ORBX_OPTIN_SMEM(k) orbx_optin_smem(reinterpret_cast<const void*>(k))
// number of kernels this library launched since load (bench.py reports the delta as gpu_launches)
#include <atomic>
extern std::atomic<unsigned long long> g_orbx_launches;     // extractor instances may run on different host threads
#define ORBX_COUNT_LAUNCH(n) (g_orbx_launches.fetch_add((n), std::memory_order_relaxed))

// host-buffer pipeline pieces of the extractor (orbx_api.cu), used by orbx_extract_match_batch
struct orbx_extractor;
int orbx_ex_configure(orbx_extractor* h, int width, int height);
int orbx_ex_stage_input(orbx_extractor* h, const uint8_t* imgs, int f0, int count, int width, int height, int stride,
                        size_t frame_stride, cudaStream_t s);
int orbx_ex_run_staged(orbx_extractor* h, int f0, int count, int lap0, int lap1, int first_slot, cudaStream_t s);
int orbx_ex_run_device(orbx_extractor* h, const uint8_t* d_imgs, int pitch, long long fstride, int f0, int count,
                       int lap0, int lap1, int first_slot, cudaStream_t s);
cudaStream_t orbx_ex_stream(orbx_extractor* h);
int orbx_ex_out_cap(orbx_extractor* h);
bool orbx_ex_can_fetch_direct(orbx_extractor* h, orbx_keypoint* kps, uint8_t* desc, int cap, int32_t* n, int32_t* mono_index);
int orbx_ex_fetch_async(orbx_extractor* h, int first_slot, int count, int host_off, orbx_keypoint* kps, uint8_t* desc, int cap,
                        int32_t* n, int32_t* mono_index, cudaStream_t s, bool direct, bool counts = true);
void orbx_ex_set_fetched(orbx_extractor* h, unsigned err, const int32_t* n, const int32_t* mono, int count, int32_t* user_n, int32_t* user_mono, bool direct);
int orbx_ex_fetch_finish(orbx_extractor* h, int count, orbx_keypoint* kps, uint8_t* desc, int cap,
                         int32_t* n, int32_t* mono_index, bool direct, bool err_fetched = false);
int orbx_ex_fetch_err_async(orbx_extractor* h, cudaStream_t s);
// input prefetch of the host pipeline (orbx_extract_match_batch_prefetch)
int orbx_ex_prefetch(orbx_extractor* h, const uint8_t* imgs, int batch, int width, int height, int stride, size_t frame_stride, cudaStream_t s_copy);
bool orbx_ex_take_prefetched(orbx_extractor* h, const uint8_t* imgs, int batch, int width, int height, const uint8_t** d_frames, cudaEvent_t* ready);
int orbx_ex_pitch0(orbx_extractor* h);
int orbx_ex_prefetch_mark_read(orbx_extractor* h, const uint8_t* d_frames, cudaStream_t s);
int orbx_ex_geom_gen(orbx_extractor* h);
unsigned* orbx_ex_err_device(orbx_extractor* h);
bool orbx_ex_profiling(orbx_extractor* h);
bool orbx_host_pinned(const void* p);
long long orbx_ex_stride0(orbx_extractor* h);

// device view of one frame's pyramid of an extractor handle (stereo refinement reads both cameras' pyramids)
struct OrbxPyrView {
    const uint8_t* lv[ORBX_MAX_LEVELS];
    int pitch[ORBX_MAX_LEVELS], w[ORBX_MAX_LEVELS], h[ORBX_MAX_LEVELS];
    long long fstride[ORBX_MAX_LEVELS];     // bytes between consecutive frames of the batch
    float scale[ORBX_MAX_LEVELS], inv_scale[ORBX_MAX_LEVELS];
    int nlevels;
};
int orbx_ex_pyramid_view(orbx_extractor* h, int frame, OrbxPyrView* out);
int orbx_ex_device(orbx_extractor* h);
int orbx_ex_max_batch(orbx_extractor* h);
