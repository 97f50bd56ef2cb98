// orbx_api.cu - C ABI of the extractor (include/orbx.h): geometry tables, device buffers, launch order.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <set>
#include <string>
#include <utility>
#include <vector>
#include "orbx_internal.h"

std::atomic<unsigned long long> g_orbx_launches{0};
// PDL policy: ORBX_PDL=1 always, ORBX_PDL=0 never, unset: only inside a few-frame step (OrbxPdlScope set by run_batch /
// match_slots_impl for <= 2 frames / pairs), where the kernels are a handful of CTAs and early-resident dependents cost nothing
thread_local int g_orbx_pdl_scope = 0;
bool orbx_pdl_enabled()
{
    static const int env = getenv("ORBX_PDL") ? atoi(getenv("ORBX_PDL")) : -1;
    return env == 1 || (env < 0 && g_orbx_pdl_scope > 0);
}
extern "C" unsigned long long orbx_launch_count(void) { return g_orbx_launches.load(std::memory_order_relaxed); }

static thread_local std::string g_last_error;
void orbx_set_error(const char* fmt, const char* a, const char* b2)
{
    char buf[512];
    snprintf(buf, sizeof(buf), fmt, a, b2);
    g_last_error = buf;
}
extern "C" const char* orbx_last_error(void) { return g_last_error.c_str(); }

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            orbx_set_error("%s failed: %s", #call, cudaGetErrorString(e_));               \
            return ORBX_E_CUDA;                                                           \
        }                                                                                 \
    } while (0)

cudaError_t orbx_optin_smem(const void* kernel)
{
    static std::mutex mu;
    static std::set<std::pair<const void*, int>> done;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    if (done.count(std::make_pair(kernel, dev))) return cudaSuccess;
    int optin = 0;
    if ((e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess) return e;
    cudaFuncAttributes fa;
    if ((e = cudaFuncGetAttributes(&fa, kernel)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes)) != cudaSuccess) return e;
    done.insert(std::make_pair(kernel, dev));
    return cudaSuccess;
}

extern "C" int orbx_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

static inline int cv_round_f(float v) { return (int)lrintf(v); }

struct orbx_extractor {
    orbx_params p;
    double scaleFactor;
    std::vector<float> scale, invScale, sigma2, invSigma2;
    std::vector<int> featuresPerLevel;
    int slots;               // result slots = max_batch + 1
    OrbxGeom geom;           // geometry of the currently configured image size (width == 0: none)
    OrbxBuffers buf;
    cudaStream_t stream;
    // level-0 staging for host images
    uint8_t* d_level0; int pitch0; long long stride0;
    uint8_t* h_stage_in;     // pinned
    uint8_t* d_raw; size_t raw_bytes;   // device landing zone for host rows whose stride is not the staging pitch
    // input prefetch (orbx_extract_match_batch_prefetch): two extra staging buffers; entries are consumed first in, first out
    struct Prefetch { const uint8_t* src; int batch, width, height; uint8_t* buf; cudaEvent_t ev; };
    uint8_t* pf_buf[2]; size_t pf_bytes; cudaEvent_t pf_ev[2]; Prefetch pf[2]; int pf_count, pf_next;
    cudaEvent_t pf_read[2]; bool pf_read_set[2];   // recorded on the kernel stream after the kernels that read pf_buf[i] were queued
    int geom_gen;                                  // counts (re)configurations: cached launch graphs of other handles key on it
    orbx_keypoint* h_kps; uint8_t* h_desc; int* h_n; int* h_mono; unsigned* h_err;   // pinned
    // what the last batch used as level 0 (for pyramid_to_host)
    const uint8_t* last_level0; int last_pitch0; long long last_stride0; int last_batch;
    std::vector<void*> allocs;
    // optional per-stage CUDA-event timing (bench.py): pyramid+blur | FAST | octree | finalize+orient+describe
    // a ring of event sets so that recording never makes the host wait inside a timed region; harvested on query
    bool profile; std::vector<cudaEvent_t> ev; int ev_head, ev_count; double stage_ms[4]; int stage_batches;
    // level-parallel launch order of the few-frame path (run_batch_dag): one branch stream per level + one for the level-0 blur
    cudaStream_t br[2 * ORBX_MAX_LEVELS]; cudaEvent_t ev_lvl[ORBX_MAX_LEVELS], ev_br[2 * ORBX_MAX_LEVELS]; bool dag_init;
    // launch graph of the one- / two-frame orbx_extract_batch call (the class API's operator()): seen 1 = ran directly once,
    // 2 = captured, -1 = not capturable; keyed on the arguments and the configuration generation
    struct SmallGraph { cudaGraphExec_t exec; cudaGraph_t graph; int seen, batch, lap0, lap1, gen, nkernels; } xg;
    int32_t* h_mailx; int32_t* d_mailx;          // mapped pinned: {error word, n[2], monoIndex[2]} written by k_ex_mailbox
    uint8_t* h_pyr_stage; size_t pyr_stage_bytes;   // pinned staging of orbx_pyramid_levels_to_host
};
#define ORBX_EV_SETS 512

static void harvest_stage_times(orbx_extractor* h, bool all)
{
    // oldest first; `all` == false only frees one set when the ring is full
    while (h->ev_count > 0) {
        const int set = (h->ev_head - h->ev_count + ORBX_EV_SETS) % ORBX_EV_SETS;
        cudaEvent_t* e = h->ev.data() + set * 5;
        if (cudaEventSynchronize(e[4]) != cudaSuccess) { cudaGetLastError(); break; }
        for (int i = 0; i < 4; i++) { float ms = 0; cudaEventElapsedTime(&ms, e[i], e[i + 1]); h->stage_ms[i] += ms; }
        h->stage_batches++;
        h->ev_count--;
        if (!all) break;
    }
}

static int dev_alloc(orbx_extractor* h, void** p, size_t bytes)
{
    if (bytes == 0) bytes = 16;
    CK(cudaMalloc(p, bytes));
    h->allocs.push_back(*p);
    CK(cudaMemset(*p, 0, bytes));        // capacity tails that travel to the host with the results are zeros, not stale memory
    return ORBX_OK;
}

// cv::resize table for one axis (OpenCV imgproc resize.cpp: fixed-point INTER_LINEAR, coefficient bits = 11)
static void axis_table(int dn, int sn, bool horizontal, short4* out)
{
    const double inv_scale = (double)dn / sn, scale = 1. / inv_scale;
    for (int d = 0; d < dn; d++) {
        float fx = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(fx);
        fx -= s;
        int s0, s1;
        if (horizontal) {
            if (s < 0) { fx = 0; s = 0; }
            if (s >= sn - 1) { fx = 0; s = sn - 1; }
            s0 = s; s1 = s + 1 < sn ? s + 1 : sn - 1;
        } else {
            s0 = s < 0 ? 0 : (s > sn - 1 ? sn - 1 : s);
            s1 = s + 1 < 0 ? 0 : (s + 1 > sn - 1 ? sn - 1 : s + 1);
        }
        out[d].x = (short)s0; out[d].y = (short)s1;
        out[d].z = (short)cv_round_f((1.f - fx) * 2048.f);
        out[d].w = (short)cv_round_f(fx * 2048.f);
    }
}

static int configure_geometry_impl(orbx_extractor* h, int width, int height);
// The geometry is committed only when every table and buffer of it exists: a failure half way (out of memory, a bad size)
// leaves the handle unconfigured (width == 0) so that the next call starts over instead of launching on partial buffers.
static int configure_geometry(orbx_extractor* h, int width, int height)
{
    if (h->geom.width == width && h->geom.height == height) return ORBX_OK;
    h->geom_gen++;
    const int rc = configure_geometry_impl(h, width, height);
    if (rc != ORBX_OK) { h->geom.width = 0; h->geom.height = 0; }
    return rc;
}

static int configure_geometry_impl(orbx_extractor* h, int width, int height)
{
    OrbxGeom& g = h->geom;
    if (width > h->p.max_width || height > h->p.max_height) {
        orbx_set_error("%s%s", "image larger than max_width/max_height of the handle", "");
        return ORBX_E_INVALID;
    }
    memset(&g, 0, sizeof(g));
    g.nlevels = h->p.nlevels; g.width = width; g.height = height;
    g.ini_th = h->p.ini_th_fast; g.min_th = h->p.min_th_fast;
    const int B = h->p.max_batch;
    int rows = 0, kpbase = 0, tab = 0;
    const int maxcand = h->p.max_candidates_per_level > 0 ? h->p.max_candidates_per_level : 16384;
    std::vector<int> row_off; std::vector<int> sort_off;
    long long row_elems = 0, sort_elems = 0;
    for (int l = 0; l < g.nlevels; l++) {
        OrbxLevel& L = g.lv[l];
        const float sc = h->invScale[l];
        L.w = cv_round_f((float)width * sc); L.h = cv_round_f((float)height * sc);   // R/src/ORBextractor.cc:1157
        if (L.w < 8 || L.h < 8 || L.w > 4096 + 2 * ORBX_BORDER || L.h > 4096 + 2 * ORBX_BORDER) {
            orbx_set_error("%s%s", "unsupported level size (need 8..4128 px per side at every level)", "");
            g.width = 0; return ORBX_E_INVALID;
        }
        L.pitch = (L.w + 63) & ~63;
        L.frame_stride = (long long)L.pitch * L.h;
        L.maxBX = L.w - ORBX_BORDER; L.maxBY = L.h - ORBX_BORDER;
        const float fw = (float)(L.maxBX - ORBX_BORDER), fh = (float)(L.maxBY - ORBX_BORDER);
        L.nCols = fw > 0 ? (int)(fw / (float)ORBX_FAST_W) : 0;
        L.nRows = fh > 0 ? (int)(fh / (float)ORBX_FAST_W) : 0;
        if (L.nCols <= 0 || L.nRows <= 0) { L.nCols = L.nRows = 0; L.wCell = L.hCell = 0; }
        else { L.wCell = (int)ceilf(fw / L.nCols); L.hCell = (int)ceilf(fh / L.nRows); }
        if (L.nCols > 128 || L.nRows > 127) {
            orbx_set_error("%s%s", "image too large for the FAST cell tables", ""); g.width = 0; return ORBX_E_INVALID;
        }
        L.quota = h->featuresPerLevel[l];
        L.scale = h->scale[l];
        L.size = (float)(int)(31 * h->scale[l]);                                    // :862
        if (L.nRows > 0) {
            L.nIni = (int)roundf((float)(L.maxBX - ORBX_BORDER) / (L.maxBY - ORBX_BORDER));   // :541
            if (L.nIni < 1) L.nIni = 1;
            if (L.nIni > 63) { orbx_set_error("%s%s", "aspect ratio too extreme for the octree key", ""); g.width = 0; return ORBX_E_INVALID; }
            L.hX = (float)(L.maxBX - ORBX_BORDER) / L.nIni;                         // :543
        } else { L.nIni = 1; L.hX = 1.f; }
        L.row_base = rows;
        L.segCells = orbx_fast_plan(L.w, L.nCols, L.wCell);
        L.nSeg = L.nCols > 0 ? (L.nCols + L.segCells - 1) / L.segCells : 0;
        if (L.nRows * L.nSeg > ORBX_MAX_UNITS) {
            orbx_set_error("%s%s", "image too large for the FAST segment tables", ""); g.width = 0; return ORBX_E_INVALID;
        }
        // capacity of one segment: in-cell NMS leaves at most ceil(w/2)*ceil(h/2) corners per cell
        long long rowcap = 0, rowsum = 0;
        for (int sg = 0; sg < L.nSeg; sg++) {
            long long segcap = 0;
            for (int j = sg * L.segCells; j < L.nCols && j < (sg + 1) * L.segCells; j++) {
                int c0 = j * L.wCell, c1 = c0 + L.wCell; const int ws = L.maxBX - ORBX_BORDER - 6;
                if (c1 > ws) c1 = ws;
                if (c1 > c0) segcap += (long long)((c1 - c0 + 1) / 2) * ((L.hCell + 1) / 2);
            }
            rowsum += segcap;
            if (segcap > rowcap) rowcap = segcap;
        }
        if (rowcap > maxcand) rowcap = maxcand;
        L.row_cap = (int)rowcap;
        long long lc = rowsum * L.nRows; if (lc > maxcand) lc = maxcand;
        L.cand_cap = (int)lc;
        for (int r = 0; r < L.nRows * L.nSeg; r++) { row_off.push_back((int)row_elems); row_elems += L.row_cap; }
        rows += L.nRows * L.nSeg;
        L.kp_cap = (L.quota > 4 * L.nIni ? L.quota : 4 * L.nIni) + 4;
        L.kp_base = kpbase; kpbase += L.kp_cap;
        L.xtab_off = tab; tab += L.w;
        L.ytab_off = tab; tab += L.h;
        int npad = 1; while (npad < L.cand_cap) npad <<= 1;
        sort_off.push_back((int)sort_elems);
        sort_elems += (npad > 4096) ? (long long)npad + npad / 2 : 0;
    }
    g.total_rows = rows;
    g.kp_total_cap = kpbase;
    g.out_cap = orbx_extractor_max_keypoints(h);

    // ---- (re)allocate device buffers ----
    for (void* p : h->allocs) cudaFree(p);
    h->allocs.clear();
    OrbxBuffers& b = h->buf;
    memset(&b, 0, sizeof(b));
    int rc;
    for (int l = 0; l < g.nlevels; l++) {
        if (l > 0 && (rc = dev_alloc(h, (void**)&b.pyr[l], (size_t)g.lv[l].frame_stride * B))) return rc;
        if ((rc = dev_alloc(h, (void**)&b.blur[l], (size_t)g.lv[l].frame_stride * B))) return rc;
    }
    // staged level 0: tight 16-byte-multiple pitch so that a pinned frame batch with the same stride is ONE contiguous DMA
    h->pitch0 = (g.lv[0].w + 15) & ~15; h->stride0 = (long long)h->pitch0 * g.lv[0].h;
    if ((rc = dev_alloc(h, (void**)&h->d_level0, (size_t)h->stride0 * B))) return rc;
    std::vector<short4> tabs(tab);
    for (int l = 1; l < g.nlevels; l++) {
        axis_table(g.lv[l].w, g.lv[l - 1].w, true, tabs.data() + g.lv[l].xtab_off);
        axis_table(g.lv[l].h, g.lv[l - 1].h, false, tabs.data() + g.lv[l].ytab_off);
    }
    if ((rc = dev_alloc(h, (void**)&b.tabs, sizeof(short4) * tab))) return rc;
    CK(cudaMemcpy(b.tabs, tabs.data(), sizeof(short4) * tab, cudaMemcpyHostToDevice));
    b.row_cand_stride = row_elems;
    if ((rc = dev_alloc(h, (void**)&b.row_cand, sizeof(uint32_t) * (size_t)row_elems * B))) return rc;
    if ((rc = dev_alloc(h, (void**)&b.row_count, sizeof(int) * (size_t)rows * B))) return rc;
    if ((rc = dev_alloc(h, (void**)&b.row_off, sizeof(int) * (rows + 1)))) return rc;
    if (rows) CK(cudaMemcpy(b.row_off, row_off.data(), sizeof(int) * rows, cudaMemcpyHostToDevice));
    {
        std::vector<int4> units;
        orbx_fast_units(g, units);
        if ((rc = dev_alloc(h, (void**)&b.unit_tab, sizeof(int4) * (units.size() + 1)))) return rc;
        if (!units.empty()) CK(cudaMemcpy(b.unit_tab, units.data(), sizeof(int4) * units.size(), cudaMemcpyHostToDevice));
    }
    if ((rc = dev_alloc(h, (void**)&b.lvl_kp, sizeof(uint32_t) * (size_t)kpbase * B))) return rc;
    if ((rc = dev_alloc(h, (void**)&b.lvl_n, sizeof(int) * (size_t)g.nlevels * B))) return rc;
    b.sort_scratch_stride = sort_elems;
    if ((rc = dev_alloc(h, (void**)&b.sort_scratch, sizeof(unsigned long long) * (size_t)sort_elems * B))) return rc;
    if ((rc = dev_alloc(h, (void**)&b.sort_off, sizeof(int) * g.nlevels))) return rc;
    CK(cudaMemcpy(b.sort_off, sort_off.data(), sizeof(int) * g.nlevels, cudaMemcpyHostToDevice));
    if ((rc = dev_alloc(h, (void**)&b.work, sizeof(uint4) * (size_t)g.out_cap * B))) return rc;
    if ((rc = dev_alloc(h, (void**)&b.kps, sizeof(orbx_keypoint) * (size_t)g.out_cap * h->slots))) return rc;
    if ((rc = dev_alloc(h, (void**)&b.desc, (size_t)32 * g.out_cap * h->slots))) return rc;
    if ((rc = dev_alloc(h, (void**)&b.n, sizeof(int) * h->slots))) return rc;
    if ((rc = dev_alloc(h, (void**)&b.mono, sizeof(int) * h->slots))) return rc;
    if ((rc = dev_alloc(h, (void**)&b.err, sizeof(unsigned)))) return rc;
    CK(cudaMemset(b.n, 0, sizeof(int) * h->slots));
    CK(cudaMemset(b.mono, 0, sizeof(int) * h->slots));
    CK(cudaMemset(b.err, 0, sizeof(unsigned)));
    orbx_octree_configure(g);
    orbx_fast_configure(g);
    CK(cudaGetLastError());
    return ORBX_OK;
}

extern "C" int orbx_extractor_create(const orbx_params* p, orbx_extractor** out)
{
    if (!p || !out || p->nlevels < 1 || p->nlevels > ORBX_MAX_LEVELS || p->nfeatures < 1 || p->max_batch < 1 ||
        p->scale_factor <= 1.0f || p->max_width < 8 || p->max_height < 8) {
        orbx_set_error("%s%s", "orbx_extractor_create: invalid parameters", "");
        return ORBX_E_INVALID;
    }
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (p->device < 0 || p->device >= ndev) { orbx_set_error("%s%s", "no such CUDA device", ""); return ORBX_E_CUDA; }
    CK(cudaSetDevice(p->device));
    orbx_extractor* h = new orbx_extractor();
    h->p = *p;
    h->slots = p->max_batch + 1;
    h->scaleFactor = (double)p->scale_factor;
    const int nl = p->nlevels;
    h->scale.resize(nl); h->invScale.resize(nl); h->sigma2.resize(nl); h->invSigma2.resize(nl); h->featuresPerLevel.resize(nl);
    // R/src/ORBextractor.cc:413-444
    h->scale[0] = 1.0f; h->sigma2[0] = 1.0f;
    for (int i = 1; i < nl; i++) {
        h->scale[i] = (float)(h->scale[i - 1] * h->scaleFactor);
        h->sigma2[i] = h->scale[i] * h->scale[i];
    }
    for (int i = 0; i < nl; i++) { h->invScale[i] = 1.0f / h->scale[i]; h->invSigma2[i] = 1.0f / h->sigma2[i]; }
    const float factor = (float)(1.0f / h->scaleFactor);
    float nDesired = p->nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nl));
    int sum = 0;
    for (int l = 0; l < nl - 1; l++) {
        h->featuresPerLevel[l] = cv_round_f(nDesired);
        sum += h->featuresPerLevel[l];
        nDesired *= factor;
    }
    h->featuresPerLevel[nl - 1] = p->nfeatures - sum > 0 ? p->nfeatures - sum : 0;
    // k_octree keeps a level's node table on chip: quota + 8 nodes must fit its ORBX_OCTREE_MAX_NODES entries
    for (int l = 0; l < nl; l++)
        if (h->featuresPerLevel[l] + 8 > ORBX_OCTREE_MAX_NODES) {
            orbx_set_error("%s%s", "orbx_extractor_create: nfeatures too large (a level quota exceeds the octree node table, 4088)", "");
            delete h;
            return ORBX_E_INVALID;
        }
    memset(&h->geom, 0, sizeof(h->geom));
    memset(&h->buf, 0, sizeof(h->buf));
    h->d_level0 = nullptr; h->last_level0 = nullptr; h->last_batch = 0;
    h->pf_buf[0] = h->pf_buf[1] = nullptr; h->pf_bytes = 0; h->pf_ev[0] = h->pf_ev[1] = nullptr; h->pf_count = 0; h->pf_next = 0;
    h->pf_read[0] = h->pf_read[1] = nullptr; h->pf_read_set[0] = h->pf_read_set[1] = false; h->geom_gen = 0;
    h->d_raw = nullptr; h->raw_bytes = 0;
    h->profile = false; h->ev_head = 0; h->ev_count = 0; h->stage_batches = 0;
    for (int i = 0; i < 4; i++) h->stage_ms[i] = 0;
    // value-initialised above: every pointer the destructor frees is null until it is allocated
#define CKH(call)                                                                         \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            orbx_set_error("%s failed: %s", #call, cudaGetErrorString(e_));               \
            orbx_extractor_destroy(h);                                                    \
            return ORBX_E_CUDA;                                                           \
        }                                                                                 \
    } while (0)
    CKH(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    const int cap = orbx_extractor_max_keypoints(h);
    CKH(cudaMallocHost((void**)&h->h_stage_in, (size_t)((p->max_width + 63) & ~63) * p->max_height * p->max_batch));
    CKH(cudaMallocHost((void**)&h->h_kps, sizeof(orbx_keypoint) * (size_t)cap * h->slots));
    CKH(cudaMallocHost((void**)&h->h_desc, (size_t)32 * cap * h->slots));
    CKH(cudaMallocHost((void**)&h->h_n, sizeof(int) * h->slots));
    CKH(cudaMallocHost((void**)&h->h_mono, sizeof(int) * h->slots));
    CKH(cudaMallocHost((void**)&h->h_err, sizeof(unsigned)));
    orbx_upload_pattern();
    CKH(cudaGetLastError());
#undef CKH
    *out = h;
    return ORBX_OK;
}

extern "C" void orbx_extractor_destroy(orbx_extractor* h)
{
    if (!h) return;
    cudaSetDevice(h->p.device);
    cudaStreamSynchronize(h->stream);
    for (void* p : h->allocs) cudaFree(p);
    if (h->d_raw) cudaFree(h->d_raw);
    for (int i = 0; i < 2; i++) { if (h->pf_buf[i]) cudaFree(h->pf_buf[i]); if (h->pf_ev[i]) cudaEventDestroy(h->pf_ev[i]); if (h->pf_read[i]) cudaEventDestroy(h->pf_read[i]); }
    cudaFreeHost(h->h_stage_in); cudaFreeHost(h->h_kps); cudaFreeHost(h->h_desc);
    cudaFreeHost(h->h_n); cudaFreeHost(h->h_mono); cudaFreeHost(h->h_err);
    for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
    if (h->xg.exec) cudaGraphExecDestroy(h->xg.exec);
    if (h->xg.graph) cudaGraphDestroy(h->xg.graph);
    if (h->h_mailx) cudaFreeHost(h->h_mailx);
    if (h->h_pyr_stage) cudaFreeHost(h->h_pyr_stage);
    if (h->dag_init) for (int i = 0; i < 2 * ORBX_MAX_LEVELS; i++) { cudaStreamDestroy(h->br[i]); cudaEventDestroy(h->ev_br[i]); if (i < ORBX_MAX_LEVELS) cudaEventDestroy(h->ev_lvl[i]); }
    cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" int orbx_extractor_tables(const orbx_extractor* h, float* scale, float* inv_scale, float* sigma2,
                                     float* inv_sigma2, int32_t* fpl)
{
    if (!h) return ORBX_E_INVALID;
    for (int i = 0; i < h->p.nlevels; i++) {
        if (scale) scale[i] = h->scale[i];
        if (inv_scale) inv_scale[i] = h->invScale[i];
        if (sigma2) sigma2[i] = h->sigma2[i];
        if (inv_sigma2) inv_sigma2[i] = h->invSigma2[i];
        if (fpl) fpl[i] = h->featuresPerLevel[i];
    }
    return ORBX_OK;
}

extern "C" int orbx_extractor_max_keypoints(const orbx_extractor* h)
{
    // every level can overshoot its quota by up to 3 (a split adds at most 3 nodes), and a level whose
    // quota is tiny still starts from up to 4 nodes per root
    int cap = 0;
    for (int l = 0; l < h->p.nlevels; l++) cap += h->featuresPerLevel[l] + 4;
    return cap + 64;
}

static int deferred_error(orbx_extractor* h)
{
    const unsigned e = *h->h_err;
    if (e) {
        char buf[64]; snprintf(buf, sizeof(buf), "0x%x", e);
        orbx_set_error("device capacity error flags %s%s", buf, " (raise max_candidates_per_level / capacities)");
        cudaMemsetAsync(h->buf.err, 0, sizeof(unsigned), h->stream);
        return ORBX_E_CAPACITY;
    }
    return ORBX_OK;
}

// OrbxBuffers view whose per-frame arrays start at frame `off` (chunked pipelines reuse the same kernels)
static OrbxBuffers shifted(const orbx_extractor* h, int off)
{
    OrbxBuffers b = h->buf;
    if (off == 0) return b;
    const OrbxGeom& g = h->geom;
    for (int l = 0; l < g.nlevels; l++) {
        if (b.pyr[l]) b.pyr[l] += (long long)off * g.lv[l].frame_stride;
        b.blur[l] += (long long)off * g.lv[l].frame_stride;
    }
    b.row_cand += (long long)off * b.row_cand_stride;
    b.row_count += (long long)off * g.total_rows;
    b.lvl_kp += (long long)off * g.kp_total_cap;
    b.lvl_n += (long long)off * g.nlevels;
    b.sort_scratch += (long long)off * b.sort_scratch_stride;
    b.work += (long long)off * g.out_cap;
    return b;
}

// The few-frame launch order.  With one or two frames every kernel is a handful of CTAs and the step is a CHAIN of latencies:
// 8 pyramid levels, then FAST, then the octree whose level-0 CTA alone runs as long as the whole pyramid chain.  But level l's
// FAST and octree need level l only, so they go to a branch stream of their own as soon as that level exists: FAST + octree of
// level 0 run beside the resize chain 1 .. 7, and so on; the level-0 blur (needed by the descriptors only) leaves the chain too.
// Everything joins before k_finalize.  Under stream capture the events become graph edges, so a replay costs one launch; issued
// directly the ~60 API calls would cost more host time than the chain saves, which is why run_batch takes this path only when
// the stream is being captured (the single-frame graph of orbx_extract_match_batch) or ORBX_DAG=1 forces it.
static int ensure_dag(orbx_extractor* h)
{
    if (h->dag_init) return ORBX_OK;
    for (int i = 0; i < 2 * ORBX_MAX_LEVELS; i++) {
        CK(cudaStreamCreateWithFlags(&h->br[i], cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&h->ev_br[i], cudaEventDisableTiming));
        if (i < ORBX_MAX_LEVELS) CK(cudaEventCreateWithFlags(&h->ev_lvl[i], cudaEventDisableTiming));
    }
    h->dag_init = true;
    return ORBX_OK;
}

static int run_batch_dag(orbx_extractor* h, const OrbxBuffers& buf, const uint8_t* l0, int pitch0, long long stride0, int batch,
                         int lap0, int lap1, int first_slot, cudaStream_t s)
{
    const OrbxGeom& g = h->geom;
    const int nl = g.nlevels;
    CK(cudaEventRecord(h->ev_lvl[0], s));                                   // level 0 = the input: whatever `s` carries so far
    for (int l = 0; l < nl; l++) {
        if (l >= 1) {
            orbx_launch_pyramid(g, buf, l0, pitch0, stride0, batch, s, l, l + 1, ORBX_PYR_RESIZE);    // the chain: level l from level l - 1
            CK(cudaEventRecord(h->ev_lvl[l], s));
        }
        cudaStream_t sf = h->br[l], sb = h->br[ORBX_MAX_LEVELS + l];
        CK(cudaStreamWaitEvent(sf, h->ev_lvl[l], 0));
        orbx_launch_fast(g, buf, l0, pitch0, stride0, batch, sf, l, l + 1);
        orbx_launch_octree(g, buf, batch, sf, l, l + 1);
        CK(cudaEventRecord(h->ev_br[l], sf));
        CK(cudaStreamWaitEvent(sb, h->ev_lvl[l], 0));
        orbx_launch_pyramid(g, buf, l0, pitch0, stride0, batch, sb, l, l + 1, ORBX_PYR_BLUR);         // needed by the descriptors only
        CK(cudaEventRecord(h->ev_br[ORBX_MAX_LEVELS + l], sb));
    }
    for (int l = 0; l < nl; l++) { CK(cudaStreamWaitEvent(s, h->ev_br[l], 0)); CK(cudaStreamWaitEvent(s, h->ev_br[ORBX_MAX_LEVELS + l], 0)); }
    orbx_launch_describe(g, buf, l0, pitch0, stride0, batch, lap0, lap1, first_slot, s);
    return ORBX_OK;
}

static int run_batch(orbx_extractor* h, const uint8_t* d_level0, int pitch0, long long stride0, int batch,
                     int lap0, int lap1, int first_slot, cudaStream_t s, int frame_off = 0)
{
    const OrbxGeom& g = h->geom;
    const OrbxBuffers buf = shifted(h, frame_off);
    const uint8_t* l0 = d_level0 + (long long)frame_off * stride0;
    const bool prof = h->profile;
    OrbxPdlScope pdl_scope(batch <= 2);
    static const int dag_env = getenv("ORBX_DAG") ? atoi(getenv("ORBX_DAG")) : -1;      // 0: never, 1: always for few frames, unset: under capture
    if (!prof && batch <= 2 && g.nlevels > 1 && dag_env != 0) {
        int rc = ensure_dag(h);                                             // (created on a direct call: not inside a capture)
        if (rc) return rc;
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (dag_env != 1) CK(cudaStreamIsCapturing(s, &cs));
        if (dag_env == 1 || cs == cudaStreamCaptureStatusActive) {
            rc = run_batch_dag(h, buf, l0, pitch0, stride0, batch, lap0, lap1, first_slot, s);
            if (rc) return rc;
            h->last_level0 = d_level0; h->last_pitch0 = pitch0; h->last_stride0 = stride0;
            if (frame_off + batch > h->last_batch || frame_off == 0) h->last_batch = frame_off + batch;
            CK(cudaGetLastError());
            return ORBX_OK;
        }
    }
    cudaEvent_t* e = nullptr;
    if (prof) {
        if (h->ev_count == ORBX_EV_SETS) harvest_stage_times(h, false);
        e = h->ev.data() + h->ev_head * 5;
        h->ev_head = (h->ev_head + 1) % ORBX_EV_SETS; h->ev_count++;
        cudaEventRecord(e[0], s);
    }
    orbx_launch_pyramid(g, buf, l0, pitch0, stride0, batch, s);
    if (prof) cudaEventRecord(e[1], s);
    orbx_launch_fast(g, buf, l0, pitch0, stride0, batch, s);
    if (prof) cudaEventRecord(e[2], s);
    orbx_launch_octree(g, buf, batch, s);
    if (prof) cudaEventRecord(e[3], s);
    orbx_launch_describe(g, buf, l0, pitch0, stride0, batch, lap0, lap1, first_slot, s);
    if (prof) cudaEventRecord(e[4], s);
    h->last_level0 = d_level0; h->last_pitch0 = pitch0; h->last_stride0 = stride0;
    if (frame_off + batch > h->last_batch || frame_off == 0) h->last_batch = frame_off + batch;
    CK(cudaGetLastError());
    return ORBX_OK;
}

extern "C" int orbx_extract_batch_device(orbx_extractor* h, const uint8_t* d_imgs, int batch, int width, int height,
                                         int stride, size_t frame_stride, int lap0, int lap1, int first_slot, void* stream)
{
    if (!h || !d_imgs || batch < 1 || batch > h->p.max_batch || first_slot < 0 || first_slot + batch > h->slots ||
        stride < width) {
        orbx_set_error("%s%s", "orbx_extract_batch_device: invalid arguments", "");
        return ORBX_E_INVALID;
    }
    if (width <= 0 || height <= 0) return ORBX_E_EMPTY;
    CK(cudaSetDevice(h->p.device));
    int rc = configure_geometry(h, width, height);
    if (rc) return rc;
    cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
    return run_batch(h, d_imgs, stride, (long long)frame_stride, batch, lap0, lap1, first_slot, s);
}

extern "C" int orbx_extractor_sync(orbx_extractor* h, void* stream)
{
    if (!h) return ORBX_E_INVALID;
    cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
    if (h->buf.err) CK(cudaMemcpyAsync(h->h_err, h->buf.err, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
    else *h->h_err = 0;
    CK(cudaStreamSynchronize(s));
    return deferred_error(h);
}

// one launch instead of four device-to-device copies: keypoints (7 words each), descriptors (8 words each), n, monoIndex
__global__ void k_copy_slot(uint32_t* kps, uint32_t* desc, int* n, int* mono, int from, int to, int cap)
{
    static_assert(sizeof(orbx_keypoint) == 28, "keypoint record = 7 words");
    const int wk = cap * 7, wd = cap * 8;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < wk + wd; i += gridDim.x * blockDim.x) {
        if (i < wk) kps[(size_t)to * wk + i] = kps[(size_t)from * wk + i];
        else desc[(size_t)to * wd + (i - wk)] = desc[(size_t)from * wd + (i - wk)];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { n[to] = n[from]; mono[to] = mono[from]; }
}

extern "C" int orbx_extractor_copy_slot(orbx_extractor* h, int from, int to, void* stream)
{
    if (!h || from < 0 || to < 0 || from >= h->slots || to >= h->slots || !h->buf.kps) return ORBX_E_INVALID;
    if (from == to) return ORBX_OK;
    cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
    const int cap = h->geom.out_cap;
    k_copy_slot<<<(cap * 15 + 1023) / 1024, 256, 0, s>>>(reinterpret_cast<uint32_t*>(h->buf.kps), reinterpret_cast<uint32_t*>(h->buf.desc),
                                                         h->buf.n, h->buf.mono, from, to, cap);
    ORBX_COUNT_LAUNCH(1);
    CK(cudaGetLastError());
    return ORBX_OK;
}

extern "C" int orbx_extractor_results_device(orbx_extractor* h, orbx_keypoint** d_kps, uint8_t** d_desc,
                                             int32_t** d_n, int32_t** d_mono, int* cap, int* slots)
{
    if (!h || !h->buf.kps) { orbx_set_error("%s%s", "no batch has been extracted yet", ""); return ORBX_E_INVALID; }
    if (d_kps) *d_kps = h->buf.kps;
    if (d_desc) *d_desc = h->buf.desc;
    if (d_n) *d_n = h->buf.n;
    if (d_mono) *d_mono = h->buf.mono;
    if (cap) *cap = h->geom.out_cap;
    if (slots) *slots = h->slots;
    return ORBX_OK;
}

static bool is_pinned(const void* p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

// ---- host-buffer pipeline pieces (shared with orbx_extract_match_batch in orbx_match.cu) ----
int orbx_ex_configure(orbx_extractor* h, int width, int height)
{
    CK(cudaSetDevice(h->p.device));
    return configure_geometry(h, width, height);
}

// H2D of frames [f0, f0+count) of a host batch into the staged level 0 (geometry already configured)
int orbx_ex_stage_input(orbx_extractor* h, const uint8_t* imgs, int f0, int count, int width, int height, int stride,
                        size_t frame_stride, cudaStream_t s)
{
    const int p0 = h->pitch0;
    uint8_t* dst = h->d_level0 + (size_t)f0 * h->stride0;
    const uint8_t* src = imgs + (size_t)f0 * frame_stride;
    if (is_pinned(imgs)) {
        // pinned caller memory: DMA straight from it
        if (frame_stride == (size_t)stride * height && stride == p0) {
            CK(cudaMemcpyAsync(dst, src, (size_t)h->stride0 * count, cudaMemcpyHostToDevice, s));
        } else if (frame_stride == (size_t)stride * height) {
            // rows at another pitch (e.g. a 1241-px-wide KITTI frame): one contiguous DMA into a device landing zone, then
            // a device-side 2-D copy to the staging pitch (a strided host->device DMA runs at a fraction of the link rate)
            const size_t need = frame_stride * (size_t)h->p.max_batch;
            if (need > h->raw_bytes) {
                if (h->d_raw) { CK(cudaStreamSynchronize(s)); cudaFree(h->d_raw); h->d_raw = nullptr; h->raw_bytes = 0; }
                CK(cudaMalloc((void**)&h->d_raw, need)); h->raw_bytes = need;
            }
            uint8_t* raw = h->d_raw + (size_t)f0 * frame_stride;
            CK(cudaMemcpyAsync(raw, src, frame_stride * count, cudaMemcpyHostToDevice, s));
            CK(cudaMemcpy2DAsync(dst, p0, raw, stride, width, (size_t)height * count, cudaMemcpyDeviceToDevice, s));
        } else {
            for (int f = 0; f < count; f++)
                CK(cudaMemcpy2DAsync(dst + (size_t)f * h->stride0, p0, src + f * frame_stride, stride, width, height,
                                     cudaMemcpyHostToDevice, s));
        }
    } else {
        uint8_t* stage = h->h_stage_in + (size_t)f0 * h->stride0;
        for (int f = 0; f < count; f++)
            for (int y = 0; y < height; y++)
                memcpy(stage + (size_t)f * h->stride0 + (size_t)y * p0, src + f * frame_stride + (size_t)y * stride, width);
        CK(cudaMemcpyAsync(dst, stage, (size_t)h->stride0 * count, cudaMemcpyHostToDevice, s));
    }
    return ORBX_OK;
}

// ---- input prefetch: the frames of the NEXT host-buffer call travel to the device while the current call computes ----
// copies `batch` pinned, tightly packed frames into one of two extra staging buffers on stream `s_copy`
int orbx_ex_prefetch(orbx_extractor* h, const uint8_t* imgs, int batch, int width, int height, int stride, size_t frame_stride, cudaStream_t s_copy)
{
    int rc = orbx_ex_configure(h, width, height);
    if (rc) return rc;
    if (batch < 1 || batch > h->p.max_batch || !is_pinned(imgs) || frame_stride != (size_t)stride * height || stride != h->pitch0) {
        orbx_set_error("%s%s", "orbx_extract_match_batch_prefetch: needs pinned frames packed at the staging pitch (width rounded up to 16)", "");
        return ORBX_E_INVALID;
    }
    if (h->pf_count == 2) { orbx_set_error("%s%s", "orbx_extract_match_batch_prefetch: two batches are already waiting", ""); return ORBX_E_CAPACITY; }
    const size_t need = (size_t)h->stride0 * h->p.max_batch;
    if (need > h->pf_bytes) {
        if (h->pf_count) { orbx_set_error("%s%s", "orbx_extract_match_batch_prefetch: image size changed with a batch waiting", ""); return ORBX_E_INVALID; }
        for (int i = 0; i < 2; i++) { if (h->pf_buf[i]) { CK(cudaDeviceSynchronize()); cudaFree(h->pf_buf[i]); h->pf_buf[i] = nullptr; } }
        for (int i = 0; i < 2; i++) {
            CK(cudaMalloc((void**)&h->pf_buf[i], need));
            if (!h->pf_ev[i]) CK(cudaEventCreateWithFlags(&h->pf_ev[i], cudaEventDisableTiming));
            if (!h->pf_read[i]) CK(cudaEventCreateWithFlags(&h->pf_read[i], cudaEventDisableTiming));
            h->pf_read_set[i] = false;
        }
        h->pf_bytes = need;
    }
    const int slot = h->pf_next; h->pf_next ^= 1;
    orbx_extractor::Prefetch& e = h->pf[h->pf_count++];
    e.src = imgs; e.batch = batch; e.width = width; e.height = height; e.buf = h->pf_buf[slot]; e.ev = h->pf_ev[slot];
    // the buffer may still be read by the kernels of the batch that used it last (streaming form: two batches in flight)
    if (h->pf_read_set[slot]) CK(cudaStreamWaitEvent(s_copy, h->pf_read[slot], 0));
    CK(cudaMemcpyAsync(e.buf, imgs, (size_t)h->stride0 * batch, cudaMemcpyHostToDevice, s_copy));
    CK(cudaEventRecord(e.ev, s_copy));
    return ORBX_OK;
}

// the oldest waiting batch if it is exactly this call's input (then *d_frames / *ready describe it and it is consumed);
// any other waiting batch is stale (the caller changed its mind): dropped after its copy finished
bool orbx_ex_take_prefetched(orbx_extractor* h, const uint8_t* imgs, int batch, int width, int height, const uint8_t** d_frames, cudaEvent_t* ready)
{
    if (h->pf_count == 0) return false;
    const orbx_extractor::Prefetch e = h->pf[0];
    if (e.src == imgs && e.batch == batch && e.width == width && e.height == height) {
        h->pf[0] = h->pf[1]; h->pf_count--;
        *d_frames = e.buf; *ready = e.ev;
        return true;
    }
    for (int i = 0; i < h->pf_count; i++) cudaEventSynchronize(h->pf[i].ev);
    h->pf_count = 0;
    return false;
}
// to be called after the kernels that read a prefetched buffer were queued on `s`
int orbx_ex_prefetch_mark_read(orbx_extractor* h, const uint8_t* d_frames, cudaStream_t s)
{
    for (int i = 0; i < 2; i++)
        if (h->pf_buf[i] == d_frames && h->pf_read[i]) { CK(cudaEventRecord(h->pf_read[i], s)); h->pf_read_set[i] = true; }
    return ORBX_OK;
}
int orbx_ex_geom_gen(orbx_extractor* h) { return h->geom_gen; }
int orbx_ex_pitch0(orbx_extractor* h) { return h->pitch0; }
unsigned* orbx_ex_err_device(orbx_extractor* h) { return h->buf.err; }
bool orbx_ex_profiling(orbx_extractor* h) { return h->profile; }
bool orbx_host_pinned(const void* p) { return is_pinned(p); }
long long orbx_ex_stride0(orbx_extractor* h) { return h->stride0; }

int orbx_ex_run_staged(orbx_extractor* h, int f0, int count, int lap0, int lap1, int first_slot, cudaStream_t s)
{
    return run_batch(h, h->d_level0, h->pitch0, h->stride0, count, lap0, lap1, first_slot, s, f0);
}

// frames [f0, f0+count) of a DEVICE batch (geometry already configured); per-frame scratch slice f0..
int orbx_ex_run_device(orbx_extractor* h, const uint8_t* d_imgs, int pitch, long long fstride, int f0, int count,
                       int lap0, int lap1, int first_slot, cudaStream_t s)
{
    return run_batch(h, d_imgs, pitch, fstride, count, lap0, lap1, first_slot, s, f0);
}

cudaStream_t orbx_ex_stream(orbx_extractor* h) { return h->stream; }
int orbx_ex_device(orbx_extractor* h) { return h->p.device; }
int orbx_ex_max_batch(orbx_extractor* h) { return h->p.max_batch; }

int orbx_ex_pyramid_view(orbx_extractor* h, int frame, OrbxPyrView* out)
{
    if (!h || h->geom.width == 0 || frame < 0 || frame >= h->last_batch || !h->last_level0) return ORBX_E_INVALID;
    const OrbxGeom& g = h->geom;
    out->nlevels = g.nlevels;
    for (int l = 0; l < g.nlevels; l++) {
        if (l == 0) { out->lv[0] = h->last_level0 + (long long)frame * h->last_stride0; out->pitch[0] = h->last_pitch0; out->fstride[0] = h->last_stride0; }
        else { out->lv[l] = h->buf.pyr[l] + (long long)frame * g.lv[l].frame_stride; out->pitch[l] = g.lv[l].pitch; out->fstride[l] = g.lv[l].frame_stride; }
        out->w[l] = g.lv[l].w; out->h[l] = g.lv[l].h;
        out->scale[l] = h->scale[l]; out->inv_scale[l] = h->invScale[l];
    }
    return ORBX_OK;
}
int orbx_ex_out_cap(orbx_extractor* h) { return h->geom.out_cap; }

// issues the D2H copies of `count` result slots starting at first_slot into host frame positions host_off..;
// direct into the caller's buffers when they are pinned and laid out with the handle's own capacity, else into the
// handle's pinned staging (unpacked by _finish)
int orbx_ex_fetch_async(orbx_extractor* h, int first_slot, int count, int host_off, orbx_keypoint* kps, uint8_t* desc, int cap,
                        int32_t* n, int32_t* mono_index, cudaStream_t s, bool direct, bool counts)
{
    const size_t ocap = h->geom.out_cap;
    const size_t ho = (size_t)host_off;
    if (counts) {
        CK(cudaMemcpyAsync(direct ? n + ho : h->h_n + ho, h->buf.n + first_slot, sizeof(int) * count, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(direct ? mono_index + ho : h->h_mono + ho, h->buf.mono + first_slot, sizeof(int) * count, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaMemcpyAsync(direct ? kps + ho * ocap : h->h_kps + ho * ocap, h->buf.kps + first_slot * ocap, sizeof(orbx_keypoint) * ocap * count, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(direct ? desc + ho * ocap * 32 : h->h_desc + ho * ocap * 32, h->buf.desc + first_slot * ocap * 32, ocap * 32 * count, cudaMemcpyDeviceToHost, s));
    return ORBX_OK;
}

bool orbx_ex_can_fetch_direct(orbx_extractor* h, orbx_keypoint* kps, uint8_t* desc, int cap, int32_t* n, int32_t* mono_index)
{
    return kps && desc && n && mono_index && cap == h->geom.out_cap && is_pinned(kps) && is_pinned(desc) && is_pinned(n) && is_pinned(mono_index);
}

// after the stream(s) were synchronised: device error flags, then unpack the staging if the fetch was not direct
// queues the D2H copy of the device error flags on `s` (the caller synchronises s, then calls orbx_ex_fetch_finish with err_fetched)
// counts and error flags that reached the host another way (the matcher's mailbox): what the copies of orbx_ex_fetch_async /
// orbx_ex_fetch_err_async would have delivered
void orbx_ex_set_fetched(orbx_extractor* h, unsigned err, const int32_t* n, const int32_t* mono, int count, int32_t* user_n, int32_t* user_mono, bool direct)
{
    *h->h_err = err;
    memcpy(direct ? user_n : h->h_n, n, sizeof(int32_t) * count);
    memcpy(direct ? user_mono : h->h_mono, mono, sizeof(int32_t) * count);
}

int orbx_ex_fetch_err_async(orbx_extractor* h, cudaStream_t s)
{
    CK(cudaMemcpyAsync(h->h_err, h->buf.err, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
    return ORBX_OK;
}

int orbx_ex_fetch_finish(orbx_extractor* h, int count, orbx_keypoint* kps, uint8_t* desc, int cap,
                         int32_t* n, int32_t* mono_index, bool direct, bool err_fetched)
{
    if (!err_fetched) CK(cudaMemcpy(h->h_err, h->buf.err, sizeof(unsigned), cudaMemcpyDeviceToHost));
    int rc = deferred_error(h);
    if (rc) return rc;
    if (direct) return ORBX_OK;
    const size_t ocap = h->geom.out_cap;
    for (int i = 0; i < count; i++) {
        const int ni = h->h_n[i];
        if (n) n[i] = ni;
        if (mono_index) mono_index[i] = h->h_mono[i];
        if (ni > cap) { orbx_set_error("%s%s", "caller keypoint capacity too small", ""); return ORBX_E_CAPACITY; }
        if (kps) memcpy(kps + (size_t)i * cap, h->h_kps + (size_t)i * ocap, sizeof(orbx_keypoint) * ni);
        if (desc) memcpy(desc + (size_t)i * cap * 32, h->h_desc + (size_t)i * ocap * 32, (size_t)32 * ni);
    }
    return ORBX_OK;
}

extern "C" int orbx_extractor_download(orbx_extractor* h, int first_slot, int count, orbx_keypoint* kps, uint8_t* desc,
                                       int cap, int32_t* n, int32_t* mono_index, void* stream)
{
    if (!h || !h->buf.kps || first_slot < 0 || count < 1 || first_slot + count > h->slots) return ORBX_E_INVALID;
    CK(cudaSetDevice(h->p.device));
    cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
    const bool direct = orbx_ex_can_fetch_direct(h, kps, desc, cap, n, mono_index);
    int rc = orbx_ex_fetch_async(h, first_slot, count, 0, kps, desc, cap, n, mono_index, s, direct);
    if (rc) return rc;
    CK(cudaStreamSynchronize(s));
    return orbx_ex_fetch_finish(h, count, kps, desc, cap, n, mono_index, direct, false);
}

// n, monoIndex and the error word of slots 0 .. batch-1 into mapped pinned memory: one kernel instead of three copies
__global__ void k_ex_mailbox(const int* n, const int* mono, const unsigned* err, int batch, int* mail)
{
    orbx_pdl_prologue();
    const int i = threadIdx.x;
    if (i == 0) mail[0] = (int)*err;
    if (i < batch) { mail[1 + i] = n[i]; mail[1 + batch + i] = mono[i]; }
}

// One or two frames through the synchronous call (ORBextractor::operator() of the class API): the kernels replay from a launch
// graph captured in the level-parallel order of run_batch_dag (the first call with a set of arguments runs directly, the second
// captures), the counts come back through k_ex_mailbox: H2D | graph | 2 copies | one synchronisation.
static int extract_small(orbx_extractor* h, int batch, int lap0, int lap1, orbx_keypoint* kps, uint8_t* desc, int cap, int32_t* n, int32_t* mono_index)
{
    cudaStream_t s = h->stream;
    if (!h->h_mailx) {
        CK(cudaHostAlloc((void**)&h->h_mailx, sizeof(int32_t) * 8, cudaHostAllocMapped));
        CK(cudaHostGetDevicePointer((void**)&h->d_mailx, h->h_mailx, 0));
    }
    auto issue = [&]() -> int {
        int r = run_batch(h, h->d_level0, h->pitch0, h->stride0, batch, lap0, lap1, 0, s, 0);
        if (r) return r;
        orbx_launch_pdl(k_ex_mailbox, dim3(1), dim3(32), 0, s, (const int*)h->buf.n, (const int*)h->buf.mono, (const unsigned*)h->buf.err, batch, h->d_mailx);
        ORBX_COUNT_LAUNCH(1);
        CK(cudaGetLastError());
        return ORBX_OK;
    };
    static const bool no_graph = getenv("ORBX_NO_GRAPH") != nullptr;
    orbx_extractor::SmallGraph& G = h->xg;
    const bool same = G.seen > 0 && G.batch == batch && G.lap0 == lap0 && G.lap1 == lap1 && G.gen == h->geom_gen;
    int rc = ORBX_OK;
    if (no_graph || h->profile) rc = issue();
    else if (same && G.seen == 2 && G.exec) { CK(cudaGraphLaunch(G.exec, s)); ORBX_COUNT_LAUNCH(G.nkernels); }
    else if (same && G.seen == 1) {
        const unsigned long long before = g_orbx_launches.load(std::memory_order_relaxed);
        CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        rc = issue();
        cudaGraph_t graph = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(s, &graph);
        const int captured = (int)(g_orbx_launches.load(std::memory_order_relaxed) - before);
        g_orbx_launches.fetch_sub(captured, std::memory_order_relaxed);                // nothing ran yet
        if (!rc && ce == cudaSuccess && graph && cudaGraphInstantiate(&G.exec, graph, 0) == cudaSuccess) {
            G.graph = graph; G.nkernels = captured; G.seen = 2;
            CK(cudaGraphLaunch(G.exec, s)); ORBX_COUNT_LAUNCH(captured);
        } else {
            cudaGetLastError();
            if (graph) cudaGraphDestroy(graph);
            G.exec = nullptr; G.seen = -1;                                             // not capturable here: stay on direct launches
            if (!rc) rc = issue();
        }
    } else {
        if (!same) {
            if (G.exec) { cudaGraphExecDestroy(G.exec); G.exec = nullptr; }
            if (G.graph) { cudaGraphDestroy(G.graph); G.graph = nullptr; }
            G.batch = batch; G.lap0 = lap0; G.lap1 = lap1; G.gen = h->geom_gen; G.seen = 1;
        }
        rc = issue();
    }
    if (rc) return rc;
    const bool direct = orbx_ex_can_fetch_direct(h, kps, desc, cap, n, mono_index);
    rc = orbx_ex_fetch_async(h, 0, batch, 0, kps, desc, cap, n, mono_index, s, direct, false);
    if (rc) return rc;
    CK(cudaStreamSynchronize(s));
    orbx_ex_set_fetched(h, (unsigned)h->h_mailx[0], h->h_mailx + 1, h->h_mailx + 1 + batch, batch, n, mono_index, direct);
    return orbx_ex_fetch_finish(h, batch, kps, desc, cap, n, mono_index, direct, true);
}

extern "C" int orbx_extract_batch(orbx_extractor* h, const uint8_t* imgs, int batch, int width, int height,
                                  int stride, size_t frame_stride, int lap0, int lap1,
                                  orbx_keypoint* kps, uint8_t* desc, int cap, int32_t* n, int32_t* mono_index)
{
    if (!h || batch < 1 || batch > h->p.max_batch) { orbx_set_error("%s%s", "orbx_extract_batch: invalid arguments", ""); return ORBX_E_INVALID; }
    if (!imgs || width <= 0 || height <= 0) return ORBX_E_EMPTY;
    if (stride < width) return ORBX_E_INVALID;
    int rc = orbx_ex_configure(h, width, height);
    if (rc) return rc;
    rc = orbx_ex_stage_input(h, imgs, 0, batch, width, height, stride, frame_stride, h->stream);
    if (rc) return rc;
    if (batch <= 2) return extract_small(h, batch, lap0, lap1, kps, desc, cap, n, mono_index);
    rc = orbx_ex_run_staged(h, 0, batch, lap0, lap1, 0, h->stream);
    if (rc) return rc;
    return orbx_extractor_download(h, 0, batch, kps, desc, cap, n, mono_index, h->stream);
}

extern "C" int orbx_extract(orbx_extractor* h, const uint8_t* img, int width, int height, int stride,
                            int lap0, int lap1, orbx_keypoint* kps, uint8_t* desc, int cap, int* n, int* mono_index)
{
    int32_t nn = 0, mm = 0;
    int rc = orbx_extract_batch(h, img, 1, width, height, stride, (size_t)stride * (height > 0 ? height : 0), lap0, lap1,
                                kps, desc, cap, &nn, &mm);
    if (n) *n = nn;
    if (mono_index) *mono_index = mm;
    return rc;
}

extern "C" int orbx_extractor_profile(orbx_extractor* h, int enable, double* stage_ms4, int* batches)
{
    if (!h) return ORBX_E_INVALID;
    harvest_stage_times(h, true);
    if (stage_ms4) for (int i = 0; i < 4; i++) stage_ms4[i] = h->stage_ms[i];
    if (batches) *batches = h->stage_batches;
    if (enable >= 0) {
        h->profile = enable != 0;
        if (h->profile && h->ev.empty()) {
            h->ev.resize(ORBX_EV_SETS * 5);
            for (cudaEvent_t& e : h->ev) CK(cudaEventCreate(&e));
        }
        for (int i = 0; i < 4; i++) h->stage_ms[i] = 0;
        h->stage_batches = 0; h->ev_head = 0; h->ev_count = 0;
    }
    return ORBX_OK;
}

extern "C" int orbx_pyramid_level_size(const orbx_extractor* h, int level, int* width, int* height)
{
    if (!h || h->geom.width == 0 || level < 0 || level >= h->geom.nlevels) return ORBX_E_INVALID;
    if (width) *width = h->geom.lv[level].w;
    if (height) *height = h->geom.lv[level].h;
    return ORBX_OK;
}

static int level_to_host(orbx_extractor* h, const uint8_t* base, int pitch, long long fstride, int slot, int level,
                         uint8_t* dst, int dst_stride)
{
    const OrbxLevel& L = h->geom.lv[level];
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy2D(dst, dst_stride, base + (long long)slot * fstride, pitch, L.w, L.h, cudaMemcpyDeviceToHost));
    return ORBX_OK;
}

extern "C" int orbx_pyramid_to_host(orbx_extractor* h, int slot, int level, uint8_t* dst, int dst_stride)
{
    if (!h || h->geom.width == 0 || level < 0 || level >= h->geom.nlevels || slot < 0 || slot >= h->last_batch || !dst)
        return ORBX_E_INVALID;
    CK(cudaSetDevice(h->p.device));
    if (level == 0) return level_to_host(h, h->last_level0, h->last_pitch0, h->last_stride0, slot, 0, dst, dst_stride);
    return level_to_host(h, h->buf.pyr[level], h->geom.lv[level].pitch, h->geom.lv[level].frame_stride, slot, level, dst, dst_stride);
}

// levels [first_level, first_level + n_levels) of one frame into the handle's pinned staging (tight rows), one synchronisation
static int stage_levels(orbx_extractor* h, int slot, int first_level, int n_levels)
{
    if (!h || h->geom.width == 0 || first_level < 0 || n_levels < 1 || first_level + n_levels > h->geom.nlevels || slot < 0 || slot >= h->last_batch)
        return ORBX_E_INVALID;
    CK(cudaSetDevice(h->p.device));
    const OrbxGeom& g = h->geom;
    size_t need = 0;
    for (int l = first_level; l < first_level + n_levels; l++) need += ((size_t)g.lv[l].w * g.lv[l].h + 63) & ~(size_t)63;
    if (need > h->pyr_stage_bytes) {
        if (h->h_pyr_stage) { CK(cudaStreamSynchronize(h->stream)); cudaFreeHost(h->h_pyr_stage); h->h_pyr_stage = nullptr; h->pyr_stage_bytes = 0; }
        CK(cudaMallocHost((void**)&h->h_pyr_stage, need));
        h->pyr_stage_bytes = need;
    }
    size_t off = 0;
    for (int l = first_level; l < first_level + n_levels; l++) {
        const OrbxLevel& L = g.lv[l];
        const uint8_t* src = l == 0 ? h->last_level0 + (long long)slot * h->last_stride0 : h->buf.pyr[l] + (long long)slot * L.frame_stride;
        const int pitch = l == 0 ? h->last_pitch0 : L.pitch;
        CK(cudaMemcpy2DAsync(h->h_pyr_stage + off, L.w, src, pitch, L.w, L.h, cudaMemcpyDeviceToHost, h->stream));
        off += ((size_t)L.w * L.h + 63) & ~(size_t)63;
    }
    CK(cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}

extern "C" int orbx_pyramid_levels_staged(orbx_extractor* h, int slot, int first_level, int n_levels, const uint8_t** ptr, int* stride)
{
    if (!ptr || !stride) return ORBX_E_INVALID;
    const int rc = stage_levels(h, slot, first_level, n_levels);
    if (rc) return rc;
    size_t off = 0;
    for (int l = first_level; l < first_level + n_levels; l++) {
        const OrbxLevel& L = h->geom.lv[l];
        ptr[l - first_level] = h->h_pyr_stage + off; stride[l - first_level] = L.w;
        off += ((size_t)L.w * L.h + 63) & ~(size_t)63;
    }
    return ORBX_OK;
}

extern "C" int orbx_pyramid_levels_to_host(orbx_extractor* h, int slot, int first_level, int n_levels, uint8_t* const* dst, const int* dst_stride)
{
    if (!h || h->geom.width == 0 || !dst || !dst_stride || first_level < 0 || n_levels < 1 || first_level + n_levels > h->geom.nlevels) return ORBX_E_INVALID;
    for (int l = first_level; l < first_level + n_levels; l++)
        if (!dst[l - first_level] || dst_stride[l - first_level] < h->geom.lv[l].w) return ORBX_E_INVALID;
    const int rc = stage_levels(h, slot, first_level, n_levels);
    if (rc) return rc;
    size_t off = 0;
    for (int l = first_level; l < first_level + n_levels; l++) {
        const OrbxLevel& L = h->geom.lv[l];
        uint8_t* d = dst[l - first_level]; const int ds = dst_stride[l - first_level];
        if (ds == L.w) memcpy(d, h->h_pyr_stage + off, (size_t)L.w * L.h);
        else for (int y = 0; y < L.h; y++) memcpy(d + (size_t)y * ds, h->h_pyr_stage + off + (size_t)y * L.w, L.w);
        off += ((size_t)L.w * L.h + 63) & ~(size_t)63;
    }
    return ORBX_OK;
}

extern "C" int orbx_blurred_to_host(orbx_extractor* h, int slot, int level, uint8_t* dst, int dst_stride)
{
    if (!h || h->geom.width == 0 || level < 0 || level >= h->geom.nlevels || slot < 0 || slot >= h->last_batch || !dst)
        return ORBX_E_INVALID;
    CK(cudaSetDevice(h->p.device));
    return level_to_host(h, h->buf.blur[level], h->geom.lv[level].pitch, h->geom.lv[level].frame_stride, slot, level, dst, dst_stride);
}

extern "C" int orbx_candidates_to_host(orbx_extractor* h, int slot, int level, float* xyr, int cap, int* n)
{
    if (!h || h->geom.width == 0 || level < 0 || level >= h->geom.nlevels || slot < 0 || slot >= h->last_batch) return ORBX_E_INVALID;
    CK(cudaSetDevice(h->p.device));
    CK(cudaStreamSynchronize(h->stream));
    const OrbxGeom& g = h->geom; const OrbxLevel& L = g.lv[level];
    const int nUnits = L.nRows * L.nSeg;
    std::vector<int> cnt(nUnits > 0 ? nUnits : 1), off(nUnits > 0 ? nUnits : 1);
    int total = 0;
    if (nUnits > 0) {
        CK(cudaMemcpy(cnt.data(), h->buf.row_count + (long long)slot * g.total_rows + L.row_base, sizeof(int) * nUnits, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(off.data(), h->buf.row_off + L.row_base, sizeof(int) * nUnits, cudaMemcpyDeviceToHost));
    }
    std::vector<uint32_t> tmp;
    for (int r = 0; r < nUnits; r++) {
        tmp.resize(cnt[r] > 0 ? cnt[r] : 1);
        if (cnt[r] > 0)
            CK(cudaMemcpy(tmp.data(), h->buf.row_cand + (long long)slot * h->buf.row_cand_stride + off[r], sizeof(uint32_t) * cnt[r], cudaMemcpyDeviceToHost));
        for (int k = 0; k < cnt[r]; k++, total++)
            if (total < cap) {
                xyr[total * 3] = (float)(tmp[k] & 0xFFF); xyr[total * 3 + 1] = (float)((tmp[k] >> 12) & 0xFFF);
                xyr[total * 3 + 2] = (float)(tmp[k] >> 24);
            }
    }
    if (n) *n = total;
    return total <= cap ? ORBX_OK : ORBX_E_CAPACITY;
}

extern "C" int orbx_level_keypoints_to_host(orbx_extractor* h, int slot, int level, float* xyr, int cap, int* n)
{
    if (!h || h->geom.width == 0 || level < 0 || level >= h->geom.nlevels || slot < 0 || slot >= h->last_batch) return ORBX_E_INVALID;
    CK(cudaSetDevice(h->p.device));
    CK(cudaStreamSynchronize(h->stream));
    const OrbxGeom& g = h->geom; const OrbxLevel& L = g.lv[level];
    int cnt = 0;
    CK(cudaMemcpy(&cnt, h->buf.lvl_n + (long long)slot * g.nlevels + level, sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<uint32_t> tmp(cnt > 0 ? cnt : 1);
    if (cnt > 0) CK(cudaMemcpy(tmp.data(), h->buf.lvl_kp + (long long)slot * g.kp_total_cap + L.kp_base, sizeof(uint32_t) * cnt, cudaMemcpyDeviceToHost));
    for (int k = 0; k < cnt && k < cap; k++) {
        xyr[k * 3] = (float)(tmp[k] & 0xFFF); xyr[k * 3 + 1] = (float)((tmp[k] >> 12) & 0xFFF); xyr[k * 3 + 2] = (float)(tmp[k] >> 24);
    }
    if (n) *n = cnt;
    return cnt <= cap ? ORBX_OK : ORBX_E_CAPACITY;
}
