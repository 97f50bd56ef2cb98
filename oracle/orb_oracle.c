/* orb_oracle.c - CPU oracle (TEST INFRASTRUCTURE, not product code).  See orb_oracle.h.
 *
 * Parity status: the reference holds no golden vectors for this path,
 * and its own sources cannot be compiled here (they need OpenCV/Eigen/ROS headers).  The
 * OpenCV primitives restated below are pinned against cv2 4.13.0 (the reference's third-party
 * dependency, whose version the reference leaves open) by tests/test_oracle_pin_cv2.py and
 * by the fixtures under tests/golden/ (made by oracle/gen_golden.py).
 *
 * Build with -ffp-contract=off: the reference is built without -march=native
 * (R/../CMakeLists.txt:15-19), so x86-64 emits separate mul/add and no FMA.
 *
 * R/ = /root/reference/src/orb_slam3_ros/orb_slam3/
 */
#include "orb_oracle.h"
#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

static const int32_t kPattern[1024] = {
#include "../multi_orbslam3_b200/csrc/orb_pattern.inc"
};

#define PATCH_SIZE 31
#define HALF_PATCH_SIZE 15
#define EDGE_THRESHOLD 19
#define TH_HIGH 100
#define TH_LOW 50
#define HISTO_LENGTH 30
#define GRID_COLS 64
#define GRID_ROWS 48

/* ------------------------------------------------------------------------------------------
 * OpenCV primitives
 * ---------------------------------------------------------------------------------------- */

/* cvRound(float): SSE cvtss2si, round-half-to-even in the default rounding mode. */
int orc_cv_round_f(float v) { return (int)lrintf(v); }
static int cv_round_d(double v) { return (int)lrint(v); }

/* cv::resize(u8, INTER_LINEAR) fixed-point path (OpenCV imgproc resize.cpp: HResizeLinear /
 * VResizeLinear with INTER_RESIZE_COEF_BITS = 11); call site R/src/ORBextractor.cc:1165. */
void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride,
                          uint8_t* dst, int dw, int dh, int dstride)
{
    double inv_scale_x = (double)dw / sw, inv_scale_y = (double)dh / sh;
    double scale_x = 1. / inv_scale_x, scale_y = 1. / inv_scale_y;
    int* xofs = (int*)malloc(sizeof(int) * dw * 2);
    short* ialpha = (short*)malloc(sizeof(short) * dw * 2);
    int* rows[2];
    rows[0] = (int*)malloc(sizeof(int) * dw);
    rows[1] = (int*)malloc(sizeof(int) * dw);
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = (int)floorf(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx * 2] = sx;
        xofs[dx * 2 + 1] = sx + 1 < sw ? sx + 1 : sw - 1;
        ialpha[dx * 2] = (short)orc_cv_round_f((1.f - fx) * 2048.f);
        ialpha[dx * 2 + 1] = (short)orc_cv_round_f(fx * 2048.f);
    }
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = (int)floorf(fy);
        fy -= sy;
        short b0 = (short)orc_cv_round_f((1.f - fy) * 2048.f);
        short b1 = (short)orc_cv_round_f(fy * 2048.f);
        int sy0 = sy < 0 ? 0 : (sy > sh - 1 ? sh - 1 : sy);
        int sy1 = sy + 1 < 0 ? 0 : (sy + 1 > sh - 1 ? sh - 1 : sy + 1);
        const uint8_t* S0 = src + (size_t)sy0 * sstride;
        const uint8_t* S1 = src + (size_t)sy1 * sstride;
        for (int dx = 0; dx < dw; dx++) {
            int x0 = xofs[dx * 2], x1 = xofs[dx * 2 + 1];
            int a0 = ialpha[dx * 2], a1 = ialpha[dx * 2 + 1];
            rows[0][dx] = S0[x0] * a0 + S0[x1] * a1;
            rows[1][dx] = S1[x0] * a0 + S1[x1] * a1;
        }
        uint8_t* D = dst + (size_t)dy * dstride;
        for (int dx = 0; dx < dw; dx++) {
            int v = (((b0 * (rows[0][dx] >> 4)) >> 16) + ((b1 * (rows[1][dx] >> 4)) >> 16) + 2) >> 2;
            D[dx] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
    }
    free(xofs); free(ialpha); free(rows[0]); free(rows[1]);
}

static int reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) {
        if (p < 0) p = -p;
        else p = 2 * (n - 1) - p;
    }
    return p;
}

/* cv::GaussianBlur(u8, Size(7,7), 2, 2, BORDER_REFLECT_101), OpenCV >= 3.4.1 fixed-point path
 * (smooth.simd.hpp, ufixedpoint16 kernel); call site R/src/ORBextractor.cc:1115. */
void orc_gaussian_blur7(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride)
{
    static const int K[7] = {18, 34, 48, 56, 48, 34, 18};
    uint16_t* H = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)w * h);
    for (int y = 0; y < h; y++) {
        const uint8_t* S = src + (size_t)y * sstride;
        for (int x = 0; x < w; x++) {
            int acc = 0;
            if (x >= 3 && x < w - 3) {
                for (int i = 0; i < 7; i++) acc += K[i] * S[x + i - 3];
            } else {
                for (int i = 0; i < 7; i++) acc += K[i] * S[reflect101(x + i - 3, w)];
            }
            H[(size_t)y * w + x] = (uint16_t)acc;
        }
    }
    for (int y = 0; y < h; y++) {
        const uint16_t* R[7];
        for (int i = 0; i < 7; i++) R[i] = H + (size_t)reflect101(y + i - 3, h) * w;
        uint8_t* D = dst + (size_t)y * dstride;
        for (int x = 0; x < w; x++) {
            uint32_t acc = 0;
            for (int i = 0; i < 7; i++) acc += (uint32_t)K[i] * R[i][x];
            D[x] = (uint8_t)((acc + 32768u) >> 16);
        }
    }
    free(H);
}

/* FAST-9/16 ring, (dx,dy) in OpenCV's order (features2d fast_score.cpp makeOffsets, patternSize 16) */
static const int kRing[16][2] = {
    {0, 3}, {1, 3}, {2, 2}, {3, 1}, {3, 0}, {3, -1}, {2, -2}, {1, -3},
    {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

/* m = max over the 16 arcs of 9 contiguous ring pixels of max(min d, min -d); corner <=> m > t;
 * cornerScore<16>() returns max(t, m) - 1 (OpenCV fast_score.cpp).  Sliding min/max by doubling. */
static int fast_arc_measure(const uint8_t* p, const int* off)
{
    int d[25], lo2[24], hi2[24], lo4[22], hi4[22];
    int v = p[0];
    for (int k = 0; k < 16; k++) d[k] = v - p[off[k]];
    for (int k = 16; k < 25; k++) d[k] = d[k - 16];
    for (int k = 0; k < 24; k++) { lo2[k] = d[k] < d[k + 1] ? d[k] : d[k + 1]; hi2[k] = d[k] > d[k + 1] ? d[k] : d[k + 1]; }
    for (int k = 0; k < 22; k++) { lo4[k] = lo2[k] < lo2[k + 2] ? lo2[k] : lo2[k + 2]; hi4[k] = hi2[k] > hi2[k + 2] ? hi2[k] : hi2[k + 2]; }
    int best = -256;
    for (int s = 0; s < 16; s++) {
        int mn = lo4[s] < lo4[s + 4] ? lo4[s] : lo4[s + 4];
        int mx = hi4[s] > hi4[s + 4] ? hi4[s] : hi4[s + 4];
        if (d[s + 8] < mn) mn = d[s + 8];
        if (d[s + 8] > mx) mx = d[s + 8];
        if (mn > best) best = mn;
        if (-mx > best) best = -mx;
    }
    return best;
}

/* OpenCV's high-speed pre-test: every arc of 9 holds one pixel of each opposite pair (k, k+8), so a
 * pixel whose pairs do not all have a member darker than v-t (or all a member brighter than v+t)
 * cannot be a corner at threshold t. */
static int fast_maybe_corner(const uint8_t* p, const int* off, int t)
{
    const int lo = p[0] - t, hi = p[0] + t;
#define CLS(k) ((p[off[k]] < lo ? 1 : 0) | (p[off[k]] > hi ? 2 : 0))
    int d = CLS(0) | CLS(8);
    if (!d) return 0;
    d &= CLS(2) | CLS(10); d &= CLS(4) | CLS(12); d &= CLS(6) | CLS(14);
    if (!d) return 0;
    d &= CLS(1) | CLS(9); d &= CLS(3) | CLS(11); d &= CLS(5) | CLS(13); d &= CLS(7) | CLS(15);
#undef CLS
    return d;
}

/* cv::FAST(img, keypoints, threshold, nonmaxSuppression=true) (features2d fast.cpp FAST_t<16>);
 * call sites R/src/ORBextractor.cc:808, :827.  Output order: row-major. */
int orc_fast9_16(const uint8_t* img, int w, int h, int stride, int threshold, int nms,
                 int32_t* out, int cap)
{
    int n = 0;
    if (threshold < 0) threshold = 0;
    if (threshold > 255) threshold = 255;
    if (w < 7 || h < 7) return 0;
    int off[16];
    for (int k = 0; k < 16; k++) off[k] = kRing[k][1] * stride + kRing[k][0];
    /* score = 0 for non-corners (OpenCV zeroes its row buffers); corner flag kept apart because a
     * corner with m == 1 (only possible at threshold 0) scores 0 */
    int* score = (int*)calloc((size_t)w * h, sizeof(int));
    uint8_t* corner = (uint8_t*)calloc((size_t)w * h, 1);
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            const uint8_t* p = img + (size_t)y * stride + x;
            if (!fast_maybe_corner(p, off, threshold)) continue;
            int m = fast_arc_measure(p, off);
            if (m > threshold) { corner[(size_t)y * w + x] = 1; score[(size_t)y * w + x] = m - 1; }
        }
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            if (!corner[(size_t)y * w + x]) continue;
            int s = score[(size_t)y * w + x];
            if (nms) {
                const int* c = score + (size_t)y * w + x;
                if (!(s > c[-1] && s > c[1] && s > c[-w - 1] && s > c[-w] && s > c[-w + 1] &&
                      s > c[w - 1] && s > c[w] && s > c[w + 1]))
                    continue;
            }
            if (n < cap) { out[n * 3] = x; out[n * 3 + 1] = y; out[n * 3 + 2] = s; }
            n++;
        }
    free(score); free(corner);
    return n;
}

/* cv::fastAtan2 (core mathfuncs_core.simd.hpp atan_f32), OpenCV 3.x/4.x constants; call site
 * R/src/ORBextractor.cc:101.  All ops fp32, no FMA. */
float orc_fast_atan2(float y, float x)
{
    const float scale = (float)(180.0 / 3.141592653589793238462643383279502884);
    const float p1 = 0.9997878412794807f * scale;
    const float p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale;
    const float p7 = -0.04432655554792128f * scale;
    float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

/* ------------------------------------------------------------------------------------------
 * Extractor
 * ---------------------------------------------------------------------------------------- */
typedef struct { float x, y, r; } Cand;

struct OrcExtractor {
    int nfeatures; double scaleFactor; int nlevels, iniTh, minTh;
    float *scale, *invScale, *sigma2, *invSigma2;
    int* featuresPerLevel;
    int umax[HALF_PATCH_SIZE + 1];
    /* per-call state */
    uint8_t** level; uint8_t** blurred; int *lw, *lh;
    Cand** cand; int *ncand, *capcand;
    OrcKeyPoint** lkp; int* nlkp;
};

/* ORBextractor::ORBextractor, R/src/ORBextractor.cc:408-468 */
OrcExtractor* orc_extractor_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th)
{
    OrcExtractor* e = (OrcExtractor*)calloc(1, sizeof(OrcExtractor));
    e->nfeatures = nfeatures; e->scaleFactor = (double)scale_factor; e->nlevels = nlevels;
    e->iniTh = ini_th; e->minTh = min_th;
    e->scale = (float*)calloc(nlevels, sizeof(float));
    e->invScale = (float*)calloc(nlevels, sizeof(float));
    e->sigma2 = (float*)calloc(nlevels, sizeof(float));
    e->invSigma2 = (float*)calloc(nlevels, sizeof(float));
    e->featuresPerLevel = (int*)calloc(nlevels, sizeof(int));
    e->scale[0] = 1.0f; e->sigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) {
        e->scale[i] = (float)(e->scale[i - 1] * e->scaleFactor);      /* float*double -> float, :419 */
        e->sigma2[i] = e->scale[i] * e->scale[i];
    }
    for (int i = 0; i < nlevels; i++) {
        e->invScale[i] = 1.0f / e->scale[i];
        e->invSigma2[i] = 1.0f / e->sigma2[i];
    }
    float factor = (float)(1.0f / e->scaleFactor);                     /* :434 */
    float nDesired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) {
        e->featuresPerLevel[l] = orc_cv_round_f(nDesired);
        sum += e->featuresPerLevel[l];
        nDesired *= factor;
    }
    e->featuresPerLevel[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;

    /* umax, :452-467 */
    int v, v0, vmax = (int)floor(HALF_PATCH_SIZE * sqrtf(2.f) / 2 + 1);
    int vmin = (int)ceil(HALF_PATCH_SIZE * sqrtf(2.f) / 2);
    const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
    for (v = 0; v <= vmax; ++v) e->umax[v] = cv_round_d(sqrt(hp2 - v * v));
    for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
        while (e->umax[v0] == e->umax[v0 + 1]) ++v0;
        e->umax[v] = v0;
        ++v0;
    }
    e->level = (uint8_t**)calloc(nlevels, sizeof(uint8_t*));
    e->blurred = (uint8_t**)calloc(nlevels, sizeof(uint8_t*));
    e->lw = (int*)calloc(nlevels, sizeof(int));
    e->lh = (int*)calloc(nlevels, sizeof(int));
    e->cand = (Cand**)calloc(nlevels, sizeof(Cand*));
    e->ncand = (int*)calloc(nlevels, sizeof(int));
    e->capcand = (int*)calloc(nlevels, sizeof(int));
    e->lkp = (OrcKeyPoint**)calloc(nlevels, sizeof(OrcKeyPoint*));
    e->nlkp = (int*)calloc(nlevels, sizeof(int));
    return e;
}

static void extractor_clear(OrcExtractor* e)
{
    for (int l = 0; l < e->nlevels; l++) {
        free(e->level[l]); e->level[l] = NULL;
        free(e->blurred[l]); e->blurred[l] = NULL;
        free(e->cand[l]); e->cand[l] = NULL; e->ncand[l] = 0; e->capcand[l] = 0;
        free(e->lkp[l]); e->lkp[l] = NULL; e->nlkp[l] = 0;
    }
}

void orc_extractor_destroy(OrcExtractor* e)
{
    if (!e) return;
    extractor_clear(e);
    free(e->scale); free(e->invScale); free(e->sigma2); free(e->invSigma2); free(e->featuresPerLevel);
    free(e->level); free(e->blurred); free(e->lw); free(e->lh);
    free(e->cand); free(e->ncand); free(e->capcand); free(e->lkp); free(e->nlkp);
    free(e);
}

void orc_extractor_tables(const OrcExtractor* e, float* scale, float* inv_scale, float* sigma2,
                          float* inv_sigma2, int32_t* fpl, int32_t* umax16)
{
    for (int i = 0; i < e->nlevels; i++) {
        if (scale) scale[i] = e->scale[i];
        if (inv_scale) inv_scale[i] = e->invScale[i];
        if (sigma2) sigma2[i] = e->sigma2[i];
        if (inv_sigma2) inv_sigma2[i] = e->invSigma2[i];
        if (fpl) fpl[i] = e->featuresPerLevel[i];
    }
    if (umax16) for (int i = 0; i <= HALF_PATCH_SIZE; i++) umax16[i] = e->umax[i];
}

/* ---- DistributeOctTree, R/src/ORBextractor.cc:479-761 ---- */
typedef struct ONode {
    int x0, y0, x1, y1;          /* UL=(x0,y0) UR=(x1,y0) BL=(x0,y1) BR=(x1,y1) */
    int* keys; int nkeys;
    int noMore;
    long seq;                    /* creation sequence = canonical stand-in for the heap address (:682) */
    struct ONode *prev, *next;
} ONode;

typedef struct { ONode *head, *tail; int size; long nextSeq; } OList;

static ONode* onode_new(OList* L, int x0, int y0, int x1, int y1, int cap)
{
    ONode* n = (ONode*)calloc(1, sizeof(ONode));
    n->x0 = x0; n->y0 = y0; n->x1 = x1; n->y1 = y1;
    n->keys = (int*)malloc(sizeof(int) * (cap > 0 ? cap : 1));
    n->seq = L->nextSeq++;
    return n;
}
static void olist_push_front(OList* L, ONode* n)
{
    n->prev = NULL; n->next = L->head;
    if (L->head) L->head->prev = n; else L->tail = n;
    L->head = n; L->size++;
}
static void olist_push_back(OList* L, ONode* n)
{
    n->next = NULL; n->prev = L->tail;
    if (L->tail) L->tail->next = n; else L->head = n;
    L->tail = n; L->size++;
}
static ONode* olist_erase(OList* L, ONode* n)
{
    ONode* nx = n->next;
    if (n->prev) n->prev->next = n->next; else L->head = n->next;
    if (n->next) n->next->prev = n->prev; else L->tail = n->prev;
    L->size--;
    free(n->keys); free(n);
    return nx;
}

/* ExtractorNode::DivideNode, :479-535 */
static void divide_node(OList* L, const ONode* p, const Cand* pts, ONode* c[4])
{
    const int halfX = (int)ceilf((float)(p->x1 - p->x0) / 2);
    const int halfY = (int)ceilf((float)(p->y1 - p->y0) / 2);
    const int mx = p->x0 + halfX, my = p->y0 + halfY;
    c[0] = onode_new(L, p->x0, p->y0, mx, my, p->nkeys);
    c[1] = onode_new(L, mx, p->y0, p->x1, my, p->nkeys);
    c[2] = onode_new(L, p->x0, my, mx, p->y1, p->nkeys);
    c[3] = onode_new(L, mx, my, p->x1, p->y1, p->nkeys);
    for (int i = 0; i < p->nkeys; i++) {
        const Cand* kp = &pts[p->keys[i]];
        ONode* d;
        if (kp->x < (float)mx) d = (kp->y < (float)my) ? c[0] : c[2];
        else d = (kp->y < (float)my) ? c[1] : c[3];
        d->keys[d->nkeys++] = p->keys[i];
    }
    for (int k = 0; k < 4; k++) if (c[k]->nkeys == 1) c[k]->noMore = 1;
}

typedef struct { int size; long seq; ONode* node; } SizeNode;
static int cmp_sizenode(const void* a, const void* b)
{
    const SizeNode* A = (const SizeNode*)a; const SizeNode* B = (const SizeNode*)b;
    if (A->size != B->size) return A->size < B->size ? -1 : 1;
    return A->seq < B->seq ? -1 : (A->seq > B->seq ? 1 : 0);
}

/* push the non-empty children to the front (n1..n4), record the multi-point ones; returns how many recorded */
static int push_children(OList* L, ONode* c[4], SizeNode* vec, int* nvec)
{
    int expand = 0;
    for (int k = 0; k < 4; k++) {
        if (c[k]->nkeys > 0) {
            olist_push_front(L, c[k]);
            if (c[k]->nkeys > 1) {
                expand++;
                vec[*nvec].size = c[k]->nkeys; vec[*nvec].seq = c[k]->seq; vec[*nvec].node = c[k];
                (*nvec)++;
            }
        } else {
            free(c[k]->keys); free(c[k]);
        }
    }
    return expand;
}

static int distribute_octree(const Cand* pts, int n, int minX, int maxX, int minY, int maxY, int N,
                             Cand* out, int cap)
{
    OList L; memset(&L, 0, sizeof(L));
    int nIni = (int)roundf((float)(maxX - minX) / (maxY - minY));
    if (nIni < 1) nIni = 1;     /* the reference would divide by zero; images this tall are out of scope */
    const float hX = (float)(maxX - minX) / nIni;
    ONode** ini = (ONode**)malloc(sizeof(ONode*) * nIni);
    for (int i = 0; i < nIni; i++) {
        ONode* r = onode_new(&L, (int)(hX * (float)i), 0, (int)(hX * (float)(i + 1)), maxY - minY, n);
        olist_push_back(&L, r);
        ini[i] = r;
    }
    for (int i = 0; i < n; i++) {
        int r = (int)(pts[i].x / hX);
        if (r >= nIni) r = nIni - 1;      /* cannot happen for x < maxX-minX; guard only */
        ini[r]->keys[ini[r]->nkeys++] = i;
    }
    free(ini);
    for (ONode* it = L.head; it;) {
        if (it->nkeys == 1) { it->noMore = 1; it = it->next; }
        else if (it->nkeys == 0) it = olist_erase(&L, it);
        else it = it->next;
    }
    int finish = 0;
    int vcap = 4 * (n + 8) + 16;
    SizeNode* vec = (SizeNode*)malloc(sizeof(SizeNode) * vcap);
    SizeNode* prevVec = (SizeNode*)malloc(sizeof(SizeNode) * vcap);
    int nvec = 0;
    while (!finish) {
        int prevSize = L.size;
        int nToExpand = 0;
        nvec = 0;
        for (ONode* it = L.head; it;) {
            if (it->noMore) { it = it->next; continue; }
            ONode* c[4];
            divide_node(&L, it, pts, c);
            nToExpand += push_children(&L, c, vec, &nvec);
            it = olist_erase(&L, it);
        }
        if (L.size >= N || L.size == prevSize) {
            finish = 1;
        } else if (L.size + nToExpand * 3 > N) {
            while (!finish) {
                prevSize = L.size;
                int nprev = nvec;
                memcpy(prevVec, vec, sizeof(SizeNode) * nvec);
                nvec = 0;
                qsort(prevVec, nprev, sizeof(SizeNode), cmp_sizenode);
                for (int j = nprev - 1; j >= 0; j--) {
                    ONode* c[4];
                    divide_node(&L, prevVec[j].node, pts, c);
                    push_children(&L, c, vec, &nvec);
                    olist_erase(&L, prevVec[j].node);
                    if (L.size >= N) break;
                }
                if (L.size >= N || L.size == prevSize) finish = 1;
            }
        }
    }
    free(vec); free(prevVec);
    int nout = 0;
    for (ONode* it = L.head; it; it = it->next) {
        int best = it->keys[0];
        float maxr = pts[best].r;
        for (int k = 1; k < it->nkeys; k++)
            if (pts[it->keys[k]].r > maxr) { best = it->keys[k]; maxr = pts[best].r; }
        if (nout < cap) out[nout] = pts[best];
        nout++;
    }
    while (L.head) olist_erase(&L, L.head);
    return nout;
}

int orc_distribute_octree(const float* xyr, int n, int minX, int maxX, int minY, int maxY, int N,
                          float* out_xyr, int cap)
{
    return distribute_octree((const Cand*)xyr, n, minX, maxX, minY, maxY, N, (Cand*)out_xyr, cap);
}

/* IC_Angle, R/src/ORBextractor.cc:75-102 */
static float ic_angle(const uint8_t* img, int step, float px, float py, const int* umax)
{
    int m_01 = 0, m_10 = 0;
    const uint8_t* center = img + (size_t)orc_cv_round_f(py) * step + orc_cv_round_f(px);
    for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
    for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
        int v_sum = 0;
        int d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int val_plus = center[u + v * step], val_minus = center[u - v * step];
            v_sum += (val_plus - val_minus);
            m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
    }
    return orc_fast_atan2((float)m_01, (float)m_10);
}

/* computeOrbDescriptor, R/src/ORBextractor.cc:105-145.  `cos(angle)` on a float under
 * `using namespace std` (:65) resolves to the float overload, i.e. cosf/sinf. */
static void orb_descriptor(const OrcKeyPoint* kpt, const uint8_t* img, int step, uint8_t* desc)
{
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    float angle = (float)kpt->angle * factorPI;
    float a = cosf(angle), b = sinf(angle);
    const uint8_t* center = img + (size_t)orc_cv_round_f(kpt->y) * step + orc_cv_round_f(kpt->x);
    const int32_t* pat = kPattern;
#define GET_VALUE(idx) \
    center[orc_cv_round_f(pat[2 * (idx)] * b + pat[2 * (idx) + 1] * a) * step + \
           orc_cv_round_f(pat[2 * (idx)] * a - pat[2 * (idx) + 1] * b)]
    for (int i = 0; i < 32; ++i, pat += 32) {
        int val = 0;
        for (int k = 0; k < 8; k++) {
            int t0 = GET_VALUE(2 * k), t1 = GET_VALUE(2 * k + 1);
            val |= (t0 < t1) << k;
        }
        desc[i] = (uint8_t)val;
    }
#undef GET_VALUE
}

static void cand_push(OrcExtractor* e, int l, float x, float y, float r)
{
    if (e->ncand[l] == e->capcand[l]) {
        e->capcand[l] = e->capcand[l] ? e->capcand[l] * 2 : 4096;
        e->cand[l] = (Cand*)realloc(e->cand[l], sizeof(Cand) * e->capcand[l]);
    }
    Cand* c = &e->cand[l][e->ncand[l]++];
    c->x = x; c->y = y; c->r = r;
}

/* ComputePyramid, R/src/ORBextractor.cc:1152-1177.  The 19-px reflected border the reference
 * materialises is never read by anything downstream (all keypoints are >= 19 px inside) and is omitted. */
static void compute_pyramid(OrcExtractor* e, const uint8_t* img, int w, int h, int stride)
{
    for (int l = 0; l < e->nlevels; l++) {
        float scale = e->invScale[l];
        int sw = orc_cv_round_f((float)w * scale), sh = orc_cv_round_f((float)h * scale);
        e->lw[l] = sw; e->lh[l] = sh;
        e->level[l] = (uint8_t*)malloc((size_t)(sw > 0 ? sw : 1) * (sh > 0 ? sh : 1));
        if (l == 0) {
            for (int y = 0; y < h; y++) memcpy(e->level[0] + (size_t)y * w, img + (size_t)y * stride, w);
        } else {
            orc_resize_linear_u8(e->level[l - 1], e->lw[l - 1], e->lh[l - 1], e->lw[l - 1],
                                 e->level[l], sw, sh, sw);
        }
    }
}

/* ComputeKeyPointsOctTree, R/src/ORBextractor.cc:763-878 */
static void compute_keypoints_octree(OrcExtractor* e)
{
    const float W = 30;
    int32_t* cell = (int32_t*)malloc(sizeof(int32_t) * 3 * 8192);
    for (int level = 0; level < e->nlevels; ++level) {
        const int minBorderX = EDGE_THRESHOLD - 3;
        const int minBorderY = minBorderX;
        const int maxBorderX = e->lw[level] - EDGE_THRESHOLD + 3;
        const int maxBorderY = e->lh[level] - EDGE_THRESHOLD + 3;
        const float width = (float)(maxBorderX - minBorderX);
        const float height = (float)(maxBorderY - minBorderY);
        const int nCols = (int)(width / W);
        const int nRows = (int)(height / W);
        e->ncand[level] = 0;
        if (nCols > 0 && nRows > 0) {
            const int wCell = (int)ceilf(width / nCols);
            const int hCell = (int)ceilf(height / nRows);
            for (int i = 0; i < nRows; i++) {
                const float iniY = (float)(minBorderY + i * hCell);
                float maxY = iniY + hCell + 6;
                if (iniY >= maxBorderY - 3) continue;
                if (maxY > maxBorderY) maxY = (float)maxBorderY;
                for (int j = 0; j < nCols; j++) {
                    const float iniX = (float)(minBorderX + j * wCell);
                    float maxX = iniX + wCell + 6;
                    if (iniX >= maxBorderX - 6) continue;
                    if (maxX > maxBorderX) maxX = (float)maxBorderX;
                    const uint8_t* sub = e->level[level] + (size_t)(int)iniY * e->lw[level] + (int)iniX;
                    int cw = (int)maxX - (int)iniX, ch = (int)maxY - (int)iniY;
                    int nc = orc_fast9_16(sub, cw, ch, e->lw[level], e->iniTh, 1, cell, 8192);
                    if (nc == 0) nc = orc_fast9_16(sub, cw, ch, e->lw[level], e->minTh, 1, cell, 8192);
                    for (int k = 0; k < nc; k++)
                        cand_push(e, level, (float)(cell[k * 3] + j * wCell), (float)(cell[k * 3 + 1] + i * hCell),
                                  (float)cell[k * 3 + 2]);
                }
            }
        }
        int N = e->featuresPerLevel[level];
        int cap = e->ncand[level] + 8;
        Cand* kept = (Cand*)malloc(sizeof(Cand) * cap);
        int nk = 0;
        if (e->ncand[level] > 0)
            nk = distribute_octree(e->cand[level], e->ncand[level], minBorderX, maxBorderX, minBorderY, maxBorderY,
                                   N, kept, cap);
        const int scaledPatchSize = (int)(PATCH_SIZE * e->scale[level]);
        e->lkp[level] = (OrcKeyPoint*)malloc(sizeof(OrcKeyPoint) * (nk > 0 ? nk : 1));
        e->nlkp[level] = nk;
        for (int i = 0; i < nk; i++) {
            OrcKeyPoint* kp = &e->lkp[level][i];
            kp->x = kept[i].x + minBorderX;
            kp->y = kept[i].y + minBorderY;
            kp->size = (float)scaledPatchSize;
            kp->angle = -1;
            kp->response = kept[i].r;
            kp->octave = level;
            kp->class_id = -1;
        }
        free(kept);
    }
    free(cell);
    for (int level = 0; level < e->nlevels; ++level)
        for (int i = 0; i < e->nlkp[level]; i++)
            e->lkp[level][i].angle = ic_angle(e->level[level], e->lw[level], e->lkp[level][i].x,
                                              e->lkp[level][i].y, e->umax);
}

/* ORBextractor::operator(), R/src/ORBextractor.cc:1068-1150 */
int orc_extract(OrcExtractor* e, const uint8_t* img, int w, int h, int stride, int lap0, int lap1,
                OrcKeyPoint* kps, uint8_t* desc, int cap, int* n_out)
{
    if (n_out) *n_out = 0;
    if (!img || w <= 0 || h <= 0) return -1;
    extractor_clear(e);
    compute_pyramid(e, img, w, h, stride);
    compute_keypoints_octree(e);
    int nkeypoints = 0;
    for (int l = 0; l < e->nlevels; l++) nkeypoints += e->nlkp[l];
    if (n_out) *n_out = nkeypoints;
    if (nkeypoints > cap) return -2;
    int monoIndex = 0, stereoIndex = nkeypoints - 1;
    uint8_t d[32];
    for (int level = 0; level < e->nlevels; ++level) {
        int nl = e->nlkp[level];
        if (nl == 0) continue;
        e->blurred[level] = (uint8_t*)malloc((size_t)e->lw[level] * e->lh[level]);
        orc_gaussian_blur7(e->level[level], e->lw[level], e->lh[level], e->lw[level], e->blurred[level], e->lw[level]);
        float scale = e->scale[level];
        for (int i = 0; i < nl; i++) {
            OrcKeyPoint kp = e->lkp[level][i];
            orb_descriptor(&kp, e->blurred[level], e->lw[level], d);
            if (level != 0) { kp.x *= scale; kp.y *= scale; }
            int slot;
            if (kp.x >= lap0 && kp.x <= lap1) slot = stereoIndex--;
            else slot = monoIndex++;
            kps[slot] = kp;
            memcpy(desc + (size_t)slot * 32, d, 32);
        }
    }
    return monoIndex;
}

int orc_level_size(const OrcExtractor* e, int level, int* w, int* h)
{
    if (level < 0 || level >= e->nlevels) return -1;
    *w = e->lw[level]; *h = e->lh[level];
    return 0;
}
const uint8_t* orc_level_image(const OrcExtractor* e, int level) { return e->level[level]; }
const uint8_t* orc_level_blurred(const OrcExtractor* e, int level) { return e->blurred[level]; }
int orc_level_candidates(const OrcExtractor* e, int level, const float** xyr)
{
    *xyr = (const float*)e->cand[level];
    return e->ncand[level];
}
int orc_level_keypoints(const OrcExtractor* e, int level, const OrcKeyPoint** kps)
{
    *kps = e->lkp[level];
    return e->nlkp[level];
}

/* ------------------------------------------------------------------------------------------
 * Matching
 * ---------------------------------------------------------------------------------------- */

/* ORBmatcher::DescriptorDistance, R/src/ORBmatcher.cc:2358-2374 (SWAR popcount of 8 x 32 bit) */
int orc_hamming256(const uint8_t* a, const uint8_t* b)
{
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t pa, pb;
        memcpy(&pa, a + 4 * i, 4); memcpy(&pb, b + 4 * i, 4);
        uint32_t v = pa ^ pb;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

/* cv::BFMatcher(NORM_HAMMING).knnMatch(q, t, 2) as called at R/src/Frame.cc:1130: top-2 by
 * (distance, trainIdx) lexicographic. */
void orc_bf_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx, int32_t* dist)
{
    for (int i = 0; i < nq; i++) {
        int b0 = INT_MAX, b1 = INT_MAX, i0 = -1, i1 = -1;
        for (int j = 0; j < nt; j++) {
            int d = orc_hamming256(q + (size_t)i * 32, t + (size_t)j * 32);
            if (d < b0) { b1 = b0; i1 = i0; b0 = d; i0 = j; }
            else if (d < b1) { b1 = d; i1 = j; }
        }
        idx[i * 2] = i0; idx[i * 2 + 1] = i1;
        dist[i * 2] = i0 >= 0 ? b0 : -1; dist[i * 2 + 1] = i1 >= 0 ? b1 : -1;
    }
}

/* Frame::AssignFeaturesToGrid / PosInGrid, R/src/Frame.cc:360-391, 699-709 */
struct OrcGrid {
    float minX, minY, maxX, maxY, wInv, hInv;
    float qminX, qminY;   /* origin of the query cell range: = minX, minY for Frame::GetFeaturesInArea (Frame.cc:639-657); the int-truncated
                           * KeyFrame::mnMinX / mnMinY for KeyFrame::GetFeaturesInArea (KeyFrame.cc:897-911, KeyFrame.h:501-504) */
    int* start;   /* GRID_COLS*GRID_ROWS+1, cell (ix,iy) -> ix*GRID_ROWS+iy */
    int* items;
};

OrcGrid* orc_grid_build(const OrcKeyPoint* kps, int n, float minX, float maxX, float minY, float maxY)
{
    OrcGrid* g = (OrcGrid*)calloc(1, sizeof(OrcGrid));
    g->minX = minX; g->minY = minY; g->maxX = maxX; g->maxY = maxY;
    g->qminX = minX; g->qminY = minY;
    g->wInv = (float)GRID_COLS / (maxX - minX);
    g->hInv = (float)GRID_ROWS / (maxY - minY);
    const int nc = GRID_COLS * GRID_ROWS;
    g->start = (int*)calloc(nc + 1, sizeof(int));
    g->items = (int*)malloc(sizeof(int) * (n > 0 ? n : 1));
    int* cellOf = (int*)malloc(sizeof(int) * (n > 0 ? n : 1));
    for (int i = 0; i < n; i++) {
        int px = (int)roundf((kps[i].x - minX) * g->wInv);
        int py = (int)roundf((kps[i].y - minY) * g->hInv);
        if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) { cellOf[i] = -1; continue; }
        cellOf[i] = px * GRID_ROWS + py;
        g->start[cellOf[i] + 1]++;
    }
    for (int c = 0; c < nc; c++) g->start[c + 1] += g->start[c];
    int* fill = (int*)calloc(nc, sizeof(int));
    for (int i = 0; i < n; i++)
        if (cellOf[i] >= 0) g->items[g->start[cellOf[i]] + fill[cellOf[i]]++] = i;
    free(fill); free(cellOf);
    return g;
}
void orc_grid_set_query_origin(OrcGrid* g, float qminX, float qminY) { g->qminX = qminX; g->qminY = qminY; }
void orc_grid_destroy(OrcGrid* g) { if (g) { free(g->start); free(g->items); free(g); } }

/* Frame::GetFeaturesInArea, R/src/Frame.cc:628-697 */
int orc_features_in_area(const OrcGrid* g, const OrcKeyPoint* kps, float x, float y, float r,
                         int minLevel, int maxLevel, int32_t* out, int cap)
{
    int n = 0;
    const float factorX = r, factorY = r;
    int nMinCellX = (int)floorf((x - g->qminX - factorX) * g->wInv); if (nMinCellX < 0) nMinCellX = 0;
    if (nMinCellX >= GRID_COLS) return 0;
    int nMaxCellX = (int)ceilf((x - g->qminX + factorX) * g->wInv); if (nMaxCellX > GRID_COLS - 1) nMaxCellX = GRID_COLS - 1;
    if (nMaxCellX < 0) return 0;
    int nMinCellY = (int)floorf((y - g->qminY - factorY) * g->hInv); if (nMinCellY < 0) nMinCellY = 0;
    if (nMinCellY >= GRID_ROWS) return 0;
    int nMaxCellY = (int)ceilf((y - g->qminY + factorY) * g->hInv); if (nMaxCellY > GRID_ROWS - 1) nMaxCellY = GRID_ROWS - 1;
    if (nMaxCellY < 0) return 0;
    const int bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
        for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
            int c = ix * GRID_ROWS + iy;
            for (int j = g->start[c]; j < g->start[c + 1]; j++) {
                const OrcKeyPoint* kp = &kps[g->items[j]];
                if (bCheckLevels) {
                    if (kp->octave < minLevel) continue;
                    if (maxLevel >= 0 && kp->octave > maxLevel) continue;
                }
                const float distx = kp->x - x, disty = kp->y - y;
                if (fabsf(distx) < factorX && fabsf(disty) < factorY) {
                    if (n < cap) out[n] = g->items[j];
                    n++;
                }
            }
        }
    return n;
}

/* ORBmatcher::ComputeThreeMaxima, R/src/ORBmatcher.cc:2312-2353 */
static void three_maxima(const int* sizes, int L, int* ind1, int* ind2, int* ind3)
{
    int max1 = 0, max2 = 0, max3 = 0;
    *ind1 = *ind2 = *ind3 = -1;
    for (int i = 0; i < L; i++) {
        const int s = sizes[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; *ind3 = *ind2; *ind2 = *ind1; *ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; *ind3 = *ind2; *ind2 = i; }
        else if (s > max3) { max3 = s; *ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { *ind2 = -1; *ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { *ind3 = -1; }
}

static int rot_bin(float a1, float a2)
{
    const float factor = 1.0f / HISTO_LENGTH;
    float rot = a1 - a2;
    if (rot < 0.0) rot += 360.0f;
    int bin = (int)roundf(rot * factor);
    if (bin == HISTO_LENGTH) bin = 0;
    return bin;
}

/* ORBmatcher::SearchForInitialization, R/src/ORBmatcher.cc:702-817 */
int orc_search_for_initialization(const OrcKeyPoint* k1, const uint8_t* d1, int n1,
                                  const OrcKeyPoint* k2, const uint8_t* d2, int n2,
                                  float minX, float maxX, float minY, float maxY,
                                  float* prev_xy, int32_t* matches12, int window,
                                  float nnratio, int check_ori)
{
    int nmatches = 0;
    OrcGrid* g = orc_grid_build(k2, n2, minX, maxX, minY, maxY);
    int* histIdx = (int*)malloc(sizeof(int) * (n1 > 0 ? n1 : 1));   /* bin of each pushed i1, in push order */
    int* histBin = (int*)malloc(sizeof(int) * (n1 > 0 ? n1 : 1));
    int nhist = 0;
    int sizes[HISTO_LENGTH]; memset(sizes, 0, sizeof(sizes));
    int* matchedDist = (int*)malloc(sizeof(int) * (n2 > 0 ? n2 : 1));
    int* matches21 = (int*)malloc(sizeof(int) * (n2 > 0 ? n2 : 1));
    int32_t* cand = (int32_t*)malloc(sizeof(int32_t) * (n2 > 0 ? n2 : 1));
    for (int i = 0; i < n1; i++) matches12[i] = -1;
    for (int i = 0; i < n2; i++) { matchedDist[i] = INT_MAX; matches21[i] = -1; }
    for (int i1 = 0; i1 < n1; i1++) {
        if (k1[i1].octave > 0) continue;
        int nc = orc_features_in_area(g, k2, prev_xy[i1 * 2], prev_xy[i1 * 2 + 1], (float)window, 0, 0, cand, n2);
        if (nc == 0) continue;
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (int c = 0; c < nc; c++) {
            int i2 = cand[c];
            int dist = orc_hamming256(d1 + (size_t)i1 * 32, d2 + (size_t)i2 * 32);
            if (matchedDist[i2] <= dist) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist <= TH_LOW) {
            if (bestDist < (float)bestDist2 * nnratio) {
                if (matches21[bestIdx2] >= 0) { matches12[matches21[bestIdx2]] = -1; nmatches--; }
                matches12[i1] = bestIdx2;
                matches21[bestIdx2] = i1;
                matchedDist[bestIdx2] = bestDist;
                nmatches++;
                if (check_ori) {
                    int bin = rot_bin(k1[i1].angle, k2[bestIdx2].angle);
                    histIdx[nhist] = i1; histBin[nhist] = bin; nhist++;
                    sizes[bin]++;
                }
            }
        }
    }
    if (check_ori) {
        int ind1, ind2, ind3;
        three_maxima(sizes, HISTO_LENGTH, &ind1, &ind2, &ind3);
        for (int k = 0; k < nhist; k++) {
            int b = histBin[k];
            if (b == ind1 || b == ind2 || b == ind3) continue;
            int idx1 = histIdx[k];
            if (matches12[idx1] >= 0) { matches12[idx1] = -1; nmatches--; }
        }
    }
    for (int i1 = 0; i1 < n1; i1++)
        if (matches12[i1] >= 0) {
            prev_xy[i1 * 2] = k2[matches12[i1]].x;
            prev_xy[i1 * 2 + 1] = k2[matches12[i1]].y;
        }
    free(histIdx); free(histBin); free(matchedDist); free(matches21); free(cand);
    orc_grid_destroy(g);
    return nmatches;
}

/* SearchByProjection, mono/left-image branches, on flat arrays.
 * mode 0: R/src/ORBmatcher.cc:1970-2091 + :2163-2185;  mode 1: R/src/ORBmatcher.cc:44-143. */
int orc_search_by_projection_ex(int mode, const OrcProjQuery* q, const uint8_t* qdesc, int nq,
                                const OrcKeyPoint* k2, const uint8_t* d2, const float* uright2, int n2,
                                float minX, float maxX, float minY, float maxY,
                                int32_t* assigned, float nnratio, int check_ori, int max_dist,
                                const float* inv_sigma2, double chi2, int32_t* best_idx, int32_t* best_dist)
{
    return orc_search_by_projection_full(mode, q, qdesc, nq, k2, d2, uright2, n2, minX, maxX, minY, maxY, minX, minY, assigned, nnratio,
                                         check_ori, max_dist, inv_sigma2, chi2, 0.0, best_idx, best_dist);
}

int orc_search_by_projection_full(int mode, const OrcProjQuery* q, const uint8_t* qdesc, int nq,
                                  const OrcKeyPoint* k2, const uint8_t* d2, const float* uright2, int n2,
                                  float minX, float maxX, float minY, float maxY, float qminX, float qminY,
                                  int32_t* assigned, float nnratio, int check_ori, int max_dist,
                                  const float* inv_sigma2, double chi2, double chi2_stereo, int32_t* best_idx, int32_t* best_dist)
{
    int nmatches = 0;
    if (mode == 3) for (int i = 0; i < nq; i++) { best_idx[i] = -1; best_dist[i] = 256; }
    OrcGrid* g = orc_grid_build(k2, n2, minX, maxX, minY, maxY);
    orc_grid_set_query_origin(g, qminX, qminY);
    int32_t* cand = (int32_t*)malloc(sizeof(int32_t) * (n2 > 0 ? n2 : 1));
    int* histIdx = (int*)malloc(sizeof(int) * (nq > 0 ? nq : 1));
    int* histBin = (int*)malloc(sizeof(int) * (nq > 0 ? nq : 1));
    int nhist = 0;
    int sizes[HISTO_LENGTH]; memset(sizes, 0, sizeof(sizes));
    /* "occupied" = the keypoint holds a MapPoint with Observations() > 0 (ORBmatcher.cc:89-91, :2045-2047).  The caller encodes the
     * state before the call in assigned[] (>= 0 = occupied); a claim made during the call occupies the keypoint only when the
     * claiming MapPoint has observations (valid bit 1 clear): a 0-observation owner (the temporal points Tracking::UpdateLastFrame
     * creates for stereo / RGB-D) leaves it free, so a later query may take it over and both count as matches. */
    uint8_t* occ = (uint8_t*)calloc(n2 > 0 ? n2 : 1, 1);
    if (mode != 3) for (int i = 0; i < n2; i++) occ[i] = assigned[i] >= 0;
    for (int i = 0; i < nq; i++) {
        if (!(q[i].valid & 1)) continue;
        int nc = orc_features_in_area(g, k2, q[i].u, q[i].v, q[i].r, q[i].minl, q[i].maxl, cand, n2);
        if (nc == 0) continue;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int c = 0; c < nc; c++) {
            int i2 = cand[c];
            if (mode != 3 && occ[i2]) continue;
            if (inv_sigma2 && chi2 > 0) {                  /* Fuse: reprojection error gates, ORBmatcher.cc:1525-1552 */
                const float ex = q[i].u - k2[i2].x, ey = q[i].v - k2[i2].y;
                if (chi2_stereo > 0 && uright2 && uright2[i2] >= 0) {
                    const float er = q[i].ur - uright2[i2];
                    const float e2 = ex * ex + ey * ey + er * er;
                    if ((double)(e2 * inv_sigma2[k2[i2].octave]) > chi2_stereo) continue;   /* 7.8 */
                } else {
                    const float e2 = ex * ex + ey * ey;
                    if ((double)(e2 * inv_sigma2[k2[i2].octave]) > chi2) continue;       /* float product against the double literal 5.99 */
                }
            } else if (uright2 && uright2[i2] > 0) {
                const float er = fabsf(q[i].ur - uright2[i2]);
                if (er > q[i].r) continue;
            }
            int dist = orc_hamming256(qdesc + (size_t)i * 32, d2 + (size_t)i2 * 32);
            if (dist < bestDist) {
                bestDist2 = bestDist; bestDist = dist;
                bestLevel2 = bestLevel; bestLevel = k2[i2].octave;
                bestIdx = i2;
            } else if (mode == 1 && dist < bestDist2) {
                bestLevel2 = k2[i2].octave; bestDist2 = dist;
            }
        }
        if (mode == 3) { best_idx[i] = bestIdx; best_dist[i] = bestDist; if (bestIdx >= 0) nmatches++; continue; }
        if (bestDist <= max_dist) {
            if (mode == 1) {
                if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
                assigned[bestIdx] = i;
                if (!(q[i].valid & 2)) occ[bestIdx] = 1;
                nmatches++;
            } else {
                assigned[bestIdx] = i;
                if (!(q[i].valid & 2)) occ[bestIdx] = 1;
                nmatches++;
                if (check_ori) {
                    int bin = rot_bin(q[i].angle, k2[bestIdx].angle);
                    histIdx[nhist] = bestIdx; histBin[nhist] = bin; nhist++;
                    sizes[bin]++;
                }
            }
        }
    }
    if (mode == 0 && check_ori) {
        int ind1, ind2, ind3;
        three_maxima(sizes, HISTO_LENGTH, &ind1, &ind2, &ind3);
        for (int k = 0; k < nhist; k++) {
            int b = histBin[k];
            if (b != ind1 && b != ind2 && b != ind3) { assigned[histIdx[k]] = -2; nmatches--; }   /* -2: claimed, then cleared (:2176-2180) */
        }
    }
    free(cand); free(histIdx); free(histBin); free(occ);
    orc_grid_destroy(g);
    return nmatches;
}

/* Modes 0 / 1 on a two-camera frame (Frame::Nleft != -1), R/src/ORBmatcher.cc:44-214 (mode 1) and :1970-2186 (mode 0), restated on
 * flat arrays: keypoints [0, nL) = mvKeys with grid mGrid, [nL, nL + nR) = mvKeysRight with grid mGridRight (Frame.cc:360-391), the
 * point loop visits the left query of point i, then its right query; one occupancy table over the combined indices. */
int orc_search_by_projection_rig(int mode, const OrcProjQuery* ql, const OrcProjQuery* qr, const uint8_t* qdesc, int nq,
                                 const OrcKeyPoint* k2, const uint8_t* d2, int nL, int nR, const int32_t* l2r, const int32_t* r2l,
                                 float minX, float maxX, float minY, float maxY, int32_t* assigned, float nnratio, int check_ori, int max_dist)
{
    int nmatches = 0;
    const int n2 = nL + nR;
    OrcGrid* gl = orc_grid_build(k2, nL, minX, maxX, minY, maxY);
    OrcGrid* gr = orc_grid_build(k2 + nL, nR, minX, maxX, minY, maxY);
    int32_t* cand = (int32_t*)malloc(sizeof(int32_t) * (n2 > 0 ? n2 : 1));
    int* histIdx = (int*)malloc(sizeof(int) * (nq > 0 ? 2 * nq : 1));
    int* histBin = (int*)malloc(sizeof(int) * (nq > 0 ? 2 * nq : 1));
    int nhist = 0;
    int sizes[HISTO_LENGTH]; memset(sizes, 0, sizeof(sizes));
    uint8_t* occ = (uint8_t*)calloc(n2 > 0 ? n2 : 1, 1);
    for (int i = 0; i < n2; i++) occ[i] = assigned[i] >= 0;
    for (int i = 0; i < nq; i++) {
        const int obs = !(ql[i].valid & 2);
        /* ---- left camera ---- */
        if (ql[i].valid & 1) {
            int nc = orc_features_in_area(gl, k2, ql[i].u, ql[i].v, ql[i].r, ql[i].minl, ql[i].maxl, cand, nL);
            if (nc == 0 && mode == 0) continue;                   /* :2033-2034: an empty left window leaves the point (mode 1 wraps it in an if, :75) */
            if (nc > 0) {
                int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
                for (int c = 0; c < nc; c++) {
                    int i2 = cand[c];
                    if (occ[i2]) continue;
                    int dist = orc_hamming256(qdesc + (size_t)i * 32, d2 + (size_t)i2 * 32);
                    if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = k2[i2].octave; bestIdx = i2; }
                    else if (mode == 1 && dist < bestDist2) { bestLevel2 = k2[i2].octave; bestDist2 = dist; }
                }
                if (mode == 1) {
                    if (bestDist <= TH_HIGH) {
                        if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;      /* :116-117: leaves the point, right search included */
                        assigned[bestIdx] = i; occ[bestIdx] = (uint8_t)obs;
                        if (l2r && l2r[bestIdx] != -1) { assigned[nL + l2r[bestIdx]] = i; occ[nL + l2r[bestIdx]] = (uint8_t)obs; nmatches++; }   /* :123-127 */
                        nmatches++;
                    }
                } else if (bestDist <= max_dist) {
                    assigned[bestIdx] = i; if (obs) occ[bestIdx] = 1;
                    nmatches++;
                    if (check_ori) { int bin = rot_bin(ql[i].angle, k2[bestIdx].angle); histIdx[nhist] = bestIdx; histBin[nhist] = bin; nhist++; sizes[bin]++; }
                }
            }
        }
        /* ---- right camera (:144-212 / :2093-2160) ---- */
        if (qr[i].valid & 1) {
            int nc = orc_features_in_area(gr, k2 + nL, qr[i].u, qr[i].v, qr[i].r, qr[i].minl, qr[i].maxl, cand, nR);
            if (nc == 0) continue;
            int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
            for (int c = 0; c < nc; c++) {
                int i2 = cand[c];
                if (occ[nL + i2]) continue;
                int dist = orc_hamming256(qdesc + (size_t)i * 32, d2 + (size_t)(nL + i2) * 32);
                if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = k2[nL + i2].octave; bestIdx = i2; }
                else if (mode == 1 && dist < bestDist2) { bestLevel2 = k2[nL + i2].octave; bestDist2 = dist; }
            }
            if (mode == 1) {
                if (bestDist <= TH_HIGH) {
                    if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
                    if (r2l && r2l[bestIdx] != -1) { assigned[r2l[bestIdx]] = i; occ[r2l[bestIdx]] = (uint8_t)obs; nmatches++; }   /* :191-195 */
                    assigned[nL + bestIdx] = i; occ[nL + bestIdx] = (uint8_t)obs;
                    nmatches++;
                }
            } else if (bestDist <= max_dist) {
                assigned[nL + bestIdx] = i; if (obs) occ[nL + bestIdx] = 1;
                nmatches++;
                if (check_ori) { int bin = rot_bin(qr[i].angle, k2[nL + bestIdx].angle); histIdx[nhist] = nL + bestIdx; histBin[nhist] = bin; nhist++; sizes[bin]++; }
            }
        }
    }
    if (mode == 0 && check_ori) {
        int ind1, ind2, ind3;
        three_maxima(sizes, HISTO_LENGTH, &ind1, &ind2, &ind3);
        for (int k = 0; k < nhist; k++) {
            int b = histBin[k];
            if (b != ind1 && b != ind2 && b != ind3) { assigned[histIdx[k]] = -2; nmatches--; }
        }
    }
    free(cand); free(histIdx); free(histBin); free(occ);
    orc_grid_destroy(gl); orc_grid_destroy(gr);
    return nmatches;
}

int orc_search_by_projection(int mode, const OrcProjQuery* q, const uint8_t* qdesc, int nq,
                             const OrcKeyPoint* k2, const uint8_t* d2, const float* uright2, int n2,
                             float minX, float maxX, float minY, float maxY,
                             int32_t* assigned, float nnratio, int check_ori)
{
    return orc_search_by_projection_ex(mode, q, qdesc, nq, k2, d2, uright2, n2, minX, maxX, minY, maxY, assigned, nnratio, check_ori,
                                       TH_HIGH, NULL, 0.0, NULL, NULL);
}

/* Frame::ComputeStereoMatches, descriptor search, R/src/Frame.cc:785-868 */
void orc_stereo_band_match(const OrcKeyPoint* kl, const uint8_t* dl, int nl,
                           const OrcKeyPoint* kr, const uint8_t* dr, int nr,
                           const float* scale_factors, int nrows, float minD, float maxD,
                           int32_t* best_idx, int32_t* best_dist)
{
    int* cnt = (int*)calloc(nrows + 1, sizeof(int));
    for (int iR = 0; iR < nr; iR++) {
        const float kpY = kr[iR].y;
        const float r = 2.0f * scale_factors[kr[iR].octave];
        const int maxr = (int)ceilf(kpY + r), minr = (int)floorf(kpY - r);
        for (int yi = minr; yi <= maxr; yi++) if (yi >= 0 && yi < nrows) cnt[yi + 1]++;
    }
    for (int y = 0; y < nrows; y++) cnt[y + 1] += cnt[y];
    int* items = (int*)malloc(sizeof(int) * (cnt[nrows] > 0 ? cnt[nrows] : 1));
    int* fill = (int*)calloc(nrows, sizeof(int));
    for (int iR = 0; iR < nr; iR++) {
        const float kpY = kr[iR].y;
        const float r = 2.0f * scale_factors[kr[iR].octave];
        const int maxr = (int)ceilf(kpY + r), minr = (int)floorf(kpY - r);
        for (int yi = minr; yi <= maxr; yi++)
            if (yi >= 0 && yi < nrows) items[cnt[yi] + fill[yi]++] = iR;
    }
    for (int iL = 0; iL < nl; iL++) {
        best_idx[iL] = -1; best_dist[iL] = TH_HIGH;
        const int levelL = kl[iL].octave;
        const float vL = kl[iL].y, uL = kl[iL].x;
        int row = (int)vL;
        if (row < 0 || row >= nrows) continue;
        if (cnt[row + 1] == cnt[row]) continue;
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = TH_HIGH, bestIdxR = -1;
        for (int c = cnt[row]; c < cnt[row + 1]; c++) {
            const int iR = items[c];
            if (kr[iR].octave < levelL - 1 || kr[iR].octave > levelL + 1) continue;
            const float uR = kr[iR].x;
            if (uR >= minU && uR <= maxU) {
                const int dist = orc_hamming256(dl + (size_t)iL * 32, dr + (size_t)iR * 32);
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        best_idx[iL] = bestIdxR; best_dist[iL] = bestDist;
    }
    free(cnt); free(items); free(fill);
}

/* Frame::ComputeStereoMatches, R/src/Frame.cc:785-962 (descriptor search :785-868 as above, then the SAD refinement) */
typedef struct { int dist; int idx; } DistIdx;
static int cmp_distidx(const void* a, const void* b)
{
    const DistIdx* A = (const DistIdx*)a; const DistIdx* B = (const DistIdx*)b;
    if (A->dist != B->dist) return A->dist < B->dist ? -1 : 1;
    return A->idx < B->idx ? -1 : (A->idx > B->idx ? 1 : 0);
}

void orc_compute_stereo_matches(const OrcExtractor* left, const OrcExtractor* right,
                                const OrcKeyPoint* kl, const uint8_t* dl, int nl,
                                const OrcKeyPoint* kr, const uint8_t* dr, int nr,
                                float mb, float mbf, float* uright, float* depth, int32_t* sad_dist)
{
    const int thOrbDist = (TH_HIGH + TH_LOW) / 2;
    const int nRows = left->lh[0];
    const float minZ = mb, minD = 0, maxD = mbf / minZ;
    int32_t* bi = (int32_t*)malloc(sizeof(int32_t) * (nl > 0 ? nl : 1));
    int32_t* bd = (int32_t*)malloc(sizeof(int32_t) * (nl > 0 ? nl : 1));
    orc_stereo_band_match(kl, dl, nl, kr, dr, nr, left->scale, nRows, minD, maxD, bi, bd);
    DistIdx* vDistIdx = (DistIdx*)malloc(sizeof(DistIdx) * (nl > 0 ? nl : 1));
    int nvd = 0;
    for (int iL = 0; iL < nl; iL++) {
        uright[iL] = -1.0f; depth[iL] = -1.0f;
        if (sad_dist) sad_dist[iL] = -1;
        if (bi[iL] < 0 || !(bd[iL] < thOrbDist)) continue;
        const OrcKeyPoint* kpL = &kl[iL];
        const float uL = kpL->x;
        const int oct = kpL->octave;
        const float uR0 = kr[bi[iL]].x;
        const float scaleFactor = left->invScale[oct];
        const float scaleduL = roundf(kpL->x * scaleFactor);
        const float scaledvL = roundf(kpL->y * scaleFactor);
        const float scaleduR0 = roundf(uR0 * scaleFactor);
        const int w = 5, L = 5;
        const uint8_t* imL = left->level[oct]; const int wL = left->lw[oct];
        const uint8_t* imR = right->level[oct]; const int wR = right->lw[oct];
        const int r0 = (int)(scaledvL - w), c0 = (int)(scaleduL - w);
        short IL[11][11];
        {
            const int ctr = imL[(size_t)(r0 + w) * wL + c0 + w];
            for (int r = 0; r < 11; r++) for (int c = 0; c < 11; c++) IL[r][c] = (short)(imL[(size_t)(r0 + r) * wL + c0 + c] - ctr);
        }
        int bestDist = INT_MAX, bestincR = 0;
        float vDists[11];
        const float iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1;
        if (iniu < 0 || endu >= right->lw[oct]) continue;
        for (int incR = -L; incR <= +L; incR++) {
            const int cr = (int)(scaleduR0 + incR - w);
            const int ctr = imR[(size_t)(r0 + w) * wR + cr + w];
            double acc = 0;
            for (int r = 0; r < 11; r++)
                for (int c = 0; c < 11; c++) acc += abs((int)IL[r][c] - ((int)imR[(size_t)(r0 + r) * wR + cr + c] - ctr));
            const float dist = (float)acc;                       /* cv::norm(IL, IR, NORM_L1) */
            if (dist < bestDist) { bestDist = (int)dist; bestincR = incR; }
            vDists[L + incR] = dist;
        }
        if (bestincR == -L || bestincR == L) continue;
        const float dist1 = vDists[L + bestincR - 1], dist2 = vDists[L + bestincR], dist3 = vDists[L + bestincR + 1];
        const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
        if (deltaR < -1 || deltaR > 1) continue;
        float bestuR = left->scale[oct] * ((float)scaleduR0 + (float)bestincR + deltaR);
        float disparity = (uL - bestuR);
        if (disparity >= minD && disparity < maxD) {
            if (disparity <= 0) { disparity = 0.01; bestuR = uL - 0.01; }
            depth[iL] = mbf / disparity;
            uright[iL] = bestuR;
            if (sad_dist) sad_dist[iL] = bestDist;
            vDistIdx[nvd].dist = bestDist; vDistIdx[nvd].idx = iL; nvd++;
        }
    }
    if (nvd > 0) {     /* the reference indexes an empty vector here when nothing matched (:950) */
        qsort(vDistIdx, nvd, sizeof(DistIdx), cmp_distidx);
        const float median = (float)vDistIdx[nvd / 2].dist;
        const float thDist = 1.5f * 1.4f * median;
        for (int i = nvd - 1; i >= 0; i--) {
            if (vDistIdx[i].dist < thDist) break;
            uright[vDistIdx[i].idx] = -1; depth[vDistIdx[i].idx] = -1;
        }
    }
    free(bi); free(bd); free(vDistIdx);
}

/* best / second-best over an explicit candidate list, e.g. ORBmatcher.cc:320-340 (SearchByBoW inner loop) */
void orc_match_candidates(const uint8_t* q, int nq, const uint8_t* t, const int32_t* offsets, const int32_t* indices,
                          int32_t* out_idx, int32_t* out_dist)
{
    for (int i = 0; i < nq; i++) {
        int b0 = INT_MAX, b1 = INT_MAX, i0 = -1, i1 = -1;
        for (int k = offsets[i]; k < offsets[i + 1]; k++) {
            const int j = indices[k];
            const int d = orc_hamming256(q + (size_t)i * 32, t + (size_t)j * 32);
            if (d < b0) { b1 = b0; i1 = i0; b0 = d; i0 = j; }
            else if (d < b1) { b1 = d; i1 = j; }
        }
        out_idx[i * 2] = i0; out_idx[i * 2 + 1] = i1;
        out_dist[i * 2] = i0 >= 0 ? b0 : -1; out_dist[i * 2 + 1] = i1 >= 0 ? b1 : -1;
    }
}

/* ------------------------------------------------------------------------------------------------------------------
 * Bag of words (DBoW2 as vendored by the reference, R/Thirdparty/DBoW2/DBoW2)
 * ------------------------------------------------------------------------------------------------------------------ */
struct OrcVocab {
    int n_nodes, L;
    int32_t* child_start;   /* CSR over children, order of appearance (loadFromTextFile: m_nodes[pid].children.push_back) */
    int32_t* child;
    uint8_t* desc;
    double* weight;
    int32_t* word_id;       /* -1 for inner nodes */
};

OrcVocab* orc_vocab_create(int n_nodes, const int32_t* parent, const uint8_t* is_leaf, const uint8_t* desc,
                           const double* weight, int L)
{
    if (n_nodes < 1) return NULL;
    for (int i = 1; i < n_nodes; i++) if (parent[i] < 0 || parent[i] >= i) return NULL;
    OrcVocab* v = (OrcVocab*)calloc(1, sizeof(OrcVocab));
    v->n_nodes = n_nodes; v->L = L;
    v->child_start = (int32_t*)calloc((size_t)n_nodes + 1, sizeof(int32_t));
    v->child = (int32_t*)calloc((size_t)n_nodes, sizeof(int32_t));
    v->desc = (uint8_t*)malloc((size_t)n_nodes * 32); memcpy(v->desc, desc, (size_t)n_nodes * 32);
    v->weight = (double*)malloc(sizeof(double) * n_nodes); memcpy(v->weight, weight, sizeof(double) * n_nodes);
    v->word_id = (int32_t*)malloc(sizeof(int32_t) * n_nodes);
    for (int i = 1; i < n_nodes; i++) v->child_start[parent[i] + 1]++;
    for (int i = 0; i < n_nodes; i++) v->child_start[i + 1] += v->child_start[i];
    int32_t* cur = (int32_t*)malloc(sizeof(int32_t) * n_nodes);
    memcpy(cur, v->child_start, sizeof(int32_t) * n_nodes);
    for (int i = 1; i < n_nodes; i++) v->child[cur[parent[i]]++] = i;
    free(cur);
    int words = 0;
    for (int i = 0; i < n_nodes; i++) v->word_id[i] = (i > 0 && is_leaf[i]) ? words++ : -1;   /* :TemplatedVocabulary.h loadFromTextFile */
    return v;
}

void orc_vocab_destroy(OrcVocab* v)
{
    if (!v) return;
    free(v->child_start); free(v->child); free(v->desc); free(v->weight); free(v->word_id); free(v);
}

void orc_bow_transform_features(const OrcVocab* v, const uint8_t* desc, int n, int levelsup,
                                int32_t* word_id, double* weight, int32_t* node_id)
{
    const int nid_level = v->L - levelsup;
    for (int f = 0; f < n; f++) {
        const uint8_t* d = desc + (size_t)f * 32;
        int final_id = 0, level = 0, nid = 0, have_nid = nid_level <= 0;
        if (v->child_start[1] == v->child_start[0]) { word_id[f] = -1; weight[f] = 0.0; node_id[f] = 0; continue; }   /* empty vocabulary */
        do {                                                         /* TemplatedVocabulary.h:1234-1254 */
            ++level;
            const int c0 = v->child_start[final_id], c1 = v->child_start[final_id + 1];
            int best = v->child[c0];
            int best_d = orc_hamming256(d, v->desc + (size_t)best * 32);
            for (int c = c0 + 1; c < c1; c++) {
                const int id = v->child[c];
                const int dd = orc_hamming256(d, v->desc + (size_t)id * 32);
                if (dd < best_d) { best_d = dd; best = id; }
            }
            final_id = best;
            if (level == nid_level) { nid = final_id; have_nid = 1; }
        } while (v->child_start[final_id + 1] > v->child_start[final_id]);
        if (!have_nid) nid = final_id;
        word_id[f] = v->word_id[final_id]; weight[f] = v->weight[final_id]; node_id[f] = nid;
    }
}

typedef struct { int32_t key; int32_t idx; } OrcKI;
static int orc_ki_cmp(const void* a, const void* b)
{
    const OrcKI* x = (const OrcKI*)a; const OrcKI* y = (const OrcKI*)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

int orc_bow_transform(const OrcVocab* v, const uint8_t* desc, int n, int levelsup,
                      int32_t* bow_words, double* bow_values,
                      int32_t* fv_nodes, int32_t* fv_start, int32_t* fv_features, int* n_fv)
{
    int32_t* w = (int32_t*)malloc(sizeof(int32_t) * (n + 1)); double* wt = (double*)malloc(sizeof(double) * (n + 1));
    int32_t* nd = (int32_t*)malloc(sizeof(int32_t) * (n + 1));
    orc_bow_transform_features(v, desc, n, levelsup, w, wt, nd);
    OrcKI* kw = (OrcKI*)malloc(sizeof(OrcKI) * (n + 1)); OrcKI* kn = (OrcKI*)malloc(sizeof(OrcKI) * (n + 1));
    int m = 0;
    for (int i = 0; i < n; i++)
        if (wt[i] > 0) { kw[m].key = w[i]; kw[m].idx = i; kn[m].key = nd[i]; kn[m].idx = i; m++; }      /* "not stopped" :1157 */
    qsort(kw, m, sizeof(OrcKI), orc_ki_cmp); qsort(kn, m, sizeof(OrcKI), orc_ki_cmp);
    /* BowVector::addWeight in feature order: the value of a word is w added once per feature that hit it */
    int nb = 0;
    for (int i = 0; i < m; ) {
        int j = i; double acc = 0.0;
        while (j < m && kw[j].key == kw[i].key) { if (j == i) acc = wt[kw[j].idx]; else acc += wt[kw[j].idx]; j++; }
        bow_words[nb] = kw[i].key; bow_values[nb] = acc; nb++;
        i = j;
    }
    /* BowVector::normalize(L1): norm summed in map (word id) order */
    double norm = 0.0;
    for (int i = 0; i < nb; i++) norm += fabs(bow_values[i]);
    if (norm > 0.0) for (int i = 0; i < nb; i++) bow_values[i] /= norm;
    int nf = 0;
    for (int i = 0; i < m; ) {
        int j = i;
        fv_nodes[nf] = kn[i].key; fv_start[nf] = i;
        while (j < m && kn[j].key == kn[i].key) { fv_features[j] = kn[j].idx; j++; }
        nf++; i = j;
    }
    fv_start[nf] = m;
    *n_fv = nf;
    free(w); free(wt); free(nd); free(kw); free(kn);
    return nb;
}

/* ORBmatcher::SearchByBoW, R/src/ORBmatcher.cc:269-471 (mode 0) and :819-959 (mode 1) */
int orc_search_by_bow(int mode, const OrcKeyPoint* k1, const uint8_t* d1, const uint8_t* valid1, int n1,
                      const int32_t* fv1_nodes, const int32_t* fv1_start, const int32_t* fv1_feat, int nfv1,
                      const OrcKeyPoint* k2, const uint8_t* d2, const uint8_t* valid2, int n2,
                      const int32_t* fv2_nodes, const int32_t* fv2_start, const int32_t* fv2_feat, int nfv2,
                      float nnratio, int check_ori, int32_t* matches12)
{
    int nmatches = 0;
    uint8_t* matched2 = (uint8_t*)calloc(n2 > 0 ? n2 : 1, 1);
    int* histIdx = (int*)malloc(sizeof(int) * (n1 > 0 ? n1 : 1));
    int* histBin = (int*)malloc(sizeof(int) * (n1 > 0 ? n1 : 1));
    int hist[HISTO_LENGTH]; int nh = 0;
    for (int i = 0; i < HISTO_LENGTH; i++) hist[i] = 0;
    for (int i = 0; i < n1; i++) matches12[i] = -1;
    int a = 0, b = 0;
    while (a < nfv1 && b < nfv2) {
        if (fv1_nodes[a] == fv2_nodes[b]) {
            for (int ia = fv1_start[a]; ia < fv1_start[a + 1]; ia++) {
                const int i1 = fv1_feat[ia];
                if (!valid1[i1]) continue;
                int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
                for (int ib = fv2_start[b]; ib < fv2_start[b + 1]; ib++) {
                    const int i2 = fv2_feat[ib];
                    if (matched2[i2]) continue;
                    if (mode == 1 && valid2 && !valid2[i2]) continue;
                    const int dist = orc_hamming256(d1 + (size_t)i1 * 32, d2 + (size_t)i2 * 32);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = i2; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                const int pass = mode == 0 ? bestDist1 <= TH_LOW : bestDist1 < TH_LOW;
                if (pass && (float)bestDist1 < nnratio * (float)bestDist2) {
                    matches12[i1] = bestIdx2; matched2[bestIdx2] = 1;
                    if (check_ori) { const int bin = rot_bin(k1[i1].angle, k2[bestIdx2].angle); histIdx[nh] = i1; histBin[nh] = bin; nh++; hist[bin]++; }
                    nmatches++;
                }
            }
            a++; b++;
        } else if (fv1_nodes[a] < fv2_nodes[b]) {
            while (a < nfv1 && fv1_nodes[a] < fv2_nodes[b]) a++;          /* lower_bound */
        } else {
            while (b < nfv2 && fv2_nodes[b] < fv1_nodes[a]) b++;
        }
    }
    if (check_ori) {
        int ind1, ind2, ind3;
        three_maxima(hist, HISTO_LENGTH, &ind1, &ind2, &ind3);
        for (int j = 0; j < nh; j++)
            if (histBin[j] != ind1 && histBin[j] != ind2 && histBin[j] != ind3) { matches12[histIdx[j]] = -1; nmatches--; }
    }
    free(matched2); free(histIdx); free(histBin);
    (void)n2;
    return nmatches;
}

/* SearchByBoW(KeyFrame*, Frame&) on a two-camera frame (Frame::Nleft != -1), R/src/ORBmatcher.cc:344-431 restated: features
 * [0, n2_left) of the frame are the left camera's.  matches12l / matches12r [n1]: the left / right feature taken by KF feature i1. */
int orc_search_by_bow_rig(const OrcKeyPoint* k1, const uint8_t* d1, const uint8_t* valid1, int n1,
                          const int32_t* fv1_nodes, const int32_t* fv1_start, const int32_t* fv1_feat, int nfv1,
                          const OrcKeyPoint* k2, const uint8_t* d2, int n2, int n2_left,
                          const int32_t* fv2_nodes, const int32_t* fv2_start, const int32_t* fv2_feat, int nfv2,
                          float nnratio, int check_ori, int32_t* matches12l, int32_t* matches12r)
{
    int nmatches = 0;
    uint8_t* matched2 = (uint8_t*)calloc(n2 > 0 ? n2 : 1, 1);
    int* histIdx = (int*)malloc(sizeof(int) * (n1 > 0 ? 2 * n1 : 1));     /* i1, or n1 + i1 for the right match */
    int* histBin = (int*)malloc(sizeof(int) * (n1 > 0 ? 2 * n1 : 1));
    int hist[HISTO_LENGTH]; int nh = 0;
    for (int i = 0; i < HISTO_LENGTH; i++) hist[i] = 0;
    for (int i = 0; i < n1; i++) { matches12l[i] = -1; matches12r[i] = -1; }
    int a = 0, b = 0;
    while (a < nfv1 && b < nfv2) {
        if (fv1_nodes[a] == fv2_nodes[b]) {
            for (int ia = fv1_start[a]; ia < fv1_start[a + 1]; ia++) {
                const int i1 = fv1_feat[ia];
                if (!valid1[i1]) continue;
                int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256, bestDist1R = 256, bestIdxFR = -1, bestDist2R = 256;
                for (int ib = fv2_start[b]; ib < fv2_start[b + 1]; ib++) {
                    const int i2 = fv2_feat[ib];
                    if (matched2[i2]) continue;
                    const int dist = orc_hamming256(d1 + (size_t)i1 * 32, d2 + (size_t)i2 * 32);
                    if (i2 < n2_left && dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = i2; }
                    else if (i2 < n2_left && dist < bestDist2) bestDist2 = dist;
                    if (i2 >= n2_left && dist < bestDist1R) { bestDist2R = bestDist1R; bestDist1R = dist; bestIdxFR = i2; }
                    else if (i2 >= n2_left && dist < bestDist2R) bestDist2R = dist;
                }
                if (bestDist1 <= TH_LOW) {
                    if ((float)bestDist1 < nnratio * (float)bestDist2) {
                        matches12l[i1] = bestIdxF; matched2[bestIdxF] = 1;
                        if (check_ori) { const int bin = rot_bin(k1[i1].angle, k2[bestIdxF].angle); histIdx[nh] = i1; histBin[nh] = bin; nh++; hist[bin]++; }
                        nmatches++;
                    }
                    if (bestDist1R <= TH_LOW) {                                  /* `|| true`: no ratio test on this side (:402) */
                        matches12r[i1] = bestIdxFR; matched2[bestIdxFR] = 1;
                        if (check_ori) { const int bin = rot_bin(k1[i1].angle, k2[bestIdxFR].angle); histIdx[nh] = n1 + i1; histBin[nh] = bin; nh++; hist[bin]++; }
                        nmatches++;
                    }
                }
                (void)bestDist2R;
            }
            a++; b++;
        } else if (fv1_nodes[a] < fv2_nodes[b]) {
            while (a < nfv1 && fv1_nodes[a] < fv2_nodes[b]) a++;
        } else {
            while (b < nfv2 && fv2_nodes[b] < fv1_nodes[a]) b++;
        }
    }
    if (check_ori) {
        int ind1, ind2, ind3;
        three_maxima(hist, HISTO_LENGTH, &ind1, &ind2, &ind3);
        for (int j = 0; j < nh; j++)
            if (histBin[j] != ind1 && histBin[j] != ind2 && histBin[j] != ind3) {
                if (histIdx[j] < n1) matches12l[histIdx[j]] = -1; else matches12r[histIdx[j] - n1] = -1;
                nmatches--;
            }
    }
    free(matched2); free(histIdx); free(histBin);
    return nmatches;
}

/* MapPoint::ComputeDistinctiveDescriptors, R/src/MapPoint.cc:487-518 */
static int orc_int_cmp(const void* a, const void* b) { const int x = *(const int*)a, y = *(const int*)b; return x < y ? -1 : (x > y); }
void orc_distinctive_descriptors(const uint8_t* desc, const int32_t* offsets, int npoints, int32_t* best)
{
    for (int p = 0; p < npoints; p++) {
        const int N = offsets[p + 1] - offsets[p];
        if (N <= 0) { best[p] = -1; continue; }
        const uint8_t* d = desc + (size_t)offsets[p] * 32;
        int* row = (int*)malloc(sizeof(int) * N);
        int BestMedian = INT_MAX, BestIdx = 0;
        for (int i = 0; i < N; i++) {
            for (int j = 0; j < N; j++) row[j] = i == j ? 0 : orc_hamming256(d + (size_t)i * 32, d + (size_t)j * 32);
            qsort(row, N, sizeof(int), orc_int_cmp);
            const int median = row[(int)(0.5 * (N - 1))];
            if (median < BestMedian) { BestMedian = median; BestIdx = i; }
        }
        best[p] = BestIdx;
        free(row);
    }
}

/* Frame::UndistortKeyPoints, R/src/Frame.cc:721-754; the arithmetic is cv::undistortPoints of OpenCV 4.x
 * (modules/calib3d/src/undistort.dispatch.cpp, cvUndistortPointsInternal) with the default TermCriteria(MAX_ITER, 5, 0.01) */
void orc_undistort_keypoints(const OrcKeyPoint* kps, int n, const float* K, const float* dist, int ndist, const float* P,
                             OrcKeyPoint* out)
{
    if (ndist < 1 || dist[0] == 0.0f) { for (int i = 0; i < n; i++) out[i] = kps[i]; return; }
    double k[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < ndist && i < 12; i++) k[i] = (double)dist[i];
    const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    const double ifx = 1. / fx, ify = 1. / fy;
    double RR[9];
    for (int i = 0; i < 9; i++) RR[i] = (double)P[i];
    for (int i = 0; i < n; i++) {
        const double u = kps[i].x, v = kps[i].y;
        double x = (u - cx) * ifx, y = (v - cy) * ify;
        const double x0 = x, y0 = y;
        for (int j = 0; j < 5; j++) {
            const double r2 = x * x + y * y;
            const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
            if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }
            const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
            const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
            x = (x0 - deltaX) * icdist;
            y = (y0 - deltaY) * icdist;
        }
        const double xx = RR[0] * x + RR[1] * y + RR[2], yy = RR[3] * x + RR[4] * y + RR[5];
        const double ww = 1. / (RR[6] * x + RR[7] * y + RR[8]);
        out[i] = kps[i];
        out[i].x = (float)(xx * ww); out[i].y = (float)(yy * ww);
    }
}

/* Converter::toCvKeyPointMsg / fromCvKeyPointMsg, R/src/Converter.cc:218-244, in ROS1 wire layout (15 bytes per keypoint) */
void orc_keypoints_to_msg(const OrcKeyPoint* kps, int n, uint8_t* msg)
{
    for (int i = 0; i < n; i++) {
        uint8_t* o = msg + (size_t)i * 15;
        const uint8_t size = (uint8_t)(int32_t)kps[i].size, response = (uint8_t)(int32_t)kps[i].response;
        const int8_t octave = (int8_t)kps[i].octave;
        memcpy(o, &kps[i].x, 4); memcpy(o + 4, &kps[i].y, 4); o[8] = size; memcpy(o + 9, &kps[i].angle, 4);
        o[13] = response; o[14] = (uint8_t)octave;
    }
}

void orc_keypoints_from_msg(const uint8_t* msg, int n, OrcKeyPoint* kps)
{
    for (int i = 0; i < n; i++) {
        const uint8_t* o = msg + (size_t)i * 15;
        memcpy(&kps[i].x, o, 4); memcpy(&kps[i].y, o + 4, 4); memcpy(&kps[i].angle, o + 9, 4);
        kps[i].size = (float)o[8]; kps[i].response = (float)o[13]; kps[i].octave = (int8_t)o[14]; kps[i].class_id = -1;
    }
}

/* Pinhole::epipolarConstrain with F12 given, R/src/CameraModels/Pinhole.cpp:128-142 */
static int epipolar_ok(const OrcKeyPoint* kp1, const OrcKeyPoint* kp2, const float* F, float unc)
{
    const float a = kp1->x * F[0] + kp1->y * F[3] + F[6];
    const float b = kp1->x * F[1] + kp1->y * F[4] + F[7];
    const float c = kp1->x * F[2] + kp1->y * F[5] + F[8];
    const float num = a * kp2->x + b * kp2->y + c;
    const float den = a * a + b * b;
    if (den == 0) return 0;
    const float dsqr = num * num / den;
    return dsqr < 3.84 * unc;
}

/* ORBmatcher::SearchForTriangulation, R/src/ORBmatcher.cc:961-1202 (mpCamera2 == NULL) */
int orc_search_for_triangulation(const OrcKeyPoint* k1, const uint8_t* d1, const uint8_t* free1, const uint8_t* stereo1, int n1,
                                 const int32_t* fv1_nodes, const int32_t* fv1_start, const int32_t* fv1_feat, int nfv1,
                                 const OrcKeyPoint* k2, const uint8_t* d2, const uint8_t* free2, const uint8_t* stereo2, int n2,
                                 const int32_t* fv2_nodes, const int32_t* fv2_start, const int32_t* fv2_feat, int nfv2,
                                 const float* F12, float ep_x, float ep_y, const float* scale_factors2, const float* level_sigma2_2,
                                 int only_stereo, int coarse, int check_ori, int32_t* matches12)
{
    int nmatches = 0;
    int* histIdx = (int*)malloc(sizeof(int) * (n1 > 0 ? n1 : 1));
    int* histBin = (int*)malloc(sizeof(int) * (n1 > 0 ? n1 : 1));
    int hist[HISTO_LENGTH]; int nh = 0;
    for (int i = 0; i < HISTO_LENGTH; i++) hist[i] = 0;
    for (int i = 0; i < n1; i++) matches12[i] = -1;
    int a = 0, b = 0;
    while (a < nfv1 && b < nfv2) {
        if (fv1_nodes[a] == fv2_nodes[b]) {
            for (int ia = fv1_start[a]; ia < fv1_start[a + 1]; ia++) {
                const int idx1 = fv1_feat[ia];
                if (!free1[idx1]) continue;
                const int bStereo1 = stereo1 ? stereo1[idx1] : 0;
                if (only_stereo && !bStereo1) continue;
                int bestDist = TH_LOW, bestIdx2 = -1;
                for (int ib = fv2_start[b]; ib < fv2_start[b + 1]; ib++) {
                    const int idx2 = fv2_feat[ib];
                    if (!free2[idx2]) continue;
                    const int bStereo2 = stereo2 ? stereo2[idx2] : 0;
                    if (only_stereo && !bStereo2) continue;
                    const int dist = orc_hamming256(d1 + (size_t)idx1 * 32, d2 + (size_t)idx2 * 32);
                    if (dist > TH_LOW || dist > bestDist) continue;
                    if (!bStereo1 && !bStereo2) {
                        const float distex = ep_x - k2[idx2].x, distey = ep_y - k2[idx2].y;
                        if (distex * distex + distey * distey < 100 * scale_factors2[k2[idx2].octave]) continue;
                    }
                    if (epipolar_ok(&k1[idx1], &k2[idx2], F12, level_sigma2_2[k2[idx2].octave]) || coarse) { bestIdx2 = idx2; bestDist = dist; }
                }
                if (bestIdx2 >= 0) {
                    matches12[idx1] = bestIdx2; nmatches++;
                    if (check_ori) { const int bin = rot_bin(k1[idx1].angle, k2[bestIdx2].angle); histIdx[nh] = idx1; histBin[nh] = bin; nh++; hist[bin]++; }
                }
            }
            a++; b++;
        } else if (fv1_nodes[a] < fv2_nodes[b]) {
            while (a < nfv1 && fv1_nodes[a] < fv2_nodes[b]) a++;
        } else {
            while (b < nfv2 && fv2_nodes[b] < fv1_nodes[a]) b++;
        }
    }
    if (check_ori) {
        int ind1, ind2, ind3;
        three_maxima(hist, HISTO_LENGTH, &ind1, &ind2, &ind3);
        for (int j = 0; j < nh; j++)
            if (histBin[j] != ind1 && histBin[j] != ind2 && histBin[j] != ind3) { matches12[histIdx[j]] = -1; nmatches--; }
    }
    free(histIdx); free(histBin);
    (void)n2;
    return nmatches;
}
