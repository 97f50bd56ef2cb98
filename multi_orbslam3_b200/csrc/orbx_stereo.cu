// orbx_stereo.cu - Frame::ComputeStereoMatches (R/src/Frame.cc:785-962) on the device: descriptor search over per-row
// CSR lists (vRowIndices), 11x11 SAD refinement with sub-pixel fit, median outlier cut; single pair, batched, and the
// host-to-host stereo stream pipeline.
#include "orbx_match_internal.h"

namespace {

// ---- Frame::ComputeStereoMatches, batched: blockIdx.y = stereo pair of the batch ----
struct StereoArgs {
    const orbx_keypoint* kL; const uint8_t* dL; const int32_t* nL; int capL;     // left results: [slot][capL]
    const orbx_keypoint* kR; const uint8_t* dR; const int32_t* nR; int capR;     // right results
    int slotL0, slotR0;                 // result slot of pair 0 (pair p uses slot*0 + p); nL/nR == nullptr: counts in nl1/nr1
    int nl1, nr1;
    int nrows;                          // level-0 rows
    float minD, maxD, mbf;
    float sf[ORBX_MAX_LEVELS];          // mvScaleFactors
    int32_t* best_idx; int32_t* best_dist;      // [pair][capL]
    float* uright; float* depth; int32_t* sad;  // [pair][ostride]
    int ostride;
};

// descriptor search (R/src/Frame.cc:785-868), one CTA per stereo pair:
//   1. vRowIndices (:798-812): every right keypoint is listed under the level-0 rows [floor(y - r), ceil(y + r)],
//      r = 2 * scale(octave), as a CSR table (row histogram in shared memory, block scan, fill into `lists`);
//   2. one warp per left keypoint walks the list of its own row (:826-865): octave gate, disparity gate, Hamming distance;
//      best = smallest distance, ties -> smallest right index (the reference visits a row's list in index order).
constexpr int STEREO_NT = 1024;
__device__ __forceinline__ void stereo_row_range(const orbx_keypoint& R, const float* sf, int nrows, int& minr, int& maxr)
{
    const float r = __fmul_rn(2.0f, sf[R.octave]);
    maxr = (int)ceilf(__fadd_rn(R.y, r)); minr = (int)floorf(__fsub_rn(R.y, r));
    if (minr < 0) minr = 0;
    if (maxr > nrows - 1) maxr = nrows - 1;
}

__global__ void __launch_bounds__(STEREO_NT) k_stereo_band(StereoArgs A, int32_t* lists, int list_cap)
{
    extern __shared__ int s_rows[];                 // [nrows + 1] starts, [nrows + 1] fill cursors
    __shared__ int s_warp[STEREO_NT / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, p = blockIdx.x;
    const int nl = A.nL ? A.nL[A.slotL0 + p] : A.nl1, nr = A.nR ? A.nR[A.slotR0 + p] : A.nr1;
    const int nrows = A.nrows;
    int* start = s_rows; int* cur = s_rows + nrows + 1;
    const orbx_keypoint* kl = A.kL + (size_t)(A.slotL0 + p) * A.capL; const uint8_t* dl = A.dL + (size_t)(A.slotL0 + p) * A.capL * 32;
    const orbx_keypoint* kr = A.kR + (size_t)(A.slotR0 + p) * A.capR; const uint8_t* dr = A.dR + (size_t)(A.slotR0 + p) * A.capR * 32;
    int32_t* list = lists + (size_t)p * list_cap;
    for (int i = tid; i <= nrows; i += STEREO_NT) start[i] = 0;
    __syncthreads();
    for (int iR = tid; iR < nr; iR += STEREO_NT) {
        int minr, maxr;
        stereo_row_range(kr[iR], A.sf, nrows, minr, maxr);
        for (int y = minr; y <= maxr; y++) atomicAdd(&start[y], 1);
    }
    __syncthreads();
    {   // exclusive scan of the row histogram: each thread owns a run of consecutive rows
        const int per = (nrows + STEREO_NT) / STEREO_NT;
        const int r0 = tid * per, r1 = min(r0 + per, nrows + 1);
        int sum = 0;
        for (int r = r0; r < r1; r++) sum += start[r];
        int inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int w = s_warp[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
            s_warp[lane] = wi - w;
        }
        __syncthreads();
        int run = s_warp[warp] + inc - sum;
        for (int r = r0; r < r1; r++) { const int c = start[r]; start[r] = run; cur[r] = run; run += c; }
    }
    __syncthreads();
    for (int iR = tid; iR < nr; iR += STEREO_NT) {
        int minr, maxr;
        stereo_row_range(kr[iR], A.sf, nrows, minr, maxr);
        for (int y = minr; y <= maxr; y++) { const int o = atomicAdd(&cur[y], 1); if (o < list_cap) list[o] = iR; }
    }
    __syncthreads();
    for (int iL = warp; iL < nl; iL += STEREO_NT / 32) {
        const orbx_keypoint L = kl[iL];
        const int row = (int)L.y;
        const float minU = __fsub_rn(L.x, A.maxD), maxU = __fsub_rn(L.x, A.minD);
        int bd = ORBX_TH_HIGH, bi = 0x7fffffff;
        if (row >= 0 && row < nrows && !(maxU < 0)) {
            const uint4 q0 = reinterpret_cast<const uint4*>(dl)[2 * iL], q1 = reinterpret_cast<const uint4*>(dl)[2 * iL + 1];
            const int k1 = min(start[row + 1], list_cap);
            for (int k = start[row] + lane; k < k1; k += 32) {
                const int iR = list[k];
                const int oct = kr[iR].octave; const float xr = kr[iR].x;
                if (oct < L.octave - 1 || oct > L.octave + 1) continue;
                if (!(xr >= minU && xr <= maxU)) continue;
                const uint4 t0 = reinterpret_cast<const uint4*>(dr)[2 * iR], t1 = reinterpret_cast<const uint4*>(dr)[2 * iR + 1];
                const int d = hamming256(q0, q1, t0, t1);
                if (d < bd || (d == bd && iR < bi && d < ORBX_TH_HIGH)) { bd = d; bi = iR; }    // strict '<' from TH_HIGH (:829)
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const int od = __shfl_xor_sync(0xffffffffu, bd, o), oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
        }
        if (lane == 0) { A.best_idx[(size_t)p * A.capL + iL] = bi == 0x7fffffff ? -1 : bi; A.best_dist[(size_t)p * A.capL + iL] = bd; }
    }
}

// rows a right keypoint can be listed under: 2 * ceil(2 * largest scale factor) + 3
static int stereo_rows_per_kp(const float* sf, int nlevels)
{
    float mx = 1.0f;
    for (int l = 0; l < nlevels; l++) if (sf[l] > mx) mx = sf[l];
    return 2 * (int)ceilf(2.0f * mx) + 3;
}

// sub-pixel refinement (R/src/Frame.cc:871-946): one warp per left keypoint.
// 11x11 patches around the keypoint (left) and around the matched right keypoint shifted by incR = -5..5, both centred on
// their own middle pixel, L1 distance per shift, parabola through the best shift and its neighbours.
__global__ void __launch_bounds__(256) k_stereo_refine(StereoArgs A, OrbxPyrView L, OrbxPyrView R)
{
    __shared__ int s_d[8][12];
    __shared__ uint8_t s_patch[8][11 * 32];
    const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5, p = blockIdx.y;
    const int iL = blockIdx.x * 8 + wq;
    const int nl = A.nL ? A.nL[A.slotL0 + p] : A.nl1;
    if (iL >= nl) return;
    const orbx_keypoint* kl = A.kL + (size_t)(A.slotL0 + p) * A.capL;
    const orbx_keypoint* kr = A.kR + (size_t)(A.slotR0 + p) * A.capR;
    float out_u = -1.0f, out_z = -1.0f; int out_s = -1;
    const int bi = A.best_idx[(size_t)p * A.capL + iL];
    const int thOrbDist = (ORBX_TH_HIGH + ORBX_TH_LOW) / 2;
    if (bi >= 0 && A.best_dist[(size_t)p * A.capL + iL] < thOrbDist) {
        const orbx_keypoint kp = kl[iL];
        const int oct = kp.octave;
        const float uL = kp.x;
        const float uR0 = kr[bi].x;
        const float sf = L.inv_scale[oct];
        const float scaleduL = roundf(__fmul_rn(kp.x, sf)), scaledvL = roundf(__fmul_rn(kp.y, sf)), scaleduR0 = roundf(__fmul_rn(uR0, sf));
        const int w = 5, Lr = 5;
        const float iniu = scaleduR0 + Lr - w, endu = scaleduR0 + Lr + w + 1;
        if (!(iniu < 0 || endu >= (float)R.w[oct])) {
            const int r0 = (int)(scaledvL - w), c0 = (int)(scaleduL - w), cr0 = (int)(scaleduR0 - w);
            const uint8_t* imL = L.lv[oct] + (long long)p * L.fstride[oct]; const int pL = L.pitch[oct];
            const uint8_t* imR = R.lv[oct] + (long long)p * R.fstride[oct]; const int pR = R.pitch[oct];
            if (lane < 11) s_d[wq][lane] = 0;
            // stage the 11x11 left patch and the 11x21 right strip (all 11 shifts) in shared memory: 32 bytes per row
            uint8_t* sp = s_patch[wq];
#pragma unroll
            for (int r = 0; r < 11; r++)
                sp[r * 32 + lane] = lane < 11 ? imL[(long long)(r0 + r) * pL + c0 + lane]
                                              : imR[(long long)(r0 + r) * pR + cr0 - Lr + (lane - 11)];
            __syncwarp();
            const int ctrL = sp[w * 32 + w];
            // 121 (shift, row) tasks of 11 pixels each: |(a - ctrL) - (b - ctrR)| = |(a + ctrR - ctrL) - b|
            for (int t = lane; t < 121; t += 32) {
                const int inc = t / 11, r = t - inc * 11;
                const int kd = (int)sp[w * 32 + 11 + inc + w] - ctrL;
                const uint8_t* a = sp + r * 32;
                const uint8_t* b = sp + r * 32 + 11 + inc;
                unsigned acc = 0;
#pragma unroll
                for (int c = 0; c < 11; c++) acc = __sad((int)a[c] + kd, (int)b[c], acc);
                atomicAdd(&s_d[wq][inc], (int)acc);
            }
            __syncwarp();
            int bestDist = 0x7fffffff, bestinc = 0;
            for (int k = 0; k < 11; k++) { const int d = s_d[wq][k]; if (d < bestDist) { bestDist = d; bestinc = k - Lr; } }   // first minimum wins (:905-909)
            if (bestinc != -Lr && bestinc != Lr) {
                const float d1 = (float)s_d[wq][Lr + bestinc - 1], d2 = (float)s_d[wq][Lr + bestinc], d3 = (float)s_d[wq][Lr + bestinc + 1];
                const float deltaR = __fdiv_rn(__fsub_rn(d1, d3), __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2))));
                if (!(deltaR < -1 || deltaR > 1)) {
                    float bestuR = __fmul_rn(L.scale[oct], __fadd_rn(__fadd_rn(scaleduR0, (float)bestinc), deltaR));
                    float disparity = __fsub_rn(uL, bestuR);
                    if (disparity >= A.minD && disparity < A.maxD) {
                        if (disparity <= 0) { disparity = 0.01f; bestuR = (float)((double)uL - 0.01); }
                        out_z = __fdiv_rn(A.mbf, disparity); out_u = bestuR; out_s = bestDist;
                    }
                }
            }
        }
    }
    if (lane == 0) {
        const size_t o = (size_t)p * A.ostride + iL;
        A.uright[o] = out_u; A.depth[o] = out_z; A.sad[o] = out_s;
    }
}

// median-based outlier cut (R/src/Frame.cc:949-962): one CTA per pair; bitonic sort of the SAD distances of the matched keypoints
__global__ void __launch_bounds__(1024) k_stereo_outliers(StereoArgs A, int npad)
{
    extern __shared__ int s_v[];
    __shared__ int s_n;
    const int p = blockIdx.x;
    const int nl = A.nL ? A.nL[A.slotL0 + p] : A.nl1;
    float* uright = A.uright + (size_t)p * A.ostride; float* depth = A.depth + (size_t)p * A.ostride;
    const int32_t* sad = A.sad + (size_t)p * A.ostride;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    int local = 0;
    for (int i = threadIdx.x; i < nl; i += 1024) local += sad[i] >= 0;
    if (local) atomicAdd(&s_n, local);
    for (int i = threadIdx.x; i < npad; i += 1024) s_v[i] = (i < nl && sad[i] >= 0) ? sad[i] : 0x7fffffff;
    __syncthreads();
    const int n = s_n;
    if (n == 0) return;
    for (int k = 2; k <= npad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (npad >> 1); t += 1024) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
                const bool up = (i & k) == 0;
                const int x = s_v[i], y = s_v[l];
                if ((x > y) == up) { s_v[i] = y; s_v[l] = x; }
            }
            __syncthreads();
        }
    const float median = (float)s_v[n / 2];
    const float thDist = __fmul_rn(1.5f * 1.4f, median);
    for (int i = threadIdx.x; i < nl; i += 1024)
        if (sad[i] >= 0 && !((float)sad[i] < thDist)) { uright[i] = -1.0f; depth[i] = -1.0f; }
}

}  // namespace

static int stereo_scratch(orbx_matcher* m, size_t bytes);

extern "C" int orbx_stereo_band_match(orbx_matcher* m, const orbx_keypoint* kl, const uint8_t* dl, int nl,
                                      const orbx_keypoint* kr, const uint8_t* dr, int nr, const float* scale_factors,
                                      int nlevels, int nrows, float min_d, float max_d, int32_t* best_idx, int32_t* best_dist)
{
    if (!m || nl < 0 || nr < 0 || nl > m->K || nr > m->K || nrows < 1 || nlevels < 1 || nlevels > ORBX_MAX_LEVELS || !scale_factors ||
        (nl > 0 && (!kl || !dl || !best_idx || !best_dist)) || (nr > 0 && (!kr || !dr))) return ORBX_E_INVALID;
    if (nl == 0) return ORBX_OK;
    CKM(cudaSetDevice(m->p.device));
    cudaStream_t s = m->stream;
    CKM(cudaMemcpyAsync(m->d_k1, kl, sizeof(orbx_keypoint) * nl, cudaMemcpyHostToDevice, s));
    CKM(cudaMemcpyAsync(m->d_d1, dl, (size_t)32 * nl, cudaMemcpyHostToDevice, s));
    if (nr) {
        CKM(cudaMemcpyAsync(m->d_k2, kr, sizeof(orbx_keypoint) * nr, cudaMemcpyHostToDevice, s));
        CKM(cudaMemcpyAsync(m->d_d2, dr, (size_t)32 * nr, cudaMemcpyHostToDevice, s));
    }
    {
        StereoArgs A{};
        A.kL = m->d_k1; A.dL = m->d_d1; A.capL = nl; A.kR = m->d_k2; A.dR = m->d_d2; A.capR = nr; A.nl1 = nl; A.nr1 = nr;
        A.nrows = nrows; A.minD = min_d; A.maxD = max_d;
        for (int l = 0; l < nlevels && l < ORBX_MAX_LEVELS; l++) A.sf[l] = scale_factors[l];
        A.best_idx = m->d_out; A.best_dist = m->d_out2;
        const int list_cap = (nr > 0 ? nr : 1) * stereo_rows_per_kp(A.sf, nlevels < ORBX_MAX_LEVELS ? nlevels : ORBX_MAX_LEVELS);
        int rc = stereo_scratch(m, sizeof(int32_t) * (size_t)list_cap);
        if (rc) return rc;
        const size_t smem = sizeof(int) * 2 * ((size_t)nrows + 1);
        if (smem > 200 * 1024) return ORBX_E_INVALID;
        if (smem > 48 * 1024) CKM(ORBX_OPTIN_SMEM(k_stereo_band));
        k_stereo_band<<<1, STEREO_NT, smem, s>>>(A, reinterpret_cast<int32_t*>(m->d_st), list_cap); ORBX_COUNT_LAUNCH(1);
    }
    CKM(cudaGetLastError());
    CKM(cudaMemcpyAsync(best_idx, m->d_out, sizeof(int32_t) * nl, cudaMemcpyDeviceToHost, s));
    CKM(cudaMemcpyAsync(best_dist, m->d_out2, sizeof(int32_t) * nl, cudaMemcpyDeviceToHost, s));
    CKM(cudaStreamSynchronize(s));
    return ORBX_OK;
}

// Frame::ComputeStereoMatches (R/src/Frame.cc:785-962) on two extractors' device-resident results and pyramids
// Device scratch of one stereo call: row lists, best index / distance, and (when the caller's outputs live on the host)
// mvuRight / mvDepth / SAD rows.
struct StereoScratch { int32_t* lists; int list_cap; int32_t* best; float* u; float* z; int32_t* sad; };
static int stereo_prepare(orbx_matcher* m, orbx_extractor* left, orbx_extractor* right, int count, StereoScratch* S)
{
    const int capL = orbx_ex_out_cap(left), capR = orbx_ex_out_cap(right);
    OrbxPyrView v;
    int rc = orbx_ex_pyramid_view(left, 0, &v);
    if (rc) return rc;
    S->list_cap = capR * stereo_rows_per_kp(v.scale, v.nlevels);
    const size_t n_lists = (size_t)count * S->list_cap, n_row = (size_t)count * capL;
    if ((rc = stereo_scratch(m, sizeof(int32_t) * (n_lists + 5 * n_row)))) return rc;
    S->lists = reinterpret_cast<int32_t*>(m->d_st);
    S->best = S->lists + n_lists;
    S->u = reinterpret_cast<float*>(S->best + 2 * n_row); S->z = S->u + n_row;
    S->sad = reinterpret_cast<int32_t*>(S->z + n_row);
    return ORBX_OK;
}

// Launches the three stereo kernels for `count` pairs: pair p = (left slot slot_l + p, frame frame_l + p) x (right ...).
// Outputs have row stride `ostride`.
static int stereo_launch(orbx_matcher* m, orbx_extractor* left, orbx_extractor* right, int slot_l, int slot_r, int frame_l, int frame_r,
                         int count, float mb, float mbf, const StereoScratch& S, float* d_u, float* d_z, int32_t* d_sad, int ostride,
                         cudaStream_t s)
{
    orbx_keypoint *kL, *kR; uint8_t *dL, *dR; int32_t *nL, *nR; int capL, capR, slotsL, slotsR;
    int rc = orbx_extractor_results_device(left, &kL, &dL, &nL, nullptr, &capL, &slotsL);
    if (rc) return rc;
    if ((rc = orbx_extractor_results_device(right, &kR, &dR, &nR, nullptr, &capR, &slotsR))) return rc;
    if (count <= 0 || slot_l < 0 || slot_l + count > slotsL || slot_r < 0 || slot_r + count > slotsR || ostride < capL) return ORBX_E_INVALID;
    OrbxPyrView vL, vR, tmp;
    if ((rc = orbx_ex_pyramid_view(left, frame_l, &vL)) || (rc = orbx_ex_pyramid_view(right, frame_r, &vR))) return rc;
    if ((rc = orbx_ex_pyramid_view(left, frame_l + count - 1, &tmp)) || (rc = orbx_ex_pyramid_view(right, frame_r + count - 1, &tmp))) return rc;
    StereoArgs A{};
    A.kL = kL; A.dL = dL; A.nL = nL; A.capL = capL; A.kR = kR; A.dR = dR; A.nR = nR; A.capR = capR;
    A.slotL0 = slot_l; A.slotR0 = slot_r;
    A.nrows = vL.h[0]; A.minD = 0.0f; A.maxD = mbf / mb; A.mbf = mbf;      // minZ = mb (R/src/Frame.cc:815-818)
    for (int l = 0; l < vL.nlevels; l++) A.sf[l] = vL.scale[l];
    A.best_idx = S.best; A.best_dist = S.best + (size_t)count * capL;
    A.uright = d_u; A.depth = d_z; A.sad = d_sad; A.ostride = ostride;
    {
        const size_t smem = sizeof(int) * 2 * ((size_t)A.nrows + 1);
        if (smem > 200 * 1024) return ORBX_E_INVALID;
        if (smem > 48 * 1024) CKM(ORBX_OPTIN_SMEM(k_stereo_band));
        k_stereo_band<<<count, STEREO_NT, smem, s>>>(A, S.lists, S.list_cap); ORBX_COUNT_LAUNCH(1);
    }
    k_stereo_refine<<<dim3((capL + 7) / 8, count), 256, 0, s>>>(A, vL, vR); ORBX_COUNT_LAUNCH(1);
    int npad = 1; while (npad < capL) npad <<= 1;
    if (sizeof(int) * npad > 48 * 1024) CKM(ORBX_OPTIN_SMEM(k_stereo_outliers));
    k_stereo_outliers<<<count, 1024, sizeof(int) * npad, s>>>(A, npad); ORBX_COUNT_LAUNCH(1);
    CKM(cudaGetLastError());
    return ORBX_OK;
}

static int stereo_scratch(orbx_matcher* m, size_t bytes)
{
    if (bytes <= m->st_bytes) return ORBX_OK;
    if (m->d_st) { cudaDeviceSynchronize(); cudaFree(m->d_st); }
    m->d_st = nullptr; m->st_bytes = 0;
    CKM(cudaMalloc((void**)&m->d_st, bytes));
    m->st_bytes = bytes;
    return ORBX_OK;
}

extern "C" int orbx_stereo_matches(orbx_matcher* m, orbx_extractor* left, orbx_extractor* right, int slot_l, int slot_r,
                                   int frame_l, int frame_r, float mb, float mbf, float* uright, float* depth,
                                   int32_t* sad_dist, int cap, int* n_left)
{
    if (!m || !left || !right || !uright || !depth) return ORBX_E_INVALID;
    orbx_keypoint* kL; uint8_t* dL; int32_t* nL; int capL, slotsL;
    int rc = orbx_extractor_results_device(left, &kL, &dL, &nL, nullptr, &capL, &slotsL);
    if (rc) return rc;
    if (slot_l < 0 || slot_l >= slotsL) return ORBX_E_INVALID;
    CKM(cudaSetDevice(m->p.device));
    // the extractors run on their own streams: wait for both, then work on the matcher's stream
    CKM(cudaStreamSynchronize(orbx_ex_stream(left)));
    CKM(cudaStreamSynchronize(orbx_ex_stream(right)));
    cudaStream_t s = m->stream;
    int nl = 0;
    CKM(cudaMemcpyAsync(&nl, nL + slot_l, sizeof(int), cudaMemcpyDeviceToHost, s));
    CKM(cudaStreamSynchronize(s));
    if (n_left) *n_left = nl;
    if (nl > cap) { orbx_set_error("%s%s", "orbx_stereo_matches: output capacity too small", ""); return ORBX_E_CAPACITY; }
    if (nl == 0) return ORBX_OK;
    StereoScratch S;
    if ((rc = stereo_prepare(m, left, right, 1, &S))) return rc;
    if ((rc = stereo_launch(m, left, right, slot_l, slot_r, frame_l, frame_r, 1, mb, mbf, S, S.u, S.z, S.sad, capL, s))) return rc;
    CKM(cudaMemcpyAsync(uright, S.u, sizeof(float) * nl, cudaMemcpyDeviceToHost, s));
    CKM(cudaMemcpyAsync(depth, S.z, sizeof(float) * nl, cudaMemcpyDeviceToHost, s));
    if (sad_dist) CKM(cudaMemcpyAsync(sad_dist, S.sad, sizeof(int32_t) * nl, cudaMemcpyDeviceToHost, s));
    CKM(cudaStreamSynchronize(s));
    return ORBX_OK;
}

// batched form on device buffers: pair p = slot / frame (first + p) of both extractors; everything is enqueued on `stream`
// (the stream the two orbx_extract_batch_device calls used); outputs are [count][capacity of the left extractor]
extern "C" int orbx_stereo_matches_batch_device(orbx_matcher* m, orbx_extractor* left, orbx_extractor* right, int first, int count,
                                                float mb, float mbf, float* d_uright, float* d_depth, int32_t* d_sad, void* stream)
{
    if (!m || !left || !right || !d_uright || !d_depth || count <= 0) return ORBX_E_INVALID;
    CKM(cudaSetDevice(m->p.device));
    cudaStream_t s = stream ? (cudaStream_t)stream : m->stream;
    StereoScratch S;
    int rc = stereo_prepare(m, left, right, count, &S);
    if (rc) return rc;
    return stereo_launch(m, left, right, first, first, first, first, count, mb, mbf, S, d_uright, d_depth, d_sad ? d_sad : S.sad,
                         orbx_ex_out_cap(left), s);
}

// batched form with host outputs: uright/depth are [count][cap] (rows beyond a frame's keypoint count are unspecified)
extern "C" int orbx_stereo_matches_batch(orbx_matcher* m, orbx_extractor* left, orbx_extractor* right, int first, int count,
                                         float mb, float mbf, float* uright, float* depth, int cap)
{
    if (!m || !left || !right || !uright || !depth || count <= 0 || cap <= 0) return ORBX_E_INVALID;
    CKM(cudaSetDevice(m->p.device));
    CKM(cudaStreamSynchronize(orbx_ex_stream(left)));
    CKM(cudaStreamSynchronize(orbx_ex_stream(right)));
    cudaStream_t s = m->stream;
    const int capL = orbx_ex_out_cap(left);
    StereoScratch S;
    int rc = stereo_prepare(m, left, right, count, &S);
    if (rc) return rc;
    if ((rc = stereo_launch(m, left, right, first, first, first, first, count, mb, mbf, S, S.u, S.z, S.sad, capL, s))) return rc;
    const int wcopy = cap < capL ? cap : capL;
    CKM(cudaMemcpy2DAsync(uright, sizeof(float) * cap, S.u, sizeof(float) * capL, sizeof(float) * wcopy, count, cudaMemcpyDeviceToHost, s));
    CKM(cudaMemcpy2DAsync(depth, sizeof(float) * cap, S.z, sizeof(float) * capL, sizeof(float) * wcopy, count, cudaMemcpyDeviceToHost, s));
    CKM(cudaStreamSynchronize(s));
    return ORBX_OK;
}

// One call = a batch of stereo frames from host memory to host results: both cameras extracted (Frame.cc:92-95 runs the
// two extractors side by side) and Frame::ComputeStereoMatches for every pair.  The batch is cut into chunks that flow
// through five streams (H2D | left extractor | right extractor | stereo kernels | D2H), so the copies and the stereo
// kernels of one chunk hide under the extraction of its neighbours.
extern "C" int orbx_extract_stereo_batch(orbx_matcher* m, orbx_extractor* left, orbx_extractor* right,
                                         const uint8_t* imgs_left, const uint8_t* imgs_right, int batch, int width, int height,
                                         int stride, size_t frame_stride, float mb, float mbf,
                                         orbx_keypoint* kps_l, uint8_t* desc_l, int32_t* n_l,
                                         orbx_keypoint* kps_r, uint8_t* desc_r, int32_t* n_r, int cap,
                                         float* uright, float* depth)
{
    if (!m || !left || !right || left == right || !imgs_left || !imgs_right || batch < 1 || !uright || !depth || cap <= 0) return ORBX_E_INVALID;
    if (width <= 0 || height <= 0) return ORBX_E_EMPTY;
    int rc;
    if ((rc = orbx_ex_configure(left, width, height)) || (rc = orbx_ex_configure(right, width, height))) return rc;
    if (orbx_ex_device(left) != m->p.device || orbx_ex_device(right) != m->p.device) return ORBX_E_INVALID;
    int slotsL = 0, slotsR = 0, capL = 0, capR = 0;
    { orbx_keypoint* k; uint8_t* d; int32_t* n;
      if ((rc = orbx_extractor_results_device(left, &k, &d, &n, nullptr, &capL, &slotsL)) || (rc = orbx_extractor_results_device(right, &k, &d, &n, nullptr, &capR, &slotsR))) return rc; }
    if (batch > slotsL - 1 || batch > slotsR - 1) { orbx_set_error("%s%s", "orbx_extract_stereo_batch: batch larger than max_batch of an extractor", ""); return ORBX_E_INVALID; }
    CKM(cudaSetDevice(m->p.device));
    if ((rc = orbx_m_ensure_pipeline(m))) return rc;
    cudaStream_t sL = orbx_ex_stream(left), sR = orbx_ex_stream(right);
    if (m->mono2_cap < batch) {
        if (m->h_mono2) cudaFreeHost(m->h_mono2);
        m->h_mono2 = nullptr; m->mono2_cap = 0;
        CKM(cudaMallocHost((void**)&m->h_mono2, sizeof(int32_t) * 2 * (size_t)batch));
        m->mono2_cap = batch;
    }
    int32_t* mono_l = m->h_mono2; int32_t* mono_r = m->h_mono2 + batch;      // monoIndex is not part of this call's results
    const bool directL = orbx_ex_can_fetch_direct(left, kps_l, desc_l, cap, n_l, mono_l), directR = orbx_ex_can_fetch_direct(right, kps_r, desc_r, cap, n_r, mono_r);
    int nchunks = batch >= 48 ? 6 : (batch >= 16 ? 4 : 1);      // measured on C2 / C3: 6 chunks 88.9 k / 62.0 k frames/s, 4: 87.6 / 61.2, 8: 86.6 / 59.1
    if (const char* e = getenv("ORBX_HOST_CHUNKS")) { const int v = atoi(e); if (v >= 1 && v <= ORBX_MAX_CHUNKS) nchunks = v; }
    if (nchunks > batch) nchunks = batch;
    const int per = (batch + nchunks - 1) / nchunks;
    nchunks = (batch + per - 1) / per;
    StereoScratch S{};
    // the side streams start after whatever the caller queued on the extractors' streams
    CKM(cudaEventRecord(m->ev_start, sL));
    CKM(cudaStreamWaitEvent(m->s_h2d, m->ev_start, 0));
    CKM(cudaStreamWaitEvent(m->s_match, m->ev_start, 0));
    CKM(cudaStreamWaitEvent(m->s_d2h, m->ev_start, 0));
    for (int c = 0; c < nchunks; c++) {
        const int f0 = c * per, cnt = f0 + per <= batch ? per : batch - f0;
        if ((rc = orbx_ex_stage_input(left, imgs_left, f0, cnt, width, height, stride, frame_stride, m->s_h2d))) return rc;
        CKM(cudaEventRecord(m->ev[c], m->s_h2d));
        if ((rc = orbx_ex_stage_input(right, imgs_right, f0, cnt, width, height, stride, frame_stride, m->s_h2d))) return rc;
        CKM(cudaEventRecord(m->ev_r[c], m->s_h2d));
    }
    for (int c = 0; c < nchunks; c++) {
        const int f0 = c * per, cnt = f0 + per <= batch ? per : batch - f0;
        CKM(cudaStreamWaitEvent(sL, m->ev[c], 0));
        if ((rc = orbx_ex_run_staged(left, f0, cnt, 0, 0, f0, sL))) return rc;
        CKM(cudaEventRecord(m->ev_ext[c], sL));
        CKM(cudaStreamWaitEvent(sR, m->ev_r[c], 0));
        if ((rc = orbx_ex_run_staged(right, f0, cnt, 0, 0, f0, sR))) return rc;
        CKM(cudaEventRecord(m->ev_r[ORBX_MAX_CHUNKS + c], sR));
        if (c == 0 && (rc = stereo_prepare(m, left, right, batch, &S))) return rc;      // needs the geometry of a queued batch
        CKM(cudaStreamWaitEvent(m->s_match, m->ev_ext[c], 0));
        CKM(cudaStreamWaitEvent(m->s_match, m->ev_r[ORBX_MAX_CHUNKS + c], 0));
        StereoScratch C = S;                                   // this chunk's slice of the scratch
        C.lists = S.lists + (size_t)f0 * S.list_cap; C.best = S.best + 2 * (size_t)f0 * capL;
        float* du = S.u + (size_t)f0 * capL; float* dz = S.z + (size_t)f0 * capL; int32_t* ds = S.sad + (size_t)f0 * capL;
        if ((rc = stereo_launch(m, left, right, f0, f0, f0, f0, cnt, mb, mbf, C, du, dz, ds, capL, m->s_match))) return rc;
        CKM(cudaEventRecord(m->ev[ORBX_MAX_CHUNKS + c], m->s_match));
        CKM(cudaStreamWaitEvent(m->s_d2h, m->ev[ORBX_MAX_CHUNKS + c], 0));
        if ((rc = orbx_ex_fetch_async(left, f0, cnt, f0, kps_l, desc_l, cap, n_l, mono_l, m->s_d2h, directL))) return rc;
        if ((rc = orbx_ex_fetch_async(right, f0, cnt, f0, kps_r, desc_r, cap, n_r, mono_r, m->s_d2h, directR))) return rc;
        const int wcopy = cap < capL ? cap : capL;
        CKM(cudaMemcpy2DAsync(uright + (size_t)f0 * cap, sizeof(float) * cap, du, sizeof(float) * capL, sizeof(float) * wcopy, cnt, cudaMemcpyDeviceToHost, m->s_d2h));
        CKM(cudaMemcpy2DAsync(depth + (size_t)f0 * cap, sizeof(float) * cap, dz, sizeof(float) * capL, sizeof(float) * wcopy, cnt, cudaMemcpyDeviceToHost, m->s_d2h));
    }
    CKM(cudaStreamSynchronize(m->s_d2h));
    CKM(cudaStreamSynchronize(sL));
    CKM(cudaStreamSynchronize(sR));
    if ((rc = orbx_ex_fetch_finish(left, batch, kps_l, desc_l, cap, n_l, nullptr, directL))) return rc;
    return orbx_ex_fetch_finish(right, batch, kps_r, desc_r, cap, n_r, nullptr, directR);
}

