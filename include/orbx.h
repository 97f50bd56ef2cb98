/* orbx.h - C ABI of liborbx_b200.so: the B200-native (sm_100a) ORB front-end hot path.
 *
 * Drop-in boundary for multi_orbslam3's ORBextractor / ORBmatcher.  The reference has no FFI or
 * plugin layer: both are concrete C++ classes.  The replacement keeps the two
 * class headers unchanged (dropin/ORBextractor.h, dropin/ORBmatcher.h) and routes their bodies to
 * the entry points below.  Every entry point cites the reference code it replaces;
 * R/ = src/orb_slam3_ros/orb_slam3/ in the reference tree.
 *
 * Groups: extractor (orbx_extractor_*, orbx_extract*), pyramid access and test taps, matcher (Hamming pairs, brute-force
 * kNN-2, SearchForInitialization, SearchByProjection family, SearchByBoW, SearchForTriangulation, candidate lists),
 * stereo (orbx_stereo_*, orbx_extract_stereo_batch), stream pipelines (orbx_extract_match_batch*), bag of words
 * (orbx_vocab_*, orbx_bow_transform*), frame / map-point helpers (undistortion, distinctive descriptors, KF.msg records).
 *
 * Conventions: extern "C", plain pointers and sizes, int status return (ORBX_OK == 0), no
 * exceptions, opaque handles, caller-allocated outputs with a capacity and a count-out.
 * `stream` arguments are a cudaStream_t passed as void* (NULL = the handle's own stream).
 * There is no CPU fallback: without a CUDA device every create call fails with ORBX_E_CUDA.
 */
#ifndef ORBX_H
#define ORBX_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORBX_OK            0
#define ORBX_E_INVALID    -1   /* bad argument */
#define ORBX_E_EMPTY      -2   /* empty image: ORBextractor::operator() returns -1 (R/src/ORBextractor.cc:1072) */
#define ORBX_E_CUDA       -3   /* CUDA runtime error; see orbx_last_error() */
#define ORBX_E_CAPACITY   -4   /* an internal or caller buffer was too small; nothing was truncated silently */
#define ORBX_E_NOMEM      -5

#define ORBX_MAX_LEVELS   12
#define ORBX_TH_HIGH      100  /* ORBmatcher::TH_HIGH  (R/src/ORBmatcher.cc:36) */
#define ORBX_TH_LOW       50   /* ORBmatcher::TH_LOW   (R/src/ORBmatcher.cc:37) */
#define ORBX_HISTO_LENGTH 30   /* ORBmatcher::HISTO_LENGTH (R/src/ORBmatcher.cc:38) */
#define ORBX_GRID_COLS    64   /* FRAME_GRID_COLS (R/include/Frame.h:39) */
#define ORBX_GRID_ROWS    48   /* FRAME_GRID_ROWS (R/include/Frame.h:38) */

/* binary-compatible with cv::KeyPoint (28 bytes): pt.x pt.y size angle response octave class_id */
typedef struct orbx_keypoint {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} orbx_keypoint;

const char* orbx_last_error(void);          /* thread-local, human readable */
int  orbx_device_count(void);
/* number of CUDA kernels this library has launched since it was loaded (bench.py: gpu_launches) */
unsigned long long orbx_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Extractor: replaces class ORBextractor (R/include/ORBextractor.h:47-113).
 * One handle per reference ORBextractor instance; like the reference it is not re-entrant, distinct
 * handles run concurrently (R/src/Frame.cc:92-95 runs left/right extractors on two threads).
 * ---------------------------------------------------------------------------------------------- */
typedef struct orbx_extractor orbx_extractor;

typedef struct orbx_params {
    int32_t nfeatures;        /* ORBextractor ctor args, R/src/ORBextractor.cc:408-411 */
    float   scale_factor;
    int32_t nlevels;
    int32_t ini_th_fast;
    int32_t min_th_fast;
    int32_t max_width;        /* largest image this handle will see */
    int32_t max_height;
    int32_t max_batch;        /* frames per batched call (1 for the class wrapper) */
    int32_t device;           /* CUDA device ordinal */
    int32_t max_candidates_per_level; /* 0 = default (16384); FAST corners per level above this -> ORBX_E_CAPACITY */
} orbx_params;

/* ORBextractor::ORBextractor, R/src/ORBextractor.cc:408-468 */
int  orbx_extractor_create(const orbx_params* p, orbx_extractor** out);
void orbx_extractor_destroy(orbx_extractor* h);

/* scale tables / per-level quotas: R/include/ORBextractor.h:64-86 getters, R/src/ORBextractor.cc:413-444.
 * Each out array has nlevels entries; NULL pointers are skipped. */
int  orbx_extractor_tables(const orbx_extractor* h, float* scale, float* inv_scale, float* sigma2,
                           float* inv_sigma2, int32_t* features_per_level);
/* upper bound on keypoints per frame (nfeatures + 3 per level): size outputs with it */
int  orbx_extractor_max_keypoints(const orbx_extractor* h);

/* ORBextractor::operator(), R/src/ORBextractor.cc:1068-1150.  Synchronous, one host image.
 * lap0/lap1 = vLappingArea.  *mono_index receives the return value of operator() (monoIndex).
 * Returns ORBX_E_EMPTY for an empty image (the class wrapper maps it to -1).  kps and / or desc may be NULL when only
 * the count is wanted. */
int  orbx_extract(orbx_extractor* h, const uint8_t* img, int width, int height, int stride,
                  int lap0, int lap1, orbx_keypoint* kps, uint8_t* desc, int cap, int* n, int* mono_index);

/* Batched operator() over host frames (the e2e path): `batch` frames of equal geometry at
 * imgs + i*frame_stride; H2D and D2H copies happen inside.  kps: batch*cap, desc: batch*cap*32. */
int  orbx_extract_batch(orbx_extractor* h, const uint8_t* imgs, int batch, int width, int height,
                        int stride, size_t frame_stride, int lap0, int lap1,
                        orbx_keypoint* kps, uint8_t* desc, int cap, int32_t* n, int32_t* mono_index);

/* Batched operator() over DEVICE-resident frames, asynchronous on `stream`; results stay on the
 * device in the handle's result slots first_slot .. first_slot+batch-1 (slots 0..max_batch). */
int  orbx_extract_batch_device(orbx_extractor* h, const uint8_t* d_imgs, int batch, int width, int height,
                               int stride, size_t frame_stride, int lap0, int lap1, int first_slot, void* stream);
/* copies result slot `from` onto slot `to` (carry the last frame of a batch over as predecessor) */
int  orbx_extractor_copy_slot(orbx_extractor* h, int from, int to, void* stream);
/* device views of the result slots: kps [slots][cap], desc [slots][cap][32], n [slots], mono [slots] */
int  orbx_extractor_results_device(orbx_extractor* h, orbx_keypoint** d_kps, uint8_t** d_desc,
                                   int32_t** d_n, int32_t** d_mono, int* cap, int* slots);
/* D2H of `count` result slots starting at first_slot; synchronises `stream`; reports ORBX_E_CAPACITY
 * raised by any kernel of the batch. */
int  orbx_extractor_download(orbx_extractor* h, int first_slot, int count, orbx_keypoint* kps, uint8_t* desc,
                             int cap, int32_t* n, int32_t* mono_index, void* stream);
int  orbx_extractor_sync(orbx_extractor* h, void* stream);   /* also returns any deferred device error */
/* per-stage CUDA-event timing: returns the milliseconds accumulated so far in stage_ms4 = {pyramid+blur, FAST,
 * octree, finalize+orientation+descriptors} over *batches batches; enable = 1/0 switches it on/off and resets the
 * sums, -1 only reads. */
int  orbx_extractor_profile(orbx_extractor* h, int enable, double* stage_ms4, int* batches);

/* Host-only introspection (no device needed): the FAST cell grid of a pyramid level of width level_width
 * (R/src/ORBextractor.cc:779-785: nCols = width/30, wCell = ceil(width/nCols) over the 16-px-bordered area) and how the
 * kernel cuts a cell row into segments of whole cells (cells per segment, segments per row, widest staged tile). */
int  orbx_fast_segment_plan(int level_width, int* n_cols, int* w_cell, int* seg_cells, int* n_seg, int* max_staged_width);

/* mvImagePyramid (R/include/ORBextractor.h:88) stays on the device; this is the explicit download the
 * stereo SAD refinement (R/src/Frame.cc:882-901) needs.  slot = frame within the last batch. */
int  orbx_pyramid_level_size(const orbx_extractor* h, int level, int* width, int* height);
int  orbx_pyramid_to_host(orbx_extractor* h, int slot, int level, uint8_t* dst, int dst_stride);
/* Levels first_level .. first_level + n_levels - 1 of one frame in ONE round trip: the copies are queued into pinned staging,
 * one synchronisation, then rows are unpacked into dst[k] (row stride dst_stride[k]) = what ORBextractor::operator() does
 * after every frame to leave mvImagePyramid as the reference does (R/src/ORBextractor.cc:1150-1177). */
int  orbx_pyramid_levels_to_host(orbx_extractor* h, int slot, int first_level, int n_levels, uint8_t* const* dst, const int* dst_stride);
/* The same without the unpack: the levels land in pinned staging owned by the handle and ptr[k] / stride[k] point into it
 * (valid until the next call of this function or of orbx_pyramid_levels_to_host on the handle, or its destruction): the class
 * layer wraps them in cv::Mat headers, as short-lived as the reference's own mvImagePyramid entries (rewritten every frame). */
int  orbx_pyramid_levels_staged(orbx_extractor* h, int slot, int first_level, int n_levels, const uint8_t** ptr, int* stride);
/* test taps: blurred level (R/src/ORBextractor.cc:1114-1115) and the FAST candidates handed to
 * DistributeOctTree (R/src/ORBextractor.cc:845-851) as (x,y,response) float triples */
int  orbx_blurred_to_host(orbx_extractor* h, int slot, int level, uint8_t* dst, int dst_stride);
int  orbx_candidates_to_host(orbx_extractor* h, int slot, int level, float* xyr, int cap, int* n);
int  orbx_level_keypoints_to_host(orbx_extractor* h, int slot, int level, float* xyr, int cap, int* n);

/* ------------------------------------------------------------------------------------------------
 * Matcher: replaces the Hamming searches of class ORBmatcher (R/include/ORBmatcher.h:35-108).
 * ---------------------------------------------------------------------------------------------- */
typedef struct orbx_matcher orbx_matcher;

typedef struct orbx_matcher_params {
    int32_t device;
    int32_t max_keypoints;     /* per frame */
    int32_t max_batch;         /* frame pairs per batched call */
    int32_t max_candidates;    /* CSR pool per pair for windowed searches; 0 = default (131072) */
} orbx_matcher_params;

int  orbx_matcher_create(const orbx_matcher_params* p, orbx_matcher** out);
void orbx_matcher_destroy(orbx_matcher* m);
int  orbx_matcher_sync(orbx_matcher* m, void* stream);

/* ORBmatcher::DescriptorDistance (R/src/ORBmatcher.cc:2358-2374) for n descriptor pairs:
 * out[i] = hamming(a[i], b[i]); host pointers. */
int  orbx_hamming_pairs(orbx_matcher* m, const uint8_t* a, const uint8_t* b, int n, int32_t* out);

/* cv::BFMatcher(NORM_HAMMING).knnMatch(q, t, 2) as used at R/src/Frame.cc:1127-1137 and by the server's
 * cross-agent matching: top-2 by (distance, trainIdx).  idx/dist are nq x 2; missing = -1.  Host pointers. */
int  orbx_bf_knn2(orbx_matcher* m, const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx, int32_t* dist);
/* same on device pointers, asynchronous.  idx_base is added to every train index (sharded DBs). */
int  orbx_bf_knn2_device(orbx_matcher* m, const uint8_t* d_q, int nq, const uint8_t* d_t, long long nt,
                         int32_t* d_idx, int32_t* d_dist, int idx_base, void* stream);
/* merges nparts partial top-2 tables ([part][nq][2], e.g. all-gathered from the DB shards of the 8 GPUs)
 * into the global top-2 by (distance, index); device pointers */
int  orbx_knn2_merge_device(orbx_matcher* m, const int32_t* d_idx_parts, const int32_t* d_dist_parts, int nparts,
                            int nq, int32_t* d_idx, int32_t* d_dist, void* stream);

/* ORBmatcher::SearchForInitialization, R/src/ORBmatcher.cc:702-817 (with Frame::AssignFeaturesToGrid /
 * GetFeaturesInArea, R/src/Frame.cc:360-391, 628-697).  bounds = {mnMinX, mnMaxX, mnMinY, mnMaxY}.
 * prev_xy (n1 x 2, vbPrevMatched) is updated in place; matches12 has n1 entries.  Host pointers. */
int  orbx_search_for_initialization(orbx_matcher* m, const orbx_keypoint* k1, const uint8_t* d1, int n1,
                                    const orbx_keypoint* k2, const uint8_t* d2, int n2, const float bounds[4],
                                    float* prev_xy, int32_t* matches12, int window, float nnratio, int check_ori,
                                    int* nmatches);
/* batched over extractor result slots on the device: pair i matches slot a[i] (F1) against slot b[i] (F2)
 * with vbPrevMatched = F1's keypoint positions (first call of Tracking.cc:2216-2217).  Also runs
 * orbx_bf_knn2 for every pair when d_knn_idx != NULL.  a, b are DEVICE pointers.  Outputs on device, rows have
 * stride K = the matcher's max_keypoints: matches12 [npairs][K], nmatches [npairs], knn idx/dist [npairs][K][2]. */
int  orbx_match_slots_device(orbx_matcher* m, orbx_extractor* ex, const int32_t* a, const int32_t* b, int npairs,
                             const float bounds[4], int window, float nnratio, int check_ori,
                             int32_t* d_matches12, int32_t* d_nmatches, int32_t* d_knn_idx, int32_t* d_knn_dist,
                             void* stream);
/* The reference matches on mvKeysUn (R/src/Frame.cc:721-754 feeds R/src/ORBmatcher.cc:702-817).  For a distorted camera
 * give the slot-based searches (orbx_match_slots_device, orbx_extract_match_batch*) the undistorted keypoints: a DEVICE
 * array laid out like the extractor's results, [slots][orbx_extractor_max_keypoints(ex)], typically written by
 * orbx_undistort_slots_device.  NULL (default) = the extractor's own keypoints (mvKeysUn == mvKeys when mDistCoef[0] == 0). */
int  orbx_matcher_set_slot_keypoints(orbx_matcher* m, const orbx_keypoint* d_kps_un);
/* The stream pipelines do that by themselves once the camera is known: K, P 3x3 row-major, dist = ndist (4..12) distortion
 * coefficients as in Frame::UndistortKeyPoints (R/src/Frame.cc:721-754); every chunk is undistorted right after extraction
 * and matched on mvKeysUn.  K = NULL clears the camera.  orbx_matcher_undistorted_device returns the device view of
 * mvKeysUn of the last pipeline call ([slots][orbx_extractor_max_keypoints], slot i + 1 = frame i; NULL without camera). */
int  orbx_matcher_set_camera(orbx_matcher* m, const float* K, const float* dist, int ndist, const float* P);
int  orbx_matcher_undistorted_device(orbx_matcher* m, orbx_keypoint** d_kps_un);


/* One tracking step over a batch of HOST frames (the end-to-end path): H2D, operator() on every frame into result
 * slots 1..batch, SearchForInitialization of each frame against its predecessor (slot i-1 -> slot i; slot 0 keeps
 * the last frame of the previous call, as Tracking keeps mLastFrame), D2H of keypoints, descriptors and matches.
 * Pinned caller buffers are DMA'd directly.  kps: batch*cap, desc: batch*cap*32, matches12: batch*cap.
 * knn_idx / knn_dist (both or neither; batch*cap*2 each): when given, the brute-force kNN-2 of every frame's predecessor
 * against the frame (cv::BFMatcher::knnMatch(k = 2), R/src/Frame.cc:1127-1137; rows = the predecessor's keypoints) is run
 * and copied back as well.  batch must fit both handles (matcher max_batch, extractor max_batch); both handles on one device. */
int  orbx_extract_match_batch(orbx_extractor* ex, orbx_matcher* m, const uint8_t* imgs, int batch, int width,
                              int height, int stride, size_t frame_stride, int lap0, int lap1,
                              const float bounds[4], int window, float nnratio, int check_ori,
                              orbx_keypoint* kps, uint8_t* desc, int cap, int32_t* n, int32_t* mono_index,
                              int32_t* matches12, int32_t* nmatches, int32_t* knn_idx, int32_t* knn_dist);

/* Input prefetch for a stream of host batches: starts the host-to-device copy of the frames of the NEXT orbx_extract_match_batch
 * call and returns at once, so that the PCIe transfer of batch k+1 runs under the kernels of batch k.  The frames must be pinned
 * and packed at the staging pitch (stride = width rounded up to 16, frame_stride = stride * height) and must stay unchanged until
 * the orbx_extract_match_batch call with the same (imgs, batch, width, height) consumes them; at most two batches can wait.  A
 * call with other arguments drops what waits and copies its own input as usual.  Results are those of the plain call. */
int  orbx_extract_match_batch_prefetch(orbx_extractor* ex, orbx_matcher* m, const uint8_t* imgs, int batch, int width,
                                       int height, int stride, size_t frame_stride);

/* Streaming form of orbx_extract_match_batch for a camera stream: submit queues one batch (input copy on the copy stream, the
 * kernels of the step, result copy on the D2H stream) and returns at once with a ticket; wait blocks until THAT batch's results
 * are in the buffers given at submit (and reports its device error flags).  Two batches may be in flight: with
 *     submit(k + 1); wait(k);
 * the input copy of batch k+1, the kernels of batch k and the result copy of batch k-1 ... run at the same time, and every frame
 * still finds its predecessor (batches are processed in submission order on one kernel stream).  The frames must be pinned and
 * packed at the staging pitch (as for the prefetch call), the result buffers pinned with cap = orbx_extractor_max_keypoints(ex)
 * and distinct for the two batches in flight; all must stay untouched until the wait returns.  Results are those of the plain
 * call.  ORBX_E_CAPACITY from submit: two batches are already in flight. */
int  orbx_stream_submit(orbx_extractor* ex, orbx_matcher* m, const uint8_t* imgs, int batch, int width, int height, int stride,
                        size_t frame_stride, int lap0, int lap1, const float bounds[4], int window, float nnratio, int check_ori,
                        orbx_keypoint* kps, uint8_t* desc, int cap, int32_t* n, int32_t* mono_index,
                        int32_t* matches12, int32_t* nmatches, int32_t* knn_idx, int32_t* knn_dist, long long* ticket);
int  orbx_stream_wait(orbx_extractor* ex, orbx_matcher* m, long long ticket);

/* The same step on DEVICE-resident frames, asynchronous on `stream` (results stay in the slots; matches12 [batch][K],
 * nmatches [batch] and the optional BF kNN-2 tables [batch][K][2] are device arrays, K = the matcher's max_keypoints).
 * Internally the batch is cut into chunks whose matcher kernels run on a second stream under the next chunk's
 * extraction; `stream` waits for all of it before the slot carry, so synchronising `stream` completes the step. */
int  orbx_extract_match_batch_device(orbx_extractor* ex, orbx_matcher* m, const uint8_t* d_imgs, int batch, int width,
                                     int height, int stride, size_t frame_stride, int lap0, int lap1,
                                     const float bounds[4], int window, float nnratio, int check_ori,
                                     int32_t* d_matches12, int32_t* d_nmatches, int32_t* d_knn_idx, int32_t* d_knn_dist,
                                     void* stream);

/* Projection-guided window searches on flat arrays (the drop-in ORBmatcher marshals Frame/MapPoint into
 * these).  Query i: window centre (u,v), radius r, octave range [minl,maxl] as GetFeaturesInArea takes them,
 * predicted right coordinate ur (stereo gate), angle (rotation histogram), valid flags: bit 0 = the query takes part; bit 1
 * (value 2) = the query's MapPoint has Observations() == 0, so a keypoint it claims stays free for later queries (the
 * reference only skips keypoints whose MapPoint has observations, R/src/ORBmatcher.cc:89-91, :2045-2047; the temporal points
 * of Tracking::UpdateLastFrame have none), every accepting query counts as a match.
 *   mode 0: SearchByProjection(Frame&, const Frame&, th, bMono)      R/src/ORBmatcher.cc:1970-2186
 *   mode 1: SearchByProjection(Frame&, vector<MapPoint*>&, th, ...)   R/src/ORBmatcher.cc:44-214
 * assigned[n2]: in = -1 for free keypoints (anything >= 0 is skipped like an occupied mvpMapPoints slot),
 * out = index of the query that took the keypoint; -2 = a keypoint that was claimed during the call and then cleared by the
 * rotation check (the reference sets such a slot to NULL even when it held a 0-observation MapPoint before, :2176-2180).
 * Host pointers. */
typedef struct orbx_proj_query {
    float u, v, r;
    int32_t minl, maxl;
    float ur;
    float angle;
    int32_t valid;
} orbx_proj_query;
int  orbx_search_by_projection(orbx_matcher* m, int mode, const orbx_proj_query* q, const uint8_t* qdesc, int nq,
                               const orbx_keypoint* k2, const uint8_t* d2, const float* uright2, int n2,
                               const float bounds[4], int32_t* assigned, float nnratio, int check_ori, int* nmatches);
/* The same machinery with the knobs of the reference's other projection searches:
 *   mode 0 + max_dist: acceptance bound of the best distance: ORBX_TH_HIGH for :1970-2186; floor(TH_LOW * ratioHamming) for
 *       the Sim3 overloads SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th, ratioHamming) (:473-586, :588-700), where
 *       assigned[] is preset from vpMatched and check_ori = 0; ORBdist for the relocalisation overload
 *       SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist) (:2188-2310);
 *   inv_level_sigma2[nlevels] + chi2: per-candidate gate of Fuse (:1497-1505), a candidate is skipped when
 *       |q - kp|^2 * invSigma2[octave] (float) > chi2 (5.99 mono, 7.8 stereo); chi2 = 0 disables it;
 *   mode 3: the best candidate of every query on its own, no bookkeeping between queries (Fuse :1395-1742, where the map
 *       update that follows stays on the host): best_idx / best_dist [nq] (-1 / 256 without a candidate), assigned unused;
 *       *nmatches = queries with a candidate. */
int  orbx_search_by_projection_ex(orbx_matcher* m, int mode, const orbx_proj_query* q, const uint8_t* qdesc, int nq,
                                  const orbx_keypoint* k2, const uint8_t* d2, const float* uright2, int n2,
                                  const float bounds[4], int32_t* assigned, float nnratio, int check_ori, int max_dist,
                                  const float* inv_level_sigma2, int nlevels, double chi2,
                                  int32_t* best_idx, int32_t* best_dist, int* nmatches);
/* The full set of knobs in one record (the entry the drop-in ORBmatcher calls):
 *   bounds        the grid of the searched frame: Frame::mnMinX, mnMaxX, mnMinY, mnMaxY (R/src/Frame.cc:314-327);
 *   query_origin  origin of the query cell range.  Frame::GetFeaturesInArea (R/src/Frame.cc:639-657) uses the same float
 *                 mnMinX / mnMinY; KeyFrame::GetFeaturesInArea (R/src/KeyFrame.cc:897-911) uses KeyFrame::mnMinX / mnMinY, which
 *                 are `const int` copies (R/include/KeyFrame.h:501-504): pass (float)(int)mnMinX for a keyframe whose grid
 *                 was built on a distorted camera's fractional bounds;
 *   chi2_mono / chi2_stereo / inv_level_sigma2[nlevels]  the reprojection gates of Fuse (R/src/ORBmatcher.cc:1525-1552): a
 *                 candidate with mvuRight >= 0 is skipped when (ex^2 + ey^2 + er^2) * invSigma2[octave] > chi2_stereo (7.8, er
 *                 from the query's `ur`), any other candidate when (ex^2 + ey^2) * invSigma2[octave] > chi2_mono (5.99);
 *                 chi2_mono = 0 disables both, chi2_stereo = 0 applies the mono form to every candidate. */
typedef struct orbx_proj_options {
    float bounds[4];
    float query_origin[2];
    float nnratio;
    int32_t check_ori;
    int32_t max_dist;
    int32_t nlevels;
    float inv_level_sigma2[ORBX_MAX_LEVELS];
    double chi2_mono, chi2_stereo;
} orbx_proj_options;
int  orbx_search_by_projection_opts(orbx_matcher* m, int mode, const orbx_proj_query* q, const uint8_t* qdesc, int nq,
                                    const orbx_keypoint* k2, const uint8_t* d2, const float* uright2, int n2,
                                    const orbx_proj_options* opt, int32_t* assigned,
                                    int32_t* best_idx, int32_t* best_dist, int* nmatches);


/* Generic candidate matching for the searches whose candidate gathering stays on the host: SearchByBoW
 * (R/src/ORBmatcher.cc:269-471, 819-959), SearchForTriangulation (:961-1394), Fuse (:1395-1742), SearchBySim3 (:1744-1968).
 * Query i is compared with train rows indices[offsets[i] .. offsets[i+1]) in list order; idx/dist (nq x 2) receive the
 * best and second best by (distance, list position), -1 when missing.  Host pointers. */
int  orbx_match_candidates(orbx_matcher* m, const uint8_t* q, int nq, const uint8_t* t, int nt, const int32_t* offsets,
                           const int32_t* indices, int32_t* idx, int32_t* dist);

/* Frame::ComputeStereoMatches, descriptor search (R/src/Frame.cc:785-868): per left keypoint the best
 * right index (-1 if none) and its distance (starts at TH_HIGH).  Host pointers. */
int  orbx_stereo_band_match(orbx_matcher* m, const orbx_keypoint* kl, const uint8_t* dl, int nl,
                            const orbx_keypoint* kr, const uint8_t* dr, int nr, const float* scale_factors,
                            int nlevels, int nrows, float min_d, float max_d, int32_t* best_idx, int32_t* best_dist);

/* Frame::ComputeStereoMatches in full (R/src/Frame.cc:785-962) on the device-resident results and pyramids of the left
 * and right extractor (R/src/Frame.cc:92-95 runs one ORBextractor per camera): descriptor search, 11x11 SAD refinement
 * over +-5 px on the pyramid level of the left keypoint, parabola sub-pixel fit, median-based outlier cut.
 * slot_* = result slot, frame_* = frame index inside each extractor's last batch (0 for orbx_extract).
 * uright/depth receive mvuRight/mvDepth (-1 = no match), sad_dist (optional) the SAD distance or -1. */
int  orbx_stereo_matches(orbx_matcher* m, orbx_extractor* left, orbx_extractor* right, int slot_l, int slot_r,
                         int frame_l, int frame_r, float mb, float mbf, float* uright, float* depth,
                         int32_t* sad_dist, int cap, int* n_left);
/* The same for a batch of stereo pairs (a KITTI / EuRoC stereo stream): pair p uses result slot and frame (first + p) of
 * both extractors' last batch.  _device: outputs are DEVICE arrays [count][orbx_extractor_max_keypoints(left)], d_sad may
 * be NULL, all work is enqueued on `stream` (the stream of the two orbx_extract_batch_device calls) and nothing is
 * synchronised.  Host form: uright/depth are host arrays [count][cap]; synchronises. */
int  orbx_stereo_matches_batch_device(orbx_matcher* m, orbx_extractor* left, orbx_extractor* right, int first, int count,
                                      float mb, float mbf, float* d_uright, float* d_depth, int32_t* d_sad, void* stream);
int  orbx_stereo_matches_batch(orbx_matcher* m, orbx_extractor* left, orbx_extractor* right, int first, int count,
                               float mb, float mbf, float* uright, float* depth, int cap);
/* A batch of stereo frames from HOST memory to HOST results in one call: ORBextractor::operator() on every left and right
 * frame (two extractor instances, as R/src/Frame.cc:92-95) and Frame::ComputeStereoMatches for every pair; chunks of the
 * batch flow through separate copy / kernel streams.  kps_* [batch][cap], desc_* [batch][cap][32], n_* [batch],
 * uright / depth [batch][cap]; any of kps_*, desc_* may be NULL.  batch <= max_batch of both extractors. */
int  orbx_extract_stereo_batch(orbx_matcher* m, orbx_extractor* left, orbx_extractor* right,
                               const uint8_t* imgs_left, const uint8_t* imgs_right, int batch, int width, int height,
                               int stride, size_t frame_stride, float mb, float mbf,
                               orbx_keypoint* kps_l, uint8_t* desc_l, int32_t* n_l,
                               orbx_keypoint* kps_r, uint8_t* desc_r, int32_t* n_r, int cap,
                               float* uright, float* depth);



/* register-only popcount micro-benchmark: returns measured 32-bit popc per second on `device` (roofline
 * denominator for the matching kernels) */
int  orbx_popc_peak(int device, double* popc_per_s, double* lop3_per_s);

/* Frame::UndistortKeyPoints (R/src/Frame.cc:721-754): mvKeysUn = cv::undistortPoints(mvKeys, K, mDistCoef,
 * R = I, P = mK) exactly as OpenCV 4.x computes it (double arithmetic, five fixed-point iterations, float result); K and P
 * are 3x3 row-major float, dist has ndist = 4..12 coefficients (k1, k2, p1, p2[, k3 ...]); dist[0] == 0 copies the
 * keypoints (:723-727).  Host form: synchronous.  _slots_device: on the result slots of an extractor, d_kps_un is a
 * DEVICE array [count][orbx_extractor_max_keypoints(ex)], asynchronous on `stream`. */
int  orbx_undistort_keypoints(orbx_matcher* m, const orbx_keypoint* kps, int n, const float* K, const float* dist, int ndist,
                              const float* P, orbx_keypoint* kps_un);
int  orbx_undistort_slots_device(orbx_extractor* ex, int first_slot, int count, const float* K, const float* dist, int ndist,
                                 const float* P, orbx_keypoint* d_kps_un, void* stream);

/* KF.msg wire format of the keypoints: msg/CvKeyPoint.msg as ROS1 serialises it, 15 packed bytes per
 * keypoint (float32 x, float32 y, uint8 size, float32 angle, uint8 response, int8 octave); Converter::toCvKeyPointMsg /
 * fromCvKeyPointMsg (R/src/Converter.cc:218-244) called from KeyFrame.cc:1430 and :1929.  size / response are truncated
 * to 8 bits, class_id is not transmitted (-1 after unpacking).  Descriptor.msg is the 32-byte descriptor row itself.
 * Host forms: synchronous.  _slot_..._device: one result slot of an extractor into a DEVICE buffer of
 * orbx_extractor_max_keypoints(ex) * 15 bytes, asynchronous on `stream`. */
int  orbx_keypoints_to_msg(orbx_matcher* m, const orbx_keypoint* kps, int n, uint8_t* msg15);
int  orbx_keypoints_from_msg(orbx_matcher* m, const uint8_t* msg15, int n, orbx_keypoint* kps);
int  orbx_slot_keypoints_to_msg_device(orbx_extractor* ex, int slot, uint8_t* d_msg15, void* stream);

/* MapPoint::ComputeDistinctiveDescriptors (R/src/MapPoint.cc:448-524) for a batch of map points, as
 * LocalMapping runs it for every point a new keyframe observes: the observed descriptors of point p are rows
 * offsets[p] .. offsets[p+1] of desc ([total][32], gathered by the caller from the observing keyframes); best[p] receives
 * the index inside that run of the descriptor with the least median Hamming distance to the others (median = sorted row
 * [(int)(0.5 * (N - 1))] with the 0 of the diagonal included; first minimum wins), -1 for a point without observations.
 * Host pointers, synchronous. */
int  orbx_distinctive_descriptors(orbx_matcher* m, const uint8_t* desc, const int32_t* offsets, int npoints, int32_t* best);

/* ORBmatcher::SearchByBoW on flat arrays.  The two FeatureVectors are CSR tables sorted by node id (the order of the
 * std::map the reference iterates): fv_nodes[nfv], fv_start[nfv + 1], fv_feat[fv_start[nfv]].  valid = "the feature has a
 * MapPoint that is not bad".  matches12[i1] receives the matched index in set 2, or -1; *nmatches the return value.
 * mode 0 = SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) (R/src/ORBmatcher.cc:269-471, Nleft == -1 branch): set 1 =
 *          keyframe, set 2 = frame, valid2 ignored, accept bestDist1 <= TH_LOW; the caller stores
 *          vpMapPointMatches[matches12[i1]] = MapPoint of i1;
 * mode 1 = SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&) (R/src/ORBmatcher.cc:819-959): candidates need valid2,
 *          accept bestDist1 < TH_LOW; vpMatches12[i1] = MapPoint of matches12[i1].
 * Host pointers, synchronous. */
int  orbx_search_by_bow(orbx_matcher* m, int mode,
                        const orbx_keypoint* k1, const uint8_t* d1, const uint8_t* valid1, int n1,
                        const int32_t* fv1_nodes, const int32_t* fv1_start, const int32_t* fv1_feat, int nfv1,
                        const orbx_keypoint* k2, const uint8_t* d2, const uint8_t* valid2, int n2,
                        const int32_t* fv2_nodes, const int32_t* fv2_start, const int32_t* fv2_feat, int nfv2,
                        float nnratio, int check_ori, int32_t* matches12, int* nmatches);

/* ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo, bCoarse) (R/src/ORBmatcher.cc:961-1202) for
 * pinhole cameras without a second camera, as LocalMapping::CreateNewMapPoints calls it for every neighbour keyframe.
 * free* = the feature has no MapPoint; stereo* (may be NULL) = mvuRight >= 0; FeatureVectors as CSR tables sorted by node id.
 * A candidate needs distance <= TH_LOW, (mono-mono) at least 100 * scale_factors2[octave2] squared pixels from the epipole
 * (ep_x, ep_y) and, unless coarse, to lie within 3.84 * level_sigma2_2[octave2] of the epipolar line of F12 (3x3 row
 * major, Pinhole::epipolarConstrain, CameraModels/Pinhole.cpp:121-143); smallest distance wins, among equals the last in
 * list order.  The reference never marks keyframe-2 features as taken in this function, so queries do not interact.
 * matches12[i1] = index in keyframe 2 or -1.  Host pointers, synchronous. */
int  orbx_search_for_triangulation(orbx_matcher* m,
                                   const orbx_keypoint* k1, const uint8_t* d1, const uint8_t* free1, const uint8_t* stereo1, int n1,
                                   const int32_t* fv1_nodes, const int32_t* fv1_start, const int32_t* fv1_feat, int nfv1,
                                   const orbx_keypoint* k2, const uint8_t* d2, const uint8_t* free2, const uint8_t* stereo2, int n2,
                                   const int32_t* fv2_nodes, const int32_t* fv2_start, const int32_t* fv2_feat, int nfv2,
                                   const float* F12, float ep_x, float ep_y, const float* scale_factors2,
                                   const float* level_sigma2_2, int nlevels, int only_stereo, int coarse, int check_ori,
                                   int32_t* matches12, int* nmatches);

/* ---- bag of words: DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB> as Frame::ComputeBoW
 * (R/src/Frame.cc:712-719) and KeyFrame::ComputeBoW (R/src/KeyFrame.cc:168-176) use it ---- */
typedef struct orbx_vocab orbx_vocab;
/* Vocabulary from its node table = the rows of ORBvoc.txt after the header, in file order
 * (R/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h, loadFromTextFile): node 0 is the root, parent[i] < i, children keep
 * their order of appearance, is_leaf marks exactly the childless nodes = the words, numbered in node order; desc is
 * [n_nodes][32], weight [n_nodes] (idf of the words); L = depth (6 for ORBvoc). */
int  orbx_vocab_create(int device, int n_nodes, const int32_t* parent, const uint8_t* is_leaf, const uint8_t* desc,
                       const double* weight, int L, orbx_vocab** out);
void orbx_vocab_destroy(orbx_vocab* v);
int  orbx_vocab_words(const orbx_vocab* v);
int  orbx_vocab_word_weights(const orbx_vocab* v, double* w, int cap);
/* Per-feature part of transform(features, BowVector&, FeatureVector&, levelsup) (TemplatedVocabulary.h:1127-1200,
 * :1218-1259; distance FORB.cpp:81-101): word id, word weight (may be NULL) and the node at level L - levelsup of every
 * descriptor (0 = root when that level is <= 0; the last node reached when a leaf lies above it, where the reference
 * leaves the value uninitialised).  The first child with the smallest distance wins, as the reference's strict '<'.
 * Host pointers, synchronous.  The BowVector / FeatureVector maps are assembled from these arrays by the caller
 * (dropin/ORBVocabulary, multi_orbslam3_b200/orbx.py): sums of equal weights and an L1 norm over ~1000 words. */
int  orbx_bow_transform(orbx_vocab* v, const uint8_t* desc, int n, int levelsup, int32_t* word_id, double* weight,
                        int32_t* node_id);
/* The same on the descriptors of an extractor's result slots (never leaving the GPU): d_word / d_node are DEVICE arrays
 * [count][orbx_extractor_max_keypoints(ex)]; asynchronous on `stream`. */
int  orbx_bow_transform_slots_device(orbx_vocab* v, orbx_extractor* ex, int first_slot, int count, int levelsup,
                                     int32_t* d_word, int32_t* d_node, void* stream);

/* Modes 0 and 1 on a two-camera frame (KannalaBrandt8 rig, Frame::Nleft != -1; R/src/ORBmatcher.cc:144-213, :2093-2160).
 * k2 / d2 hold the left camera's keypoints (indices [0, n_left): Frame::mvKeys) followed by the right camera's ([n_left, n_left +
 * n_right): Frame::mvKeysRight, descriptor rows Nleft ..), each half with its own grid (mGrid / mGridRight, R/src/Frame.cc:360-391,
 * GetFeaturesInArea(..., bRight) :628-697) over the same bounds.  Map point i has a left query ql[i] and a right query qr[i] (valid
 * bit 0 = that search takes part, bit 1 = the point has no observations) and one descriptor qdesc[i]; the reference interleaves
 * them - left of point i, right of point i, left of point i+1, ... - over ONE occupancy table, and so does this call:
 *   mode 1  SearchByProjection(Frame&, vector<MapPoint*>&): best / second with the level rule on both sides; an accepted left match
 *           also takes the keypoint's stereo partner l2r[i2] (Frame::mvLeftToRightMatch, -1 = none) and counts twice, likewise r2l
 *           (mvRightToLeftMatch) for a right match; a left search rejected by the ratio test leaves the point (the `continue` of
 *           :116-117 skips its right search);
 *   mode 0  SearchByProjection(Frame&, const Frame&, th, bMono): left as in the one-camera form without the stereo gate, right: the
 *           best candidate alone, accepted when <= opt->max_dist; both feed the one rotation histogram; a left search that takes
 *           part but finds no keypoint in its window leaves the point (the `continue` of :2033-2034 skips its right search).
 * assigned[n_left + n_right] in / out as for orbx_search_by_projection (-2 = cleared by the rotation check).  Needs a matcher with
 * max_batch >= 2 and n_left + n_right <= max_keypoints.  Host pointers. */
int  orbx_search_by_projection_rig(orbx_matcher* m, int mode, const orbx_proj_query* ql, const orbx_proj_query* qr,
                                   const uint8_t* qdesc, int nq, const orbx_keypoint* k2, const uint8_t* d2, int n_left, int n_right,
                                   const int32_t* l2r, const int32_t* r2l, const orbx_proj_options* opt, int32_t* assigned,
                                   int* nmatches);

/* SearchByBoW(KeyFrame*, Frame&) when the frame has two cameras (Frame::Nleft != -1; R/src/ORBmatcher.cc:344-431): features
 * [0, n2_left) of set 2 are the left camera's (Frame::mvKeys), the rest the right camera's (mvKeysRight).  Per keyframe feature the
 * reference keeps a best / second pair per camera: the left match needs best <= TH_LOW and the ratio test; the right match needs
 * bestLeft <= TH_LOW (it sits inside that branch) and bestRight <= TH_LOW (its ratio test is disabled by `|| true`, :402).  Both
 * claim their feature and share the rotation histogram.  matches12_left / matches12_right [n1]; *nmatches counts both. */
int  orbx_search_by_bow_rig(orbx_matcher* m,
                            const orbx_keypoint* k1, const uint8_t* d1, const uint8_t* valid1, int n1,
                            const int32_t* fv1_nodes, const int32_t* fv1_start, const int32_t* fv1_feat, int nfv1,
                            const orbx_keypoint* k2, const uint8_t* d2, int n2, int n2_left,
                            const int32_t* fv2_nodes, const int32_t* fv2_start, const int32_t* fv2_feat, int nfv2,
                            float nnratio, int check_ori, int32_t* matches12_left, int32_t* matches12_right, int* nmatches);

/* ---- server keyframe database on one GPU (SURVEY 8f row 3 -> 8e) ----
 * A keyframe reaches the server as a KF.msg (R/msg/KF.msg:24-29): `CvKeyPoint[] mvKeysUn` = N records of 15 packed bytes
 * (R/msg/CvKeyPoint.msg; Converter::toCvKeyPointMsg, R/src/Converter.cc:218-230) and `Descriptor[] mDescriptors` = N x
 * uint8[32] (R/msg/Descriptor.msg:1: fixed-size arrays carry no length prefix, the N descriptors are 32 N contiguous bytes
 * of the serialised message).  Replaces the element-by-element rebuild of KeyFrame.cc:1929-1944 / Converter.cc:232-244:
 * both byte runs go to the device as they are, the descriptors straight into this GPU's shard of the descriptor DB, the
 * keypoint records through the device unpack kernel.  The shard is searched by orbx_kfdb_knn2 / orbx_bf_knn2_device
 * (multi_orbslam3_b200.server.ShardedDescriptorDB for the multi-GPU exchange).  Ingestion is asynchronous on the DB's own
 * stream and thread-safe (one communication thread per client in the reference, Communicator.cc:110-148). */
typedef struct orbx_kfdb orbx_kfdb;
int  orbx_kfdb_create(int device, long long capacity_rows, int max_keyframes, orbx_kfdb** out);
void orbx_kfdb_destroy(orbx_kfdb* db);
/* appends one keyframe (host pointers into the received message; msg_keys15 may be NULL = descriptors only).
 * *first_row = DB row of its first descriptor.  ORBX_E_CAPACITY when rows or keyframes are exhausted: nothing is appended. */
int  orbx_kfdb_ingest_msg(orbx_kfdb* db, int64_t kf_id, const uint8_t* msg_keys15, const uint8_t* msg_desc32, int n,
                          long long* first_row);
/* the same from a result slot of an extractor on the same device (an agent whose frames are extracted on this GPU: no host hop) */
int  orbx_kfdb_ingest_slot_device(orbx_kfdb* db, int64_t kf_id, orbx_extractor* ex, int slot, long long* first_row);
int  orbx_kfdb_size(orbx_kfdb* db, long long* rows, int* keyframes, long long* capacity_rows);
/* device views [capacity_rows][32] / [capacity_rows]; rows ingested before the last orbx_kfdb_sync are complete */
int  orbx_kfdb_device(orbx_kfdb* db, const uint8_t** d_desc, const orbx_keypoint** d_kps);
int  orbx_kfdb_sync(orbx_kfdb* db);
/* DB row -> (keyframe id, feature index inside that keyframe); -1 / -1 for rows outside the DB.  Host arrays. */
int  orbx_kfdb_locate(orbx_kfdb* db, const long long* rows, int n, int64_t* kf_id, int32_t* feature);
/* cv::BFMatcher kNN-2 (Frame.cc:1127-1137 semantics) of host queries against everything ingested so far;
 * idx = idx_base + DB row (idx_base = first global row of this GPU's shard). */
int  orbx_kfdb_knn2(orbx_kfdb* db, orbx_matcher* m, const uint8_t* q, int nq, long long idx_base, int32_t* idx, int32_t* dist);

#ifdef __cplusplus
}
#endif
#endif
