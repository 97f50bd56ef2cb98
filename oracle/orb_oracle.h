/* orb_oracle.h - CPU oracle (TEST INFRASTRUCTURE, not product code).
 *
 * Plain-C restatement of the reference's ORB front-end hot path.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product path (multi_orbslam3_b200/) never does.
 *
 * R/ = /root/reference/src/orb_slam3_ros/orb_slam3/
 * Reference logic restated:   R/src/ORBextractor.cc:70-145, 408-878, 1059-1177
 *                             R/src/ORBmatcher.cc:36-222, 702-817, 1970-2186, 2312-2374
 *                             R/src/Frame.cc:360-391, 628-709, 785-868, 1127-1137
 * Third-party arithmetic the reference calls but does not contain (OpenCV, unpinned by
 * R/../CMakeLists.txt:55-65) is restated from OpenCV 4.x's published algorithms and
 * PINNED against cv2 4.13.0 by tests/test_oracle_pin_cv2.py and tests/golden/.
 */
#ifndef ORB_ORACLE_H
#define ORB_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* binary-compatible with cv::KeyPoint (28 bytes) */
typedef struct {
    float x, y;
    float size;
    float angle;
    float response;
    int32_t octave;
    int32_t class_id;
} OrcKeyPoint;

/* ---- OpenCV primitives (restated) ---- */
void  orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride,
                           uint8_t* dst, int dw, int dh, int dstride);
void  orc_gaussian_blur7(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride);
/* cv::FAST(img, t, nonmaxSuppression, TYPE_9_16): writes (x,y,score) int triples, returns count */
int   orc_fast9_16(const uint8_t* img, int w, int h, int stride, int threshold, int nms,
                   int32_t* out_xys, int cap);
float orc_fast_atan2(float y, float x);
int   orc_cv_round_f(float v);

/* ---- extractor ---- */
typedef struct OrcExtractor OrcExtractor;
OrcExtractor* orc_extractor_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th);
void  orc_extractor_destroy(OrcExtractor* e);
/* tables: out arrays have nlevels entries */
void  orc_extractor_tables(const OrcExtractor* e, float* scale, float* inv_scale, float* sigma2,
                           float* inv_sigma2, int32_t* features_per_level, int32_t* umax16);
/* ORBextractor::operator(): returns monoIndex, or -1 on empty image; *n_out = #keypoints */
int   orc_extract(OrcExtractor* e, const uint8_t* img, int w, int h, int stride, int lap0, int lap1,
                  OrcKeyPoint* kps, uint8_t* desc, int cap, int* n_out);
/* stage taps (valid until the next orc_extract on this handle) */
int   orc_level_size(const OrcExtractor* e, int level, int* w, int* h);
const uint8_t* orc_level_image(const OrcExtractor* e, int level);   /* tight stride = w */
const uint8_t* orc_level_blurred(const OrcExtractor* e, int level); /* NULL when level had no keypoints */
/* FAST candidates handed to the octree: float triples (x,y,response), coords relative to (16,16) */
int   orc_level_candidates(const OrcExtractor* e, int level, const float** xyr);
/* octree output with border added, octave/size/angle set, BEFORE scaling (list order) */
int   orc_level_keypoints(const OrcExtractor* e, int level, const OrcKeyPoint** kps);

/* standalone DistributeOctTree (canonical tie rule: equal size -> later-created node first) */
int   orc_distribute_octree(const float* xyr, int n, int minX, int maxX, int minY, int maxY, int N,
                            float* out_xyr, int cap);

/* ---- matching ---- */
int   orc_hamming256(const uint8_t* a, const uint8_t* b);
/* cv::BFMatcher(NORM_HAMMING).knnMatch(k=2): idx/dist are nq*2; missing -> idx -1 */
void  orc_bf_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx, int32_t* dist);

typedef struct OrcGrid OrcGrid;   /* Frame::mGrid, 64 x 48 */
OrcGrid* orc_grid_build(const OrcKeyPoint* kps, int n, float minX, float maxX, float minY, float maxY);
void  orc_grid_destroy(OrcGrid* g);
int   orc_features_in_area(const OrcGrid* g, const OrcKeyPoint* kps, float x, float y, float r,
                           int minLevel, int maxLevel, int32_t* out, int cap);

/* ORBmatcher::SearchForInitialization; prev_xy (n1*2) is updated in place; returns nmatches */
int   orc_search_for_initialization(const OrcKeyPoint* k1, const uint8_t* d1, int n1,
                                    const OrcKeyPoint* k2, const uint8_t* d2, int n2,
                                    float minX, float maxX, float minY, float maxY,
                                    float* prev_xy, int32_t* matches12, int window,
                                    float nnratio, int check_ori);

/* Projection-guided searches on flat arrays (the drop-in ORBmatcher marshals Frame/MapPoint into these).
 * query i: window centre (u,v), radius r, level range [minl,maxl], 32-byte descriptor, angle.
 * frame: keypoints/descriptors/grid bounds, uright[n2] (<=0: mono), assigned[n2] in/out (-1 free,
 * else the query index that owns the keypoint).
 * mode 0 = SearchByProjection(Frame&,const Frame&,th,bMono)   (ORBmatcher.cc:1970-2186): best only, <=TH_HIGH
 * mode 1 = SearchByProjection(Frame&,vector<MapPoint*>&,th..) (ORBmatcher.cc:44-214): best+second, level ratio rule */
typedef struct {
    float u, v, r;
    int32_t minl, maxl;
    float ur;          /* predicted right coordinate for the stereo gate, used when frame uright>0 */
    float angle;       /* for the rotation histogram (mode 0) */
    int32_t valid;     /* bit 0: the query takes part; bit 1 (value 2): its MapPoint has Observations() == 0, so a keypoint it
                        * claims stays free for later queries (ORBmatcher.cc:89-91, :2045-2047) */
} OrcProjQuery;
int   orc_search_by_projection(int mode, const OrcProjQuery* q, const uint8_t* qdesc, int nq,
                               const OrcKeyPoint* k2, const uint8_t* d2, const float* uright2, int n2,
                               float minX, float maxX, float minY, float maxY,
                               int32_t* assigned, float nnratio, int check_ori);

/* The same with the knobs the other projection searches of the reference use:
 *   max_dist  acceptance threshold of mode 0 (TH_HIGH for :1970-2186; TH_LOW * ratioHamming for the Sim3 overloads :473-700,
 *             ORBdist for the relocalisation overload :2188-2310);
 *   inv_sigma2 / chi2  per-candidate reprojection gate of Fuse (:1497-1505): skip when |q - kp|^2 * inv_sigma2[octave] > chi2
 *             (the product in float, the comparison against the double literal 5.99); NULL / 0 = no gate;
 *   mode 3    independent best per query, no bookkeeping (Fuse :1395-1742): best_idx / best_dist [nq] (-1 / 256 if none);
 *             returns the number of queries with a candidate. */
int   orc_search_by_projection_ex(int mode, const OrcProjQuery* q, const uint8_t* qdesc, int nq,
                                  const OrcKeyPoint* k2, const uint8_t* d2, const float* uright2, int n2,
                                  float minX, float maxX, float minY, float maxY,
                                  int32_t* assigned, float nnratio, int check_ori, int max_dist,
                                  const float* inv_sigma2, double chi2, int32_t* best_idx, int32_t* best_dist);

/* The same with the two remaining knobs: (qminX, qminY) = origin of the query cell range (KeyFrame::GetFeaturesInArea uses the
 * int-truncated KeyFrame::mnMinX / mnMinY, R/src/KeyFrame.cc:897-911), and chi2_stereo = Fuse's 3-dof gate for keypoints with
 * mvuRight >= 0 (R/src/ORBmatcher.cc:1525-1540; 0 = mono form for every candidate). */
int   orc_search_by_projection_full(int mode, const OrcProjQuery* q, const uint8_t* qdesc, int nq,
                                    const OrcKeyPoint* k2, const uint8_t* d2, const float* uright2, int n2,
                                    float minX, float maxX, float minY, float maxY, float qminX, float qminY,
                                    int32_t* assigned, float nnratio, int check_ori, int max_dist,
                                    const float* inv_sigma2, double chi2, double chi2_stereo, int32_t* best_idx, int32_t* best_dist);
void  orc_grid_set_query_origin(OrcGrid* g, float qminX, float qminY);
/* modes 0 / 1 on a two-camera frame (Frame::Nleft != -1; R/src/ORBmatcher.cc:144-213, :2093-2160): see orb_oracle.c */
int   orc_search_by_projection_rig(int mode, const OrcProjQuery* ql, const OrcProjQuery* qr, const uint8_t* qdesc, int nq,
                                   const OrcKeyPoint* k2, const uint8_t* d2, int nL, int nR, const int32_t* l2r, const int32_t* r2l,
                                   float minX, float maxX, float minY, float maxY, int32_t* assigned, float nnratio, int check_ori, int max_dist);

/* Frame::ComputeStereoMatches descriptor part (Frame.cc:785-868): per left keypoint the best right
 * index and distance (dist starts at TH_HIGH=100; idx -1 if none < 100). nrows = level-0 rows. */
void  orc_stereo_band_match(const OrcKeyPoint* kl, const uint8_t* dl, int nl,
                            const OrcKeyPoint* kr, const uint8_t* dr, int nr,
                            const float* scale_factors, int nrows, float minD, float maxD,
                            int32_t* best_idx, int32_t* best_dist);

/* Generic candidate matching used by the host-side searches (SearchByBoW ORBmatcher.cc:269-471/819-959,
 * SearchForTriangulation :961-1394, Fuse :1395-1742, SearchBySim3 :1744-1968): for query i the candidates are
 * indices[offsets[i] .. offsets[i+1]) into the train descriptors, visited in list order; strict '<' updates give the
 * top-2 by (distance, list position).  out_idx/out_dist are nq x 2 (train index / distance, -1 when missing). */
void  orc_match_candidates(const uint8_t* q, int nq, const uint8_t* t, const int32_t* offsets, const int32_t* indices,
                           int32_t* out_idx, int32_t* out_dist);

/* Frame::ComputeStereoMatches in full (Frame.cc:785-962): descriptor search, 11x11 SAD sliding-window refinement on
 * the two extractors' pyramids (their last orc_extract), parabola sub-pixel fit, median-based outlier cut.
 * uright/depth have nl entries (-1 = no match); sad_dist (optional) receives the SAD best distance or -1. */
void  orc_compute_stereo_matches(const OrcExtractor* left, const OrcExtractor* right,
                                 const OrcKeyPoint* kl, const uint8_t* dl, int nl,
                                 const OrcKeyPoint* kr, const uint8_t* dr, int nr,
                                 float mb, float mbf, float* uright, float* depth, int32_t* sad_dist);

/* ---- bag of words: DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB> as ORB-SLAM3 uses it ----
 * Restated from R/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h (loadFromTextFile, transform :1127-1200, :1218-1259),
 * BowVector.cpp (addWeight, normalize), FeatureVector.cpp (addFeature), FORB.cpp:81-101 (distance).
 * Vocabulary = node table in file order: node 0 is the root; node i > 0 has parent[i] < i; children keep their order of
 * appearance; nodes flagged is_leaf are the words, numbered in node order.  TF_IDF weighting, L1 scoring (ORBvoc.txt). */
typedef struct OrcVocab OrcVocab;
OrcVocab* orc_vocab_create(int n_nodes, const int32_t* parent, const uint8_t* is_leaf, const uint8_t* desc,
                           const double* weight, int L);
void  orc_vocab_destroy(OrcVocab* v);
/* per-feature descent (:1218-1259): word id, word weight and the node at level L - levelsup (0 = root when that level
 * is <= 0; the last node reached when the leaf lies above that level, where the reference leaves it uninitialised) */
void  orc_bow_transform_features(const OrcVocab* v, const uint8_t* desc, int n, int levelsup,
                                 int32_t* word_id, double* weight, int32_t* node_id);
/* transform(features, BowVector&, FeatureVector&, levelsup) (:1127-1200): bow_words/bow_values sorted by word id and L1
 * normalised (capacity n), fv_nodes sorted node ids with fv_start[i] .. fv_start[i+1] into fv_features (feature indices
 * in insertion order).  Returns the BowVector size; *n_fv = FeatureVector size. */
int   orc_bow_transform(const OrcVocab* v, const uint8_t* desc, int n, int levelsup,
                        int32_t* bow_words, double* bow_values,
                        int32_t* fv_nodes, int32_t* fv_start, int32_t* fv_features, int* n_fv);

/* Frame::UndistortKeyPoints (R/src/Frame.cc:721-754): cv::undistortPoints(pts, K, distCoef, R = I, P = K_new) as OpenCV 4.x
 * computes it (double arithmetic, exactly 5 fixed-point iterations of the radial-tangential model, result rounded to
 * float).  K and P are 3x3 row-major float; dist holds ndist = 4 or 5 coefficients (k1, k2, p1, p2[, k3]).  The x / y
 * of every keypoint is replaced, the other fields are copied.  dist[0] == 0 copies the keypoints unchanged (:723-727). */
void  orc_undistort_keypoints(const OrcKeyPoint* kps, int n, const float* K, const float* dist, int ndist, const float* P,
                              OrcKeyPoint* out);

/* KF.msg wire format of the keypoints (R/../msg/CvKeyPoint.msg; Converter::toCvKeyPointMsg / fromCvKeyPointMsg,
 * R/src/Converter.cc:218-244; used by KeyFrame.cc:1430 and :1929): ROS1 serialises the record without padding as
 * float32 x, float32 y, uint8 size, float32 angle, uint8 response, int8 octave = 15 bytes; size and response are
 * truncated to 8 bits on the way out, class_id is not transmitted (-1 after unpacking). */
void  orc_keypoints_to_msg(const OrcKeyPoint* kps, int n, uint8_t* msg15);
void  orc_keypoints_from_msg(const uint8_t* msg15, int n, OrcKeyPoint* kps);

/* ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo, bCoarse) (R/src/ORBmatcher.cc:961-1202),
 * pinhole cameras without a second camera (mpCamera2 == NULL): for every feature of keyframe 1 without a MapPoint, the
 * features of keyframe 2 in the same vocabulary node that have no MapPoint either; Hamming distance <= TH_LOW and <= the
 * best so far (a later candidate with the same distance replaces an earlier one); monocular-monocular candidates closer
 * to the epipole than 100 * scaleFactor[octave2] (squared pixels) are skipped; the candidate must lie within
 * 3.84 * levelSigma2[octave2] of the epipolar line of F12 (Pinhole::epipolarConstrain, CameraModels/Pinhole.cpp:121-143)
 * unless coarse; rotation consistency as everywhere.  This fork never sets vbMatched2, so queries do not interact.
 * free1 / free2: 1 = the feature has no MapPoint; stereo1 / stereo2 (may be NULL): mvuRight >= 0.
 * matches12[i1] = index in keyframe 2 or -1; returns nmatches. */
int   orc_search_for_triangulation(const OrcKeyPoint* k1, const uint8_t* d1, const uint8_t* free1, const uint8_t* stereo1, int n1,
                                   const int32_t* fv1_nodes, const int32_t* fv1_start, const int32_t* fv1_feat, int nfv1,
                                   const OrcKeyPoint* k2, const uint8_t* d2, const uint8_t* free2, const uint8_t* stereo2, int n2,
                                   const int32_t* fv2_nodes, const int32_t* fv2_start, const int32_t* fv2_feat, int nfv2,
                                   const float* F12, float ep_x, float ep_y, const float* scale_factors2, const float* level_sigma2_2,
                                   int only_stereo, int coarse, int check_ori, int32_t* matches12);

/* SearchByBoW(KeyFrame*, Frame&) on a two-camera frame (R/src/ORBmatcher.cc:344-431): see orb_oracle.c */
int   orc_search_by_bow_rig(const OrcKeyPoint* k1, const uint8_t* d1, const uint8_t* valid1, int n1,
                            const int32_t* fv1_nodes, const int32_t* fv1_start, const int32_t* fv1_feat, int nfv1,
                            const OrcKeyPoint* k2, const uint8_t* d2, int n2, int n2_left,
                            const int32_t* fv2_nodes, const int32_t* fv2_start, const int32_t* fv2_feat, int nfv2,
                            float nnratio, int check_ori, int32_t* matches12l, int32_t* matches12r);
/* MapPoint::ComputeDistinctiveDescriptors (R/src/MapPoint.cc:448-524) for a batch of map points: the observed
 * descriptors of point p are rows offsets[p] .. offsets[p+1] of desc; best[p] = index inside that run of the descriptor
 * with the least median distance to the others (median = sorted row [(int)(0.5 * (N - 1))], the row includes the 0 on
 * the diagonal; first minimum wins); -1 for a point without observations. */
void  orc_distinctive_descriptors(const uint8_t* desc, const int32_t* offsets, int npoints, int32_t* best);

/* ORBmatcher::SearchByBoW on flat arrays.  FeatureVectors are CSR tables sorted by node id (std::map order):
 * fv_nodes[nfv], fv_start[nfv+1], fv_feat[].  valid1 / valid2 = "has a MapPoint that is not bad" (valid2 NULL = every
 * feature is a candidate).  matches12[i1] = matched index in set 2 or -1; returns nmatches.
 * mode 0 = SearchByBoW(KeyFrame*, Frame&, ...)  (ORBmatcher.cc:269-471, Nleft == -1 branch): accept bestDist1 <= TH_LOW;
 *          set 1 = keyframe, set 2 = frame (the reference stores the result per frame feature: vpMapPointMatches[i2]);
 * mode 1 = SearchByBoW(KeyFrame*, KeyFrame*, ...) (ORBmatcher.cc:819-959): candidates need valid2, accept bestDist1 < TH_LOW. */
int   orc_search_by_bow(int mode, const OrcKeyPoint* k1, const uint8_t* d1, const uint8_t* valid1, int n1,
                        const int32_t* fv1_nodes, const int32_t* fv1_start, const int32_t* fv1_feat, int nfv1,
                        const OrcKeyPoint* k2, const uint8_t* d2, const uint8_t* valid2, int n2,
                        const int32_t* fv2_nodes, const int32_t* fv2_start, const int32_t* fv2_feat, int nfv2,
                        float nnratio, int check_ori, int32_t* matches12);

#ifdef __cplusplus
}
#endif
#endif
