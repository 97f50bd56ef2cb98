/* ref_api.h - C entry points of oracle/_ref/libref_orb.so: the reference's OWN code (compiled unmodified from
 * /root/reference by oracle/ref/Makefile against the minimal OpenCV stand-in dropin/cvmin) behind the same flat-array
 * interface as the C restatement oracle/orb_oracle.h, so that tests can demand  _ref == oracle == GPU  byte for byte.
 * TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this library. */
#ifndef REF_API_H
#define REF_API_H
#include <stdint.h>
#include "../orb_oracle.h"
#ifdef __cplusplus
extern "C" {
#endif

/* 1 when the list nodes of DistributeOctTree come from the monotonic arena (libref_orb.so), 0 on glibc malloc (libref_orb_malloc.so) */
int   ref_uses_arena(void);

/* ---- ORB_SLAM3::ORBextractor (R/src/ORBextractor.cc, whole file) ---- */
typedef struct RefExtractor RefExtractor;
RefExtractor* ref_extractor_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th);
void  ref_extractor_destroy(RefExtractor* e);
void  ref_extractor_tables(RefExtractor* e, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                           int32_t* features_per_level, int32_t* umax16);
/* operator()(image, mask, keypoints, descriptors, vLappingArea): returns monoIndex (-1 on an empty image) */
int   ref_extract(RefExtractor* e, const uint8_t* img, int w, int h, int stride, int lap0, int lap1,
                  OrcKeyPoint* kps, uint8_t* desc, int cap, int* n_out);
/* mvImagePyramid[level] of the last call */
int   ref_level_size(RefExtractor* e, int level, int* w, int* h);
int   ref_level_image(RefExtractor* e, int level, uint8_t* dst, int dst_stride);
/* ComputePyramid + ComputeKeyPointsOctTree on an image: allKeypoints[level] (border added, octave / size / angle set, not scaled) */
int   ref_octree_keypoints(RefExtractor* e, const uint8_t* img, int w, int h, int stride, int level, OrcKeyPoint* kps, int cap);
/* DistributeOctTree on an explicit candidate list (x, y, response triples) */
int   ref_distribute_octree(RefExtractor* e, const float* xyr, int n, int minX, int maxX, int minY, int maxY, int N, int level,
                            float* out_xyr, int cap);


/* ---- ORB_SLAM3::ORBmatcher (R/src/ORBmatcher.cc, whole file) over the reference's own Frame / KeyFrame / MapPoint bodies ---- */
int   ref_hamming256(const uint8_t* a, const uint8_t* b);                                   /* ORBmatcher::DescriptorDistance */
/* Frame::AssignFeaturesToGrid + GetFeaturesInArea (R/src/Frame.cc:360-391, 628-709) */
int   ref_features_in_area(const OrcKeyPoint* kps, int n, float minX, float maxX, float minY, float maxY,
                           float x, float y, float r, int minLevel, int maxLevel, int32_t* out, int cap);
/* same contract as orc_search_for_initialization */
int   ref_search_for_initialization(const OrcKeyPoint* k1, const uint8_t* d1, int n1, const OrcKeyPoint* k2, const uint8_t* d2, int n2,
                                    float minX, float maxX, float minY, float maxY, float* prev_xy, int32_t* matches12, int window,
                                    float nnratio, int check_ori);
/* Frame::ComputeStereoMatches (R/src/Frame.cc:785-962) on the pyramids of the two extractors' last ref_extract calls */
void  ref_compute_stereo_matches(RefExtractor* left, RefExtractor* right, const OrcKeyPoint* kl, const uint8_t* dl, int nl,
                                 const OrcKeyPoint* kr, const uint8_t* dr, int nr, const float* scale, int nlevels,
                                 float mb, float mbf, float* uright, float* depth);
/* DBoW2 vocabulary from a node table (through TemplatedVocabulary::loadFromTextFile); transform as orc_bow_transform[_features] */
typedef struct RefVocab RefVocab;
RefVocab* ref_vocab_create(int n_nodes, const int32_t* parent, const uint8_t* is_leaf, const uint8_t* desc, const double* weight, int k, int L);
void  ref_vocab_destroy(RefVocab* v);
int   ref_bow_transform(RefVocab* v, const uint8_t* desc, int n, int levelsup, int32_t* bow_words, double* bow_values,
                        int32_t* fv_nodes, int32_t* fv_start, int32_t* fv_features, int* n_fv);
void  ref_bow_transform_features(RefVocab* v, const uint8_t* desc, int n, int levelsup, int32_t* word_id, double* weight, int32_t* node_id);
/* same contract as orc_search_by_bow */
int   ref_search_by_bow(int mode, const OrcKeyPoint* k1, const uint8_t* d1, const uint8_t* valid1, int n1,
                        const int32_t* fv1_nodes, const int32_t* fv1_start, const int32_t* fv1_feat, int nfv1,
                        const OrcKeyPoint* k2, const uint8_t* d2, const uint8_t* valid2, int n2,
                        const int32_t* fv2_nodes, const int32_t* fv2_start, const int32_t* fv2_feat, int nfv2,
                        float nnratio, int check_ori, int32_t* matches12);
/* same contract as orc_distinctive_descriptors (MapPoint::ComputeDistinctiveDescriptors, R/src/MapPoint.cc:448-524) */
void  ref_distinctive_descriptors(const uint8_t* desc, const int32_t* offsets, int npoints, int32_t* best);

#ifdef __cplusplus
}
#endif
#endif
