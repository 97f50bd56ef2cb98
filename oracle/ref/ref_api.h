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

#ifdef __cplusplus
}
#endif
#endif
