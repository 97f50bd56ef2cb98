// orbx_on_oracle.cc - TEST INFRASTRUCTURE.  A stand-in for the subset of the C ABI (include/orbx.h) that the drop-in classes call,
// implemented on the CPU oracle (oracle/orb_oracle.h).  It exists for ONE purpose: to run the host logic of dropin/ORBmatcher.cc,
// ORBextractor.cc and ORBVocabulary.cc (query marshalling, chunking, map bookkeeping) against the reference's own ORBmatcher.cc in
// the CPU test suite, where there is no GPU.  It is linked only into tests/scenario/_build/scenario_dropin_cpu and never ships:
// the product library is multi_orbslam3_b200/liborbx_b200.so, which has no CPU path.  The GPU suite runs the same scenario
// against the real library (scenario_dropin_gpu).
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "orbx.h"
#include "../../oracle/orb_oracle.h"

static thread_local std::string g_err;
extern "C" const char* orbx_last_error(void) { return g_err.c_str(); }
extern "C" int orbx_device_count(void) { return 1; }
extern "C" unsigned long long orbx_launch_count(void) { return 0; }

struct orbx_extractor { OrcExtractor* e; orbx_params p; int w, h; std::vector<std::vector<uint8_t> > staged; };
struct orbx_matcher { orbx_matcher_params p; };
struct orbx_vocab { OrcVocab* v; int words; };

extern "C" int orbx_extractor_create(const orbx_params* p, orbx_extractor** out)
{
    orbx_extractor* h = new orbx_extractor();
    h->p = *p; h->w = h->h = 0;
    h->e = orc_extractor_create(p->nfeatures, p->scale_factor, p->nlevels, p->ini_th_fast, p->min_th_fast);
    *out = h;
    return ORBX_OK;
}
extern "C" void orbx_extractor_destroy(orbx_extractor* h) { if (h) { orc_extractor_destroy(h->e); delete h; } }
extern "C" int orbx_extractor_tables(const orbx_extractor* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int32_t* fpl)
{
    int32_t umax[16];
    std::vector<float> a(h->p.nlevels), b(h->p.nlevels), c(h->p.nlevels), d(h->p.nlevels);
    std::vector<int32_t> f(h->p.nlevels);
    orc_extractor_tables(h->e, a.data(), b.data(), c.data(), d.data(), f.data(), umax);
    for (int i = 0; i < h->p.nlevels; i++) {
        if (scale) scale[i] = a[i];
        if (inv_scale) inv_scale[i] = b[i];
        if (sigma2) sigma2[i] = c[i];
        if (inv_sigma2) inv_sigma2[i] = d[i];
        if (fpl) fpl[i] = f[i];
    }
    return ORBX_OK;
}
extern "C" int orbx_extractor_max_keypoints(const orbx_extractor* h) { return h->p.nfeatures + 4 * h->p.nlevels + 64; }
extern "C" int orbx_extract(orbx_extractor* h, const uint8_t* img, int width, int height, int stride, int lap0, int lap1,
                            orbx_keypoint* kps, uint8_t* desc, int cap, int* n, int* mono_index)
{
    if (!img || width <= 0 || height <= 0) return ORBX_E_EMPTY;
    int nn = 0;
    const int mono = orc_extract(h->e, img, width, height, stride, lap0, lap1, reinterpret_cast<OrcKeyPoint*>(kps), desc, cap, &nn);
    h->w = width; h->h = height;
    if (n) *n = nn;
    if (mono_index) *mono_index = mono;
    return mono == -2 ? ORBX_E_CAPACITY : ORBX_OK;
}
extern "C" int orbx_pyramid_level_size(const orbx_extractor* h, int level, int* width, int* height)
{
    return orc_level_size(h->e, level, width, height) == 0 ? ORBX_OK : ORBX_E_INVALID;
}
extern "C" int orbx_pyramid_to_host(orbx_extractor* h, int slot, int level, uint8_t* dst, int dst_stride);
extern "C" int orbx_pyramid_level_size(const orbx_extractor* h, int level, int* width, int* height);
extern "C" int orbx_pyramid_levels_staged(orbx_extractor* h, int slot, int first_level, int n_levels, const uint8_t** ptr, int* stride)
{
    std::vector<std::vector<uint8_t> >& keep = h->staged;                 // owned by the handle, like the pinned staging of the library
    keep.assign(n_levels, std::vector<uint8_t>());
    for (int k = 0; k < n_levels; k++) {
        int w = 0, hh = 0;
        if (orbx_pyramid_level_size(h, first_level + k, &w, &hh)) return ORBX_E_INVALID;
        keep[k].resize((size_t)w * hh);
        const int rc = orbx_pyramid_to_host(h, slot, first_level + k, keep[k].data(), w);
        if (rc) return rc;
        ptr[k] = keep[k].data(); stride[k] = w;
    }
    return ORBX_OK;
}
extern "C" int orbx_pyramid_levels_to_host(orbx_extractor* h, int slot, int first_level, int n_levels, uint8_t* const* dst, const int* dst_stride)
{
    for (int k = 0; k < n_levels; k++) { const int rc = orbx_pyramid_to_host(h, slot, first_level + k, dst[k], dst_stride[k]); if (rc) return rc; }
    return ORBX_OK;
}
extern "C" int orbx_pyramid_to_host(orbx_extractor* h, int slot, int level, uint8_t* dst, int dst_stride)
{
    (void)slot;
    int w = 0, hh = 0;
    if (orc_level_size(h->e, level, &w, &hh) != 0) return ORBX_E_INVALID;
    const uint8_t* src = orc_level_image(h->e, level);
    for (int y = 0; y < hh; y++) memcpy(dst + (size_t)y * dst_stride, src + (size_t)y * w, w);
    return ORBX_OK;
}

extern "C" int orbx_matcher_create(const orbx_matcher_params* p, orbx_matcher** out) { *out = new orbx_matcher(); (*out)->p = *p; return ORBX_OK; }
extern "C" void orbx_matcher_destroy(orbx_matcher* m) { delete m; }

extern "C" int orbx_hamming_pairs(orbx_matcher*, const uint8_t* a, const uint8_t* b, int n, int32_t* out)
{
    for (int i = 0; i < n; i++) out[i] = orc_hamming256(a + (size_t)i * 32, b + (size_t)i * 32);
    return ORBX_OK;
}

extern "C" int orbx_bf_knn2(orbx_matcher*, const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx, int32_t* dist)
{
    orc_bf_knn2(q, nq, t, nt, idx, dist);
    return ORBX_OK;
}
extern "C" int orbx_search_for_initialization(orbx_matcher* m, const orbx_keypoint* k1, const uint8_t* d1, int n1, const orbx_keypoint* k2,
                                              const uint8_t* d2, int n2, const float bounds[4], float* prev_xy, int32_t* matches12, int window,
                                              float nnratio, int check_ori, int* nmatches)
{
    if (n1 > m->p.max_keypoints || n2 > m->p.max_keypoints) { g_err = "more keypoints than max_keypoints"; return ORBX_E_INVALID; }
    const int nm = orc_search_for_initialization(reinterpret_cast<const OrcKeyPoint*>(k1), d1, n1, reinterpret_cast<const OrcKeyPoint*>(k2), d2, n2,
                                                 bounds[0], bounds[1], bounds[2], bounds[3], prev_xy, matches12, window, nnratio, check_ori);
    if (nmatches) *nmatches = nm;
    return ORBX_OK;
}

static_assert(sizeof(orbx_proj_query) == sizeof(OrcProjQuery), "query layouts");
extern "C" int orbx_search_by_projection_opts(orbx_matcher* m, int mode, const orbx_proj_query* q, const uint8_t* qdesc, int nq,
                                              const orbx_keypoint* k2, const uint8_t* d2, const float* uright2, int n2,
                                              const orbx_proj_options* o, int32_t* assigned, int32_t* best_idx, int32_t* best_dist, int* nmatches)
{
    if (nq > m->p.max_keypoints || n2 > m->p.max_keypoints) { g_err = "more keypoints than max_keypoints"; return ORBX_E_INVALID; }
    std::vector<int32_t> bi(nq > 0 ? nq : 1), bd(nq > 0 ? nq : 1);
    const int nm = orc_search_by_projection_full(mode, reinterpret_cast<const OrcProjQuery*>(q), qdesc, nq, reinterpret_cast<const OrcKeyPoint*>(k2), d2,
                                                 uright2, n2, o->bounds[0], o->bounds[1], o->bounds[2], o->bounds[3], o->query_origin[0],
                                                 o->query_origin[1], assigned, o->nnratio, o->check_ori, o->max_dist,
                                                 o->chi2_mono > 0 ? o->inv_level_sigma2 : NULL, o->chi2_mono, o->chi2_stereo, bi.data(), bd.data());
    if (mode == 3) for (int i = 0; i < nq; i++) { best_idx[i] = bi[i]; best_dist[i] = bd[i]; }
    if (nmatches) *nmatches = nm;
    return ORBX_OK;
}

extern "C" int orbx_search_by_projection_rig(orbx_matcher* m, int mode, const orbx_proj_query* ql, const orbx_proj_query* qr, const uint8_t* qdesc, int nq,
                                             const orbx_keypoint* k2, const uint8_t* d2, int n_left, int n_right, const int32_t* l2r, const int32_t* r2l,
                                             const orbx_proj_options* o, int32_t* assigned, int* nmatches)
{
    if (nq > m->p.max_keypoints || n_left + n_right > m->p.max_keypoints) { g_err = "more keypoints than max_keypoints"; return ORBX_E_INVALID; }
    const int nm = orc_search_by_projection_rig(mode, reinterpret_cast<const OrcProjQuery*>(ql), reinterpret_cast<const OrcProjQuery*>(qr), qdesc, nq,
                                                reinterpret_cast<const OrcKeyPoint*>(k2), d2, n_left, n_right, l2r, r2l, o->bounds[0], o->bounds[1],
                                                o->bounds[2], o->bounds[3], assigned, o->nnratio, o->check_ori, o->max_dist);
    if (nmatches) *nmatches = nm;
    return ORBX_OK;
}

extern "C" int orbx_search_by_bow(orbx_matcher*, int mode, const orbx_keypoint* k1, const uint8_t* d1, const uint8_t* valid1, int n1,
                                  const int32_t* fv1_nodes, const int32_t* fv1_start, const int32_t* fv1_feat, int nfv1,
                                  const orbx_keypoint* k2, const uint8_t* d2, const uint8_t* valid2, int n2,
                                  const int32_t* fv2_nodes, const int32_t* fv2_start, const int32_t* fv2_feat, int nfv2,
                                  float nnratio, int check_ori, int32_t* matches12, int* nmatches)
{
    const int nm = orc_search_by_bow(mode, reinterpret_cast<const OrcKeyPoint*>(k1), d1, valid1, n1, fv1_nodes, fv1_start, fv1_feat, nfv1,
                                     reinterpret_cast<const OrcKeyPoint*>(k2), d2, valid2, n2, fv2_nodes, fv2_start, fv2_feat, nfv2,
                                     nnratio, check_ori, matches12);
    if (nmatches) *nmatches = nm;
    return ORBX_OK;
}

extern "C" int orbx_search_by_bow_rig(orbx_matcher*, const orbx_keypoint* k1, const uint8_t* d1, const uint8_t* valid1, int n1,
                                      const int32_t* fv1_nodes, const int32_t* fv1_start, const int32_t* fv1_feat, int nfv1,
                                      const orbx_keypoint* k2, const uint8_t* d2, int n2, int n2_left,
                                      const int32_t* fv2_nodes, const int32_t* fv2_start, const int32_t* fv2_feat, int nfv2,
                                      float nnratio, int check_ori, int32_t* ml, int32_t* mr, int* nmatches)
{
    const int nm = orc_search_by_bow_rig(reinterpret_cast<const OrcKeyPoint*>(k1), d1, valid1, n1, fv1_nodes, fv1_start, fv1_feat, nfv1,
                                         reinterpret_cast<const OrcKeyPoint*>(k2), d2, n2, n2_left, fv2_nodes, fv2_start, fv2_feat, nfv2,
                                         nnratio, check_ori, ml, mr);
    if (nmatches) *nmatches = nm;
    return ORBX_OK;
}

extern "C" int orbx_stereo_matches(orbx_matcher*, orbx_extractor* left, orbx_extractor* right, int, int, int, int, float mb, float mbf,
                                   float* uright, float* depth, int32_t* sad_dist, int cap, int* n_left)
{
    // the oracle keeps the last extraction's keypoints per level, not the final arrays: re-run is not possible here, so the
    // stand-in extracts nothing itself; the scenario passes through Frame::ComputeStereoMatches (reference body) on CPU instead.
    (void)left; (void)right; (void)mb; (void)mbf; (void)uright; (void)depth; (void)sad_dist; (void)cap; (void)n_left;
    g_err = "orbx_stereo_matches is not available in the oracle-backed stand-in";
    return ORBX_E_INVALID;
}

extern "C" int orbx_vocab_create(int, int n_nodes, const int32_t* parent, const uint8_t* is_leaf, const uint8_t* desc, const double* weight, int L,
                                 orbx_vocab** out)
{
    OrcVocab* v = orc_vocab_create(n_nodes, parent, is_leaf, desc, weight, L);
    if (!v) { g_err = "invalid vocabulary"; return ORBX_E_INVALID; }
    *out = new orbx_vocab();
    (*out)->v = v; (*out)->words = 0;
    for (int i = 0; i < n_nodes; i++) (*out)->words += is_leaf[i] ? 1 : 0;
    return ORBX_OK;
}
extern "C" void orbx_vocab_destroy(orbx_vocab* v) { if (v) { orc_vocab_destroy(v->v); delete v; } }
extern "C" int orbx_vocab_words(const orbx_vocab* v) { return v->words; }
extern "C" int orbx_bow_transform(orbx_vocab* v, const uint8_t* desc, int n, int levelsup, int32_t* word_id, double* weight, int32_t* node_id)
{
    std::vector<double> w(n > 0 ? n : 1);
    orc_bow_transform_features(v->v, desc, n, levelsup, word_id, weight ? weight : w.data(), node_id);
    return ORBX_OK;
}
