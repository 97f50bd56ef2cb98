// orbx_bow.cu - bag-of-words transform of ORB descriptors.
//
// Replaces the per-feature tree descent of DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>::transform
// (R/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1200 and :1218-1259, distance = FORB.cpp:81-101) that
// Frame::ComputeBoW (R/src/Frame.cc:712-719) and KeyFrame::ComputeBoW (R/src/KeyFrame.cc:168-176) run on every frame /
// keyframe: ~60 Hamming distances per descriptor down a k = 10, L = 6 tree.
//
// Layout: the vocabulary is stored level-agnostic as a CSR over children in order of appearance, with the children's
// descriptors copied into CSR order, so the k candidates of one descent step are k consecutive 32-byte rows (one 320-byte
// run for ORBvoc).  Sixteen lanes serve one descriptor (two descriptors per warp): lane c scores child c (+16, +32 ...
// for wider nodes), the group takes the minimum of (distance << 8 | child rank) by shuffles, i.e. the first child with
// the smallest distance, exactly the strict '<' scan of the reference.  The whole vocabulary (35 MB for ORBvoc) stays in L2.
#include <cstring>
#include <vector>
#include "orbx_internal.h"

#define CKB(call)                                                                         \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            orbx_set_error("%s failed: %s", #call, cudaGetErrorString(e_));               \
            return ORBX_E_CUDA;                                                           \
        }                                                                                 \
    } while (0)

struct orbx_vocab {
    int device, n_nodes, n_words, L;
    int32_t* d_cstart;     // [n_nodes + 1] CSR over children
    int32_t* d_cnode;      // [n_nodes - 1] node id of the child at a CSR position
    uint4* d_cdesc;        // [n_nodes - 1][2] descriptor of the child at a CSR position
    int32_t* d_word;       // [n_nodes] word id, -1 for inner nodes
    std::vector<double> word_weight;   // host: weight of word w
    std::vector<int32_t> node_word;    // host copy of d_word
    std::vector<double> node_weight;
    cudaStream_t stream;
    uint8_t* d_in; int32_t* d_out; size_t in_cap;   // staging of the host entry point
};

namespace {

constexpr int BOW_G = 16;        // lanes per descriptor
constexpr int BOW_NT = 256;

// desc: [n][32]; word / node: [n].  n_ptr (optional): per-slot counts for the slot form (blockIdx.y = slot).
__global__ void __launch_bounds__(BOW_NT) k_bow_transform(const int32_t* cstart, const int32_t* cnode, const uint4* cdesc, const int32_t* word_of,
                                                          const uint8_t* desc, int n_fixed, const int32_t* n_slot, int first_slot, int cap,
                                                          int nid_level, int32_t* word, int32_t* node)
{
    const int slot = blockIdx.y;
    const int n = n_slot ? n_slot[first_slot + slot] : n_fixed;
    const int g = (blockIdx.x * BOW_NT + threadIdx.x) / BOW_G, sub = threadIdx.x & (BOW_G - 1);
    // all 16 lanes of a group share g; groups past the end idle through the shuffles with a dummy descriptor
    const bool live = g < n;
    const size_t row = (size_t)(n_slot ? first_slot + slot : 0) * cap + (live ? g : 0);
    const uint4 q0 = reinterpret_cast<const uint4*>(desc)[2 * row], q1 = reinterpret_cast<const uint4*>(desc)[2 * row + 1];
    int final_id = 0, level = 0, nid = 0;
    bool have = nid_level <= 0;
    int c0 = cstart[0], c1 = cstart[1];
    const unsigned gmask = 0xFFFFu << (threadIdx.x & 16);   // the two groups of a warp may leave the loop at different depths
    while (c1 > c0) {                                      // group-uniform
        ++level;
        unsigned best = 0xFFFFFFFFu;
        for (int c = c0 + sub; c < c1; c += BOW_G) {
            const uint4 t0 = __ldg(cdesc + 2 * (size_t)c), t1 = __ldg(cdesc + 2 * (size_t)c + 1);
            const unsigned d = (unsigned)(__popc(q0.x ^ t0.x) + __popc(q0.y ^ t0.y) + __popc(q0.z ^ t0.z) + __popc(q0.w ^ t0.w) +
                                          __popc(q1.x ^ t1.x) + __popc(q1.y ^ t1.y) + __popc(q1.z ^ t1.z) + __popc(q1.w ^ t1.w));
            const unsigned key = (d << 20) | (unsigned)(c - c0);          // up to 2^20 children per node
            best = min(best, key);
        }
#pragma unroll
        for (int o = BOW_G / 2; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(gmask, best, o));
        const int pos = c0 + (int)(best & 0xFFFFFu);
        final_id = __ldg(cnode + pos);
        if (level == nid_level) { nid = final_id; have = true; }
        c0 = __ldg(cstart + final_id); c1 = __ldg(cstart + final_id + 1);
    }
    if (!have) nid = final_id;
    if (live && sub == 0) {
        const size_t o = (size_t)slot * cap * (n_slot ? 1 : 0) + g;
        word[o] = word_of[final_id]; node[o] = nid;
    }
}

}  // namespace

// DBoW2 vocabulary from its node table (the rows of ORBvoc.txt after the header, in file order, see
// TemplatedVocabulary.h loadFromTextFile): node 0 = root; parent[i] < i; is_leaf marks the words (numbered in node order).
extern "C" int orbx_vocab_create(int device, int n_nodes, const int32_t* parent, const uint8_t* is_leaf, const uint8_t* desc,
                                 const double* weight, int L, orbx_vocab** out)
{
    if (!out) return ORBX_E_INVALID;
    *out = nullptr;
    if (n_nodes < 2 || !parent || !is_leaf || !desc || !weight || L < 1) { orbx_set_error("%s%s", "orbx_vocab_create: invalid arguments", ""); return ORBX_E_INVALID; }
    std::vector<int32_t> cstart((size_t)n_nodes + 1, 0), cnode((size_t)n_nodes - 1), word((size_t)n_nodes, -1);
    for (int i = 1; i < n_nodes; i++) {
        if (parent[i] < 0 || parent[i] >= i) { orbx_set_error("%s%s", "orbx_vocab_create: parent ids must precede their children", ""); return ORBX_E_INVALID; }
        cstart[parent[i] + 1]++;
    }
    for (int i = 0; i < n_nodes; i++) cstart[i + 1] += cstart[i];
    {
        std::vector<int32_t> cur(cstart.begin(), cstart.end() - 1);
        for (int i = 1; i < n_nodes; i++) cnode[cur[parent[i]]++] = i;
    }
    orbx_vocab* v = new orbx_vocab();
    v->device = device; v->n_nodes = n_nodes; v->L = L; v->n_words = 0;
    v->d_cstart = nullptr; v->d_cnode = nullptr; v->d_cdesc = nullptr; v->d_word = nullptr; v->stream = nullptr;
    v->d_in = nullptr; v->d_out = nullptr; v->in_cap = 0;
    for (int i = 1; i < n_nodes; i++) {
        const bool childless = cstart[i + 1] == cstart[i];
        if (childless != (is_leaf[i] != 0)) { delete v; orbx_set_error("%s%s", "orbx_vocab_create: is_leaf must mark exactly the childless nodes", ""); return ORBX_E_INVALID; }
        if (is_leaf[i]) { word[i] = v->n_words++; v->word_weight.push_back(weight[i]); }
    }
    v->node_word = word; v->node_weight.assign(weight, weight + n_nodes);
    std::vector<uint8_t> cdesc((size_t)(n_nodes - 1) * 32);
    for (int c = 0; c < n_nodes - 1; c++) memcpy(cdesc.data() + (size_t)c * 32, desc + (size_t)cnode[c] * 32, 32);
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc((void**)&v->d_cstart, sizeof(int32_t) * ((size_t)n_nodes + 1));
    if (e == cudaSuccess) e = cudaMalloc((void**)&v->d_cnode, sizeof(int32_t) * ((size_t)n_nodes - 1));
    if (e == cudaSuccess) e = cudaMalloc((void**)&v->d_cdesc, (size_t)(n_nodes - 1) * 32);
    if (e == cudaSuccess) e = cudaMalloc((void**)&v->d_word, sizeof(int32_t) * (size_t)n_nodes);
    if (e == cudaSuccess) e = cudaMemcpy(v->d_cstart, cstart.data(), sizeof(int32_t) * ((size_t)n_nodes + 1), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(v->d_cnode, cnode.data(), sizeof(int32_t) * ((size_t)n_nodes - 1), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(v->d_cdesc, cdesc.data(), cdesc.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(v->d_word, word.data(), sizeof(int32_t) * (size_t)n_nodes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        orbx_set_error("%s failed: %s", "orbx_vocab_create (no CPU fallback exists)", cudaGetErrorString(e));
        orbx_vocab_destroy(v);
        return ORBX_E_CUDA;
    }
    *out = v;
    return ORBX_OK;
}

extern "C" void orbx_vocab_destroy(orbx_vocab* v)
{
    if (!v) return;
    cudaSetDevice(v->device);
    if (v->d_cstart) cudaFree(v->d_cstart);
    if (v->d_cnode) cudaFree(v->d_cnode);
    if (v->d_cdesc) cudaFree(v->d_cdesc);
    if (v->d_word) cudaFree(v->d_word);
    if (v->d_in) cudaFree(v->d_in);
    if (v->d_out) cudaFree(v->d_out);
    if (v->stream) cudaStreamDestroy(v->stream);
    delete v;
}

extern "C" int orbx_vocab_words(const orbx_vocab* v) { return v ? v->n_words : 0; }

// weight of every word (idf for ORBvoc), n_words doubles
extern "C" int orbx_vocab_word_weights(const orbx_vocab* v, double* w, int cap)
{
    if (!v || !w || cap < v->n_words) return ORBX_E_INVALID;
    memcpy(w, v->word_weight.data(), sizeof(double) * v->n_words);
    return ORBX_OK;
}

static int bow_launch(orbx_vocab* v, const uint8_t* d_desc, int n_fixed, const int32_t* d_n, int first_slot, int count, int cap,
                      int levelsup, int32_t* d_word, int32_t* d_node, cudaStream_t s)
{
    const int rows = d_n ? cap : n_fixed;
    if (rows <= 0 || count <= 0) return ORBX_OK;
    const dim3 grid((unsigned)(((size_t)rows * BOW_G + BOW_NT - 1) / BOW_NT), (unsigned)count);
    k_bow_transform<<<grid, BOW_NT, 0, s>>>(v->d_cstart, v->d_cnode, v->d_cdesc, v->d_word, d_desc, n_fixed, d_n, first_slot, cap,
                                            v->L - levelsup, d_word, d_node);
    ORBX_COUNT_LAUNCH(1);
    CKB(cudaGetLastError());
    return ORBX_OK;
}

// Per-feature part of transform(features, BowVector&, FeatureVector&, levelsup) (:1127-1200, :1218-1259): for every descriptor
// its word id, the word's weight and the id of the node at level L - levelsup.  Host pointers; synchronous.
extern "C" int orbx_bow_transform(orbx_vocab* v, const uint8_t* desc, int n, int levelsup, int32_t* word_id, double* weight, int32_t* node_id)
{
    if (!v || n < 0 || (n > 0 && (!desc || !word_id || !node_id))) return ORBX_E_INVALID;
    if (n == 0) return ORBX_OK;
    CKB(cudaSetDevice(v->device));
    if ((size_t)n > v->in_cap) {
        if (v->d_in) cudaFree(v->d_in);
        if (v->d_out) cudaFree(v->d_out);
        v->d_in = nullptr; v->d_out = nullptr; v->in_cap = 0;
        CKB(cudaMalloc((void**)&v->d_in, (size_t)n * 32));
        CKB(cudaMalloc((void**)&v->d_out, sizeof(int32_t) * 2 * (size_t)n));
        v->in_cap = (size_t)n;
    }
    cudaStream_t s = v->stream;
    CKB(cudaMemcpyAsync(v->d_in, desc, (size_t)n * 32, cudaMemcpyHostToDevice, s));
    int rc = bow_launch(v, v->d_in, n, nullptr, 0, 1, n, levelsup, v->d_out, v->d_out + n, s);
    if (rc) return rc;
    CKB(cudaMemcpyAsync(word_id, v->d_out, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
    CKB(cudaMemcpyAsync(node_id, v->d_out + n, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
    CKB(cudaStreamSynchronize(s));
    if (weight)
        for (int i = 0; i < n; i++) weight[i] = word_id[i] >= 0 ? v->word_weight[word_id[i]] : 0.0;
    return ORBX_OK;
}

// The same on the descriptors an extractor holds in its result slots first_slot .. first_slot+count-1 (Frame::ComputeBoW
// right after ORBextractor::operator(), descriptors never leave the GPU): d_word / d_node are DEVICE arrays
// [count][orbx_extractor_max_keypoints(ex)]; entries past a slot's keypoint count are left untouched.  Asynchronous on `stream`.
extern "C" int orbx_bow_transform_slots_device(orbx_vocab* v, orbx_extractor* ex, int first_slot, int count, int levelsup,
                                               int32_t* d_word, int32_t* d_node, void* stream)
{
    if (!v || !ex || !d_word || !d_node || count <= 0) return ORBX_E_INVALID;
    orbx_keypoint* dk; uint8_t* dd; int32_t* dn; int cap, slots;
    int rc = orbx_extractor_results_device(ex, &dk, &dd, &dn, nullptr, &cap, &slots);
    if (rc) return rc;
    if (first_slot < 0 || first_slot + count > slots || orbx_ex_device(ex) != v->device) return ORBX_E_INVALID;
    CKB(cudaSetDevice(v->device));
    return bow_launch(v, dd, 0, dn, first_slot, count, cap, levelsup, d_word, d_node, stream ? (cudaStream_t)stream : orbx_ex_stream(ex));
}
