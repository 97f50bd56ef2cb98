// Drop-in ORBVocabulary over the B200 C ABI (include/orbx.h); see ORBVocabulary.h.
#include "ORBVocabulary.h"
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include "orbx.h"

namespace ORB_SLAM3
{

int ORBVocabulary::sDevice = 0;
void ORBVocabulary::SetDevice(int device) { sDevice = device; }

ORBVocabulary::ORBVocabulary() : mpHandle(nullptr), m_k(0), m_L(0) {}
ORBVocabulary::~ORBVocabulary() { if (mpHandle) orbx_vocab_destroy(mpHandle); }

bool ORBVocabulary::loadFromTextFile(const std::string &filename)
{
    std::ifstream f(filename.c_str());
    if (!f.is_open()) return false;
    std::string s;
    std::getline(f, s);
    std::stringstream ss(s);
    int n1 = -1, n2 = -1;
    ss >> m_k >> m_L >> n1 >> n2;
    if (m_k < 0 || m_k > 20 || m_L < 1 || m_L > 10 || n1 < 0 || n1 > 5 || n2 < 0 || n2 > 3) return false;   // TemplatedVocabulary.h, same check
    if (n1 != 0 || n2 != 0) return false;        // L1_NORM + TF_IDF only (what ORBvoc.txt declares)
    std::vector<int32_t> parent(1, -1);
    std::vector<uint8_t> leaf(1, 0), desc(32, 0);
    std::vector<double> weight(1, 0.0);
    while (std::getline(f, s)) {
        if (s.empty()) continue;
        std::stringstream sn(s);
        int pid = -1, isLeaf = 0;
        sn >> pid >> isLeaf;
        uint8_t d[32];
        for (int i = 0; i < 32; i++) { int b = 0; sn >> b; d[i] = (uint8_t)b; }          // FORB::fromString
        double w = 0.0;
        sn >> w;
        if (sn.fail()) return false;
        parent.push_back(pid); leaf.push_back(isLeaf > 0 ? 1 : 0); weight.push_back(w);
        desc.insert(desc.end(), d, d + 32);
    }
    if (mpHandle) { orbx_vocab_destroy(mpHandle); mpHandle = nullptr; }
    if (orbx_vocab_create(sDevice, (int)parent.size(), parent.data(), leaf.data(), desc.data(), weight.data(), m_L, &mpHandle) != ORBX_OK)
        throw std::runtime_error(std::string("ORBVocabulary (B200): ") + orbx_last_error());      // no CPU fallback exists
    return true;
}

unsigned int ORBVocabulary::size() const { return mpHandle ? (unsigned int)orbx_vocab_words(mpHandle) : 0u; }
bool ORBVocabulary::empty() const { return size() == 0; }

void ORBVocabulary::assemble(int n, const int* word, const double* weight, const int* node, DBoW2::BowVector &v, DBoW2::FeatureVector &fv) const
{
    for (int i = 0; i < n; i++) {
        if (!(weight[i] > 0)) continue;                                   // stopped word (:1157)
        DBoW2::BowVector::iterator vit = v.lower_bound((DBoW2::WordId)word[i]);        // BowVector::addWeight
        if (vit != v.end() && vit->first == (DBoW2::WordId)word[i]) vit->second += weight[i];
        else v.insert(vit, DBoW2::BowVector::value_type((DBoW2::WordId)word[i], weight[i]));
        fv[(DBoW2::NodeId)node[i]].push_back((unsigned int)i);            // FeatureVector::addFeature
    }
    double norm = 0.0;                                                    // BowVector::normalize(L1)
    for (DBoW2::BowVector::iterator it = v.begin(); it != v.end(); ++it) norm += std::fabs(it->second);
    if (norm > 0.0)
        for (DBoW2::BowVector::iterator it = v.begin(); it != v.end(); ++it) it->second /= norm;
}

void ORBVocabulary::transform(const cv::Mat& descriptors, DBoW2::BowVector &v, DBoW2::FeatureVector &fv, int levelsup) const
{
    v.clear(); fv.clear();
    if (empty() || descriptors.rows == 0) return;
    const int n = descriptors.rows;
    std::vector<uint8_t> tmp;
    const uint8_t* d = descriptors.ptr(0);
    if (!descriptors.isContinuous()) {
        tmp.resize((size_t)n * 32);
        for (int i = 0; i < n; i++) std::memcpy(tmp.data() + (size_t)i * 32, descriptors.ptr(i), 32);
        d = tmp.data();
    }
    std::vector<int32_t> word(n), node(n);
    std::vector<double> weight(n);
    if (orbx_bow_transform(mpHandle, d, n, levelsup, word.data(), weight.data(), node.data()) != ORBX_OK)
        throw std::runtime_error(std::string("ORBVocabulary (B200): ") + orbx_last_error());
    assemble(n, word.data(), weight.data(), node.data(), v, fv);
}

void ORBVocabulary::transform(const std::vector<cv::Mat>& features, DBoW2::BowVector &v, DBoW2::FeatureVector &fv, int levelsup) const
{
    v.clear(); fv.clear();
    if (empty() || features.empty()) return;
    const int n = (int)features.size();
    std::vector<uint8_t> rows((size_t)n * 32);
    for (int i = 0; i < n; i++) std::memcpy(rows.data() + (size_t)i * 32, features[i].ptr(0), 32);
    cv::Mat m(n, 32, CV_8U, rows.data());
    transform(m, v, fv, levelsup);
}

} //namespace ORB_SLAM
