// Drop-in replacement of R/orb_slam3/include/ORBVocabulary.h
//   typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> ORBVocabulary;
// for the calls ORB-SLAM3 makes on it: loadFromTextFile (System / ClientSystem construction), transform(features,
// BowVector&, FeatureVector&, levelsup) (Frame::ComputeBoW R/src/Frame.cc:712-719, KeyFrame::ComputeBoW
// R/src/KeyFrame.cc:168-176), size(), empty().  The tree descent runs on the B200 (orbx_bow_transform, include/orbx.h);
// the two small maps are assembled on the host exactly as BowVector::addWeight / normalize and
// FeatureVector::addFeature do.  score() (KeyFrameDatabase) is untouched DBoW2 code operating on BowVector.
#ifndef ORBVOCABULARY_H
#define ORBVOCABULARY_H

#include <map>
#include <string>
#include <vector>
#include <opencv2/core/core.hpp>

#include "Thirdparty/DBoW2/DBoW2/BowVector.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"

struct orbx_vocab;   // include/orbx.h

namespace ORB_SLAM3
{

class ORBVocabulary
{
public:
    ORBVocabulary();
    ~ORBVocabulary();

    // Same text format and header checks as TemplatedVocabulary::loadFromTextFile ("k L scoring weighting", then one
    // line per node: parent, isLeaf, 32 descriptor bytes, weight).  Only the combination ORBvoc.txt ships with
    // (L1_NORM scoring, TF_IDF weighting) is implemented; anything else returns false.
    bool loadFromTextFile(const std::string &filename);

    // Number of words / whether a vocabulary is loaded
    unsigned int size() const;
    bool empty() const;

    // TemplatedVocabulary.h:1127-1200
    void transform(const std::vector<cv::Mat>& features, DBoW2::BowVector &v, DBoW2::FeatureVector &fv, int levelsup) const;
    // Addition: the same on the N x 32 descriptor matrix itself (Frame::mDescriptors), sparing Converter::toDescriptorVector
    void transform(const cv::Mat& descriptors, DBoW2::BowVector &v, DBoW2::FeatureVector &fv, int levelsup) const;

    static void SetDevice(int device);

private:
    ORBVocabulary(const ORBVocabulary&);
    ORBVocabulary& operator=(const ORBVocabulary&);
    void assemble(int n, const int* word, const double* weight, const int* node, DBoW2::BowVector &v, DBoW2::FeatureVector &fv) const;

    orbx_vocab* mpHandle;
    int m_k, m_L;
    static int sDevice;
};

} //namespace ORB_SLAM

#endif // ORBVOCABULARY_H
