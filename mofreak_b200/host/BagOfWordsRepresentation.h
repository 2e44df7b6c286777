// Host-side mirror of the reference's BagOfWordsRepresentation (src/MoFREAK/BagOfWordsRepresentation.h:23-87) for the
// path main.cpp:294-349 drives: constructor (loads SVM_PATH/clusters.txt), intializeBOWMemory,
// convertFileToBOWFeature, computeSlidingBagOfWords, writeBOWFeaturesToFiles.  The per-row arg-min Hamming
// assignment (bruteForceMatch / hammingDistance, BagOfWordsRepresentation.cpp:22-72) runs on the GPU through
// mfk_bow_histogram / mfk_bow_assign; text parsing, float normalisation and libsvm line formatting stay on the host and
// are byte-compatible.  convertFileToBOWFeature may be called from several threads (the reference does so from an
// OpenMP region, main.cpp:305-338, with unsynchronised push_backs); here it is internally serialised.
#ifndef MOFREAK_B200_BAGOFWORDSREPRESENTATION_H
#define MOFREAK_B200_BAGOFWORDSREPRESENTATION_H

#include <stdint.h>

#include <fstream>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/mofreak_b200.h"

class BagOfWordsRepresentation {
public:
    enum datasets { KTH, TRECVID, HOLLYWOOD, UTI1, UTI2, HMDB51, UCF101 };

    BagOfWordsRepresentation(int num_clust, int ftr_dim, std::string svm_path, int num_groups, int dset);
    ~BagOfWordsRepresentation();
    BagOfWordsRepresentation(const BagOfWordsRepresentation&) = delete;
    BagOfWordsRepresentation& operator=(const BagOfWordsRepresentation&) = delete;

    void intializeBOWMemory(std::string SVM_PATH);       // (sic) reference spelling
    void computeSlidingBagOfWords(std::string& input_file, int alpha, int label, std::ofstream& out);
    void convertFileToBOWFeature(std::string file);
    void writeBOWFeaturesToFiles();

    void setDevice(int device) { device_ = device; }
    // the codebook as loadClusters parsed it (K x 16), for inspection
    const std::vector<uint8_t>& codebook() const { return codebook_; }
    // normalised histogram of one .mofreak file (buildHistogram, :74-138); false for a file without rows
    bool buildHistogram(const std::string& file, std::vector<float>& histogram);

private:
    void extractMetadata(std::string filename, int& action, int& group, int& clip_number);
    void loadClusters();
    void ensureContext();
    [[noreturn]] void die(const std::string& what);

    const unsigned int NUMBER_OF_CLUSTERS;
    const unsigned int FEATURE_DIMENSIONALITY;
    const unsigned int NUMBER_OF_GROUPS;
    const std::string SVM_PATH;
    int dataset;
    int device_;
    mfk_ctx* ctx_;
    std::mutex mutex_;
    std::vector<uint8_t> codebook_;
    std::unordered_map<std::string, int> actions;
    std::vector<std::vector<std::string> > bow_training_crossvalidation_sets;
    std::vector<std::vector<std::string> > bow_testing_crossvalidation_sets;
    std::vector<std::ofstream*> training_files;
    std::vector<std::ofstream*> testing_files;
};

#endif
