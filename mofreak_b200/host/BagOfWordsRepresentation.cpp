// See BagOfWordsRepresentation.h.  Reference: src/MoFREAK/BagOfWordsRepresentation.cpp:74-173 (histogram, clusters),
// :176-378 (sliding window), :845-1090 (metadata, cross-validation fan-out, writers), :1145-1155 (constructor).
#include "BagOfWordsRepresentation.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <iostream>
#include <list>
#include <sstream>

namespace {

std::string baseName(const std::string& path) {
    const size_t p = path.find_last_of("/\\");
    return p == std::string::npos ? path : path.substr(p + 1);
}

std::vector<std::string> split(const std::string& s, char delim) {
    std::vector<std::string> elems;
    std::stringstream ss(s);
    std::string item;
    while (std::getline(ss, item, delim)) elems.push_back(item);
    return elems;
}

// Parses .mofreak text: per line 6 numbers to skip, then FEATURE_DIMENSIONALITY numbers read as float and cast to
// unsigned char (BagOfWordsRepresentation.cpp:86-105).  Returns frames (3rd number) and descriptors.
void parseMofreak(const std::string& file, unsigned dim, std::vector<int>& frames, std::vector<uint8_t>& desc) {
    std::ifstream input_file(file);
    std::string line;
    while (std::getline(input_file, line)) {
        const char* p = line.c_str();
        char* end;
        double head[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 6; i++) { head[i] = strtod(p, &end); p = end; }
        frames.push_back((int)head[2]);
        for (unsigned i = 0; i < dim; i++) {
            const float elem = strtof(p, &end);
            p = end;
            desc.push_back((unsigned char)elem);
        }
    }
}

// "label 1:v 2:v ... " with operator<< float formatting (%g)
std::string libsvmLine(int label, const float* hist, unsigned n) {
    std::string s;
    char buf[64];
    s.append(buf, snprintf(buf, sizeof buf, "%d ", label));
    for (unsigned col = 0; col < n; ++col) s.append(buf, snprintf(buf, sizeof buf, "%d:%g ", (int)(col + 1), (double)hist[col]));
    return s;
}

}  // namespace

BagOfWordsRepresentation::BagOfWordsRepresentation(int num_clust, int ftr_dim, std::string svm_path, int num_groups, int dset)
    : NUMBER_OF_CLUSTERS(num_clust), FEATURE_DIMENSIONALITY(ftr_dim), NUMBER_OF_GROUPS(num_groups), SVM_PATH(svm_path),
      dataset(dset), device_(0), ctx_(nullptr) {
    if (FEATURE_DIMENSIONALITY != 16) die("mofreak_b200: the assignment kernel handles 16-byte MoFREAK descriptors (8 appearance + 8 motion)");
    loadClusters();
}

BagOfWordsRepresentation::~BagOfWordsRepresentation() {
    for (auto* f : training_files) delete f;
    for (auto* f : testing_files) delete f;
    mfk_destroy(ctx_);
}

void BagOfWordsRepresentation::die(const std::string& what) {
    std::cout << what << std::endl;
    exit(1);
}

void BagOfWordsRepresentation::ensureContext() {
    if (ctx_) return;
    // the bag-of-words entry points do not touch the frame store; the smallest legal geometry is enough
    if (mfk_create(&ctx_, device_, 48, 48, MFK_FRAME_GAP + 1, 1) != MFK_OK) die(std::string("mofreak_b200: ") + mfk_last_error(nullptr));
    if (mfk_bow_set_codebook(ctx_, codebook_.data(), (int)(codebook_.size() / 16)) != MFK_OK) die(std::string("mofreak_b200: ") + mfk_last_error(ctx_));
}

// getline consumes a line and the 16 extractions then take the NEXT line's numbers, so the first line of clusters.txt
// is never used and the last iteration parses nothing (the reference leaves `elem` uninitialised there; as observed
// with its own code it repeats the last parsed value, which is what is done here).  :140-173
void BagOfWordsRepresentation::loadClusters() {
    const std::string cluster_path = SVM_PATH + "/clusters.txt";
    std::ifstream cluster_file(cluster_path);
    if ((!cluster_file.is_open()) || cluster_file.bad()) die("could not open clusters file");
    std::string line;
    float elem = 0;
    while (std::getline(cluster_file, line)) {
        for (unsigned col = 0; col < FEATURE_DIMENSIONALITY; ++col) {
            float parsed;
            if (cluster_file >> parsed) elem = parsed;
            codebook_.push_back((unsigned char)elem);
        }
    }
    cluster_file.close();
    if (codebook_.empty()) die("could not open clusters file");
}

bool BagOfWordsRepresentation::buildHistogram(const std::string& file, std::vector<float>& histogram) {
    std::vector<int> frames;
    std::vector<uint8_t> desc;
    parseMofreak(file, FEATURE_DIMENSIONALITY, frames, desc);
    const unsigned K = (unsigned)(codebook_.size() / 16);
    histogram.assign(K, 0.0f);
    if (frames.empty()) return false;
    std::vector<uint32_t> counts(K);
    {
        std::lock_guard<std::mutex> lock(mutex_);
        ensureContext();
        if (mfk_bow_histogram(ctx_, desc.data(), frames.size(), counts.data()) != MFK_OK) die(std::string("mofreak_b200: ") + mfk_last_error(ctx_));
    }
    // the reference accumulates float(+1) per row and divides by the float sum (:115, :125-135); counts below 2^24 are
    // exact in float, so count -> float reproduces it
    float histogram_sum = 0;
    for (unsigned col = 0; col < K; ++col) { histogram[col] = (float)counts[col]; histogram_sum += histogram[col]; }
    for (unsigned col = 0; col < K; ++col) histogram[col] = histogram[col] / histogram_sum;
    return true;
}

void BagOfWordsRepresentation::extractMetadata(std::string filename, int& action, int& group, int& clip_number) {
    action = 0; group = 0; clip_number = 0;          // left uninitialised by the reference for HMDB51 / TRECVID
    if (dataset == UTI2) {
        const std::vector<std::string> parts = split(filename, '_');
        if (parts.size() < 3) die("Unexpected UTI2 file name: " + filename);
        std::stringstream(parts[1]) >> group;
        std::stringstream(parts[2].substr(0, 1)) >> action;
        std::stringstream(parts[0]) >> clip_number;
    } else if (dataset == UCF101 || dataset == KTH) {
        const std::vector<std::string> parts = split(filename, '_');
        if (parts.size() < 4) die("Unexpected file name: " + filename);
        const std::string parsed_action = parts[1];
        if (actions.find(parsed_action) == actions.end()) { const int next = (int)actions.size(); actions[parsed_action] = next; }
        action = actions[parsed_action];
        if (dataset == KTH) std::stringstream(parts[0].substr(parts[0].length() - 2, 2)) >> group;
        else std::stringstream(parts[2].substr(1, 2)) >> group;
        group--;
        std::stringstream(parts[3].substr(2, 2)) >> clip_number;
    }
}

void BagOfWordsRepresentation::intializeBOWMemory(std::string svm_path) {
    for (unsigned group = 0; group < NUMBER_OF_GROUPS; ++group) {
        std::stringstream training_filepath, testing_filepath;
        training_filepath << svm_path << "/" << group + 1 << ".train";
        testing_filepath << svm_path << "/" << group + 1 << ".test";
        training_files.push_back(new std::ofstream(training_filepath.str()));
        testing_files.push_back(new std::ofstream(testing_filepath.str()));
        bow_training_crossvalidation_sets.push_back(std::vector<std::string>());
        bow_testing_crossvalidation_sets.push_back(std::vector<std::string>());
    }
}

void BagOfWordsRepresentation::convertFileToBOWFeature(std::string file) {
    const std::string file_name_str = baseName(file);
    int action, group, video_number;
    std::vector<float> bow_feature;
    const bool success = buildHistogram(file, bow_feature);
    if (!success) die("Bag-of-words feature construction was unsuccessful.  Investigate.");
    std::lock_guard<std::mutex> lock(mutex_);
    extractMetadata(file_name_str, action, group, video_number);      // touches the shared action map
    const std::string current_line = libsvmLine(action + 1, bow_feature.data(), (unsigned)bow_feature.size());
    // every training file gets the line except the one whose group is held out (:1016-1030)
    for (unsigned g = 0; g < NUMBER_OF_GROUPS && g < bow_training_crossvalidation_sets.size(); ++g) {
        if ((int)g == group) bow_testing_crossvalidation_sets[g].push_back(current_line);
        else bow_training_crossvalidation_sets[g].push_back(current_line);
    }
}

void BagOfWordsRepresentation::writeBOWFeaturesToFiles() {
    if (bow_training_crossvalidation_sets.size() != NUMBER_OF_GROUPS || bow_testing_crossvalidation_sets.size() != NUMBER_OF_GROUPS)
        die("Incorrect number of training or testing file lines.  Check mapping from bow feature to test/train files");
    for (unsigned i = 0; i < NUMBER_OF_GROUPS; ++i) {
        std::cout << "number of training features: " << bow_training_crossvalidation_sets[i].size() << std::endl;
        for (const std::string& l : bow_training_crossvalidation_sets[i]) *training_files[i] << l << std::endl;
        std::cout << "number of testing features: " << bow_testing_crossvalidation_sets[i].size() << std::endl;
        for (const std::string& l : bow_testing_crossvalidation_sets[i]) *testing_files[i] << l << std::endl;
    }
    std::cout << "Finished writing to cross-validation files." << std::endl;
    for (unsigned i = 0; i < NUMBER_OF_GROUPS; ++i) {
        training_files[i]->close(); testing_files[i]->close();
        delete training_files[i]; delete testing_files[i];
    }
    std::cout << "Closed all cross-validation files. " << std::endl;
    training_files.clear();
    testing_files.clear();
}

// TRECVID sliding window (:176-378): one histogram per key frame, a libsvm line for every window of `alpha`
// consecutive frame histograms (sum, then L1 normalisation); the rows of the file's LAST frame never form a
// histogram in the reference (the loop only closes a frame when the next one starts), which is kept.  All rows are
// assigned in one device call; the per-frame bookkeeping is a running sum on the host.
void BagOfWordsRepresentation::computeSlidingBagOfWords(std::string& file, int alpha, int label, std::ofstream& out) {
    std::vector<int> frames;
    std::vector<uint8_t> desc;
    parseMofreak(file, FEATURE_DIMENSIONALITY, frames, desc);
    if (frames.empty()) return;
    const unsigned K = (unsigned)(codebook_.size() / 16);
    std::vector<int32_t> idx(frames.size());
    {
        std::lock_guard<std::mutex> lock(mutex_);
        ensureContext();
        if (mfk_bow_assign(ctx_, desc.data(), frames.size(), idx.data()) != MFK_OK) die(std::string("mofreak_b200: ") + mfk_last_error(ctx_));
    }
    std::list<std::vector<float> > histograms_per_frame;
    bool over_alpha_frames = false;
    auto emit = [&]() {
        std::vector<float> histogram(K, 0.0f);
        for (const auto& h : histograms_per_frame)
            for (unsigned col = 0; col < K; ++col) histogram[col] += h[col];
        float normalizer = 0.0f;
        for (unsigned col = 0; col < K; ++col) normalizer += histogram[col];
        for (unsigned col = 0; col < K; ++col) histogram[col] = histogram[col] / normalizer;
        out << libsvmLine(label, histogram.data(), K) << std::endl;
    };
    size_t begin = 0;
    for (size_t i = 1; i <= frames.size(); ++i) {
        if (i < frames.size() && frames[i] == frames[begin]) continue;
        if (i == frames.size()) break;                      // the last frame's rows are never closed
        std::vector<float> h(K, 0.0f);
        for (size_t r = begin; r < i; ++r) h[idx[r]] += 1.0f;
        histograms_per_frame.push_back(h);
        if ((int)histograms_per_frame.size() == alpha) {
            over_alpha_frames = true;
            emit();
            histograms_per_frame.pop_front();
        }
        begin = i;
    }
    if (!over_alpha_frames) emit();
}
