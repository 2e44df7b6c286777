#include "Clustering.h"

#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <fstream>
#include <iostream>

namespace {

// rand() of the reference's runtime (MSVC: 15-bit LCG), so that `rand() % row0` keeps its range and its bias for more
// than 32768 rows (Clustering.cpp:17)
struct Lcg15 {
    uint32_t state;
    explicit Lcg15(uint32_t seed) : state(seed) {}
    int next() { state = state * 214013u + 2531011u; return (int)((state >> 16) & 0x7fffu); }
};

}  // namespace

Clustering::Clustering(int dim, int num_clust, int num_pts, int num_classes, std::vector<int> poss_classes, std::string svm_path)
    : seed_(0), shuffles_(0), data_rows_(0), centers_((size_t)num_clust * dim, 0), center_row(0), possible_classes(poss_classes),
      appearance_descriptor_size(0), motion_descriptor_size(0), appearance_is_binary(true), motion_is_binary(true),
      SVM_PATH(svm_path), DIMENSIONALITY(dim), NUMBER_OF_CLUSTERS(num_clust), NUMBER_OF_POINTS_TO_SAMPLE(num_pts),
      NUMBER_OF_CLASSES(num_classes) {}

void Clustering::setMotionDescriptor(unsigned int size, bool binary) { motion_descriptor_size = (int)size; motion_is_binary = binary; }
void Clustering::setAppearanceDescriptor(unsigned int size, bool binary) { appearance_descriptor_size = (int)size; appearance_is_binary = binary; }

// Clustering.cpp:6-25: "Fisher-Yates" whose partner index is rand() % row0, i.e. never row0 itself
void Clustering::shuffleRows(std::vector<uint8_t>& mx, int cols) {
    Lcg15 rng(seed_ ? seed_ + 0x9E3779B9u * shuffles_ : (uint32_t)time(NULL));
    shuffles_++;
    std::vector<uint8_t> tmp((size_t)cols);
    int row0 = data_rows_ - 1;
    while (row0 > 0) {
        const int row1 = rng.next() % row0;
        memcpy(tmp.data(), &mx[(size_t)row1 * cols], (size_t)cols);
        memcpy(&mx[(size_t)row1 * cols], &mx[(size_t)row0 * cols], (size_t)cols);
        memcpy(&mx[(size_t)row0 * cols], tmp.data(), (size_t)cols);
        row0--;
    }
}

void Clustering::buildDataFromMoFREAK(std::deque<MoFREAKFeature> mofreak_ftrs, bool sample_pts, bool, bool fix_class, int fixed_class) {
    const int cols = DIMENSIONALITY + 3;
    const int num_ftrs = (int)mofreak_ftrs.size();
    std::cout << "allocating data_pts in builddatafrommofreaK: " << num_ftrs << std::endl;
    data_pts_.assign((size_t)num_ftrs * cols, 0);
    data_rows_ = num_ftrs;
    for (int row = 0; row < num_ftrs; ++row) {
        const MoFREAKFeature& ftr = mofreak_ftrs[row];
        uint8_t* d = &data_pts_[(size_t)row * cols];
        d[0] = fix_class ? (uint8_t)fixed_class : (uint8_t)ftr.action;
        d[1] = (uint8_t)ftr.person;
        d[2] = (uint8_t)ftr.video_number;
        for (int col = 0; col < appearance_descriptor_size; ++col) d[col + 3] = (uint8_t)ftr.appearance[col];
        for (int col = 0; col < motion_descriptor_size; ++col) d[col + appearance_descriptor_size + 3] = (uint8_t)ftr.motion[col];
    }
    if (sample_pts) {
        shuffleRows(data_pts_, cols); shuffleRows(data_pts_, cols); shuffleRows(data_pts_, cols);
        if (data_rows_ > NUMBER_OF_POINTS_TO_SAMPLE) { data_rows_ = NUMBER_OF_POINTS_TO_SAMPLE; data_pts_.resize((size_t)data_rows_ * cols); }
    }
}

void Clustering::randomClusters(bool only_one_class) {
    std::cout << "Random clustering" << std::endl;
    const int cols = DIMENSIONALITY + 3;
    shuffleRows(data_pts_, cols); shuffleRows(data_pts_, cols); shuffleRows(data_pts_, cols);
    // an even number of points from each class keeps the classes balanced
    const int CLUSTERS_PER_CLASS = NUMBER_OF_CLUSTERS / NUMBER_OF_CLASSES;
    std::cout << "Sampling " << CLUSTERS_PER_CLASS << std::endl;
    std::cout << "Current center row:" << center_row << std::endl;
    for (size_t ci = 0; ci < possible_classes.size(); ++ci) {
        int sampled_from_this_class = 0;
        for (int row = 0; row < data_rows_; ++row) {
            const uint8_t* d = &data_pts_[(size_t)row * cols];
            if ((int)d[0] == possible_classes[ci] || only_one_class) {
                if (center_row >= NUMBER_OF_CLUSTERS) {            // cv::Mat::at would throw here; the reference prints and exits (:190-195)
                    std::cout << "Error: more cluster centres than NUMBER_OF_CLUSTERS" << std::endl;
                    exit(1);
                }
                memcpy(&centers_[(size_t)center_row * DIMENSIONALITY], d + 3, (size_t)DIMENSIONALITY);
                center_row++;
                sampled_from_this_class++;
                if (sampled_from_this_class == CLUSTERS_PER_CLASS) break;
            }
        }
        if (only_one_class) break;
    }
}

void Clustering::writeClusters(bool append) {
    std::cout << "writing" << std::endl;
    std::ofstream output_file;
    if (append) output_file.open(SVM_PATH + "/clusters.txt", std::ios_base::app);
    else output_file.open(SVM_PATH + "/clusters.txt");
    std::string text;
    char buf[8];
    for (int i = 0; i < NUMBER_OF_CLUSTERS; ++i) {
        for (int j = 0; j < DIMENSIONALITY; ++j) {             // (float)uchar through ostream: an integer, then one space
            const int n = snprintf(buf, sizeof buf, "%d ", (int)centers_[(size_t)i * DIMENSIONALITY + j]);
            text.append(buf, n);
        }
        text += '\n';
    }
    output_file.write(text.data(), (std::streamsize)text.size());
    output_file.close();
}
