// Codebook selection with the reference's call surface (src/MoFREAK/Clustering.h:15-53; driver main.cpp:160-213).
// the reference draws the codewords at random (rand() seeded with time(NULL) on every shuffle,
// Clustering.cpp:6-25), so its clusters.txt is not reproducible; here the generator is explicit and seedable, everything
// else - the three shuffles, the per-class quota, the row layout, the clusters.txt format - follows the reference.
// Host-side bookkeeping: no GPU work (k-means on Hamming data, Clustering.cpp:215-231, is dead code in the reference).
#pragma once
#include <stdint.h>

#include <deque>
#include <string>
#include <vector>

#include "MoFREAKUtilities.h"

class Clustering {
public:
    // ---- extensions (not in the reference) ----
    // 0 (default) = the reference's behaviour, srand(time(NULL)) at every shuffle; any other value makes the selection reproducible
    void setSeed(uint32_t seed) { seed_ = seed; }
    const std::vector<uint8_t>& centers() const { return centers_; }      // num_clust x dim bytes, row-major
    int centersFilled() const { return center_row; }

    // ---- the reference's surface (Clustering.h:15-27), same names, argument order and meaning ----
    Clustering(int dim, int num_clust, int num_pts, int num_classes, std::vector<int> poss_classes, std::string svm_path);
    void setAppearanceDescriptor(unsigned int size, bool binary);
    void setMotionDescriptor(unsigned int size, bool binary);
    // one data row per feature: action, person, video_number, then the appearance and the motion bytes (Clustering.cpp:57-116)
    void buildDataFromMoFREAK(std::deque<MoFREAKFeature> mofreak_ftrs, bool sample_pts, bool img_diff = false, bool fix_class = false,
                              int fixed_class = 1);
    void randomClusters(bool only_one_class = false);     // Clustering.cpp:165-213
    void writeClusters(bool append = false);              // Clustering.cpp:233-257

private:
    void shuffleRows(std::vector<uint8_t>& mx, int cols);

    uint32_t seed_;
    uint32_t shuffles_;                  // distinguishes the shuffles of one seeded run
    std::vector<uint8_t> data_pts_;      // data_rows_ x (DIMENSIONALITY + 3) bytes
    int data_rows_;
    std::vector<uint8_t> centers_;
    int center_row;                      // next free codeword; classes append one after the other
    std::vector<int> possible_classes;
    int appearance_descriptor_size, motion_descriptor_size;
    bool appearance_is_binary, motion_is_binary;
    const std::string SVM_PATH;
    const int DIMENSIONALITY, NUMBER_OF_CLUSTERS, NUMBER_OF_POINTS_TO_SAMPLE, NUMBER_OF_CLASSES;
};
