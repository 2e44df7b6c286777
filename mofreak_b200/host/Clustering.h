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
    Clustering(int dim, int num_clust, int num_pts, int num_classes, std::vector<int> poss_classes, std::string svm_path);
    // rows = 3 metadata bytes (action, person, video_number) + appearance bytes + motion bytes (Clustering.cpp:57-116)
    void buildDataFromMoFREAK(std::deque<MoFREAKFeature> mofreak_ftrs, bool sample_pts, bool img_diff = false, bool fix_class = false,
                              int fixed_class = 1);
    void randomClusters(bool only_one_class = false);     // Clustering.cpp:165-213
    void writeClusters(bool append = false);              // Clustering.cpp:233-257
    void setMotionDescriptor(unsigned int size, bool binary);
    void setAppearanceDescriptor(unsigned int size, bool binary);

    // --- not in the reference ---
    // 0 (default) = the reference's behaviour, srand(time(NULL)) at every shuffle; any other value makes the selection reproducible
    void setSeed(uint32_t seed) { seed_ = seed; }
    const std::vector<uint8_t>& centers() const { return centers_; }      // NUMBER_OF_CLUSTERS x DIMENSIONALITY
    int centersFilled() const { return center_row; }

private:
    void shuffleRows(std::vector<uint8_t>& mx, int cols);

    const int DIMENSIONALITY;
    const int NUMBER_OF_CLUSTERS;
    const int NUMBER_OF_POINTS_TO_SAMPLE;
    const int NUMBER_OF_CLASSES;
    const std::string SVM_PATH;

    std::vector<uint8_t> centers_;
    std::vector<uint8_t> data_pts_;      // rows x (DIMENSIONALITY + 3)
    int data_rows_;
    std::vector<int> possible_classes;
    int motion_descriptor_size;
    int appearance_descriptor_size;
    bool motion_is_binary;
    bool appearance_is_binary;
    int center_row;
    uint32_t seed_;
    uint32_t shuffles_;                  // distinguishes the shuffles of one seeded run
};
