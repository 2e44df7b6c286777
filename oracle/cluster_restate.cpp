// TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// Restatement of the codebook selection of reference src/MoFREAK/Clustering.cpp:
//   shuffleCVMat :6-25      for row0 = rows-1 .. 1: swap row0 with row (rand() % row0)  - the partner is never row0 itself
//   randomClusters :165-213 three shuffles, then for every possible class (or once, if only_one_class) walk the rows from
//                           the top and copy the descriptor part (columns 3..) of each row of that class into the next free
//                           centre until NUMBER_OF_CLUSTERS / NUMBER_OF_CLASSES were taken
//   writeClusters :233-257  every centre as "(float)byte " tokens, one line per centre
// The reference seeds rand() with time(NULL) inside every shuffle, so its output is not reproducible; the random numbers
// are therefore INJECTED here: `rands` is the sequence rand() would return, consumed in call order (three shuffles of
// rows-1 draws each).  The product's Clustering class (mofreak_b200/host) is compared against this with the sequence its
// seeded generator produces.
#include "oracle.h"

#include <sstream>

namespace orc {

// data: rows x (dim + 3) bytes (action, person, video, then the descriptor).  Returns the number of draws consumed.
size_t shuffle_rows(uint8_t* data, int rows, int cols, const int* rands) {
    size_t used = 0;
    int row0 = rows - 1;
    std::vector<uint8_t> mt(cols);
    while (row0 > 0) {
        const int row1 = rands[used++] % row0;
        uint8_t* m1 = data + (size_t)row1 * cols;
        uint8_t* m0 = data + (size_t)row0 * cols;
        mt.assign(m1, m1 + cols);                     // mt = m1.clone()
        for (int c = 0; c < cols; c++) m1[c] = m0[c]; // mx.row(row0).copyTo(m1)
        for (int c = 0; c < cols; c++) m0[c] = mt[c]; // mt.copyTo(mx.row(row0))
        row0--;
    }
    return used;
}

// centers: num_clusters x dim bytes, filled from *center_row on (the reference's member of the same name).
// Returns the draws consumed; *center_row is advanced.
size_t random_clusters(uint8_t* data, int rows, int dim, int num_clusters, int num_classes, const int* classes, int n_classes,
                       bool only_one_class, const int* rands, uint8_t* centers, int* center_row) {
    const int cols = dim + 3;
    size_t used = 0;
    used += shuffle_rows(data, rows, cols, rands + used);
    used += shuffle_rows(data, rows, cols, rands + used);
    used += shuffle_rows(data, rows, cols, rands + used);
    const int CLUSTERS_PER_CLASS = num_clusters / num_classes;
    for (int ci = 0; ci < n_classes; ci++) {
        int sampled_from_this_class = 0;
        for (int row = 0; row < rows; ++row) {
            if ((int)data[(size_t)row * cols] == classes[ci] || only_one_class) {
                for (int col = 3; col < cols; ++col) centers[(size_t)(*center_row) * dim + (col - 3)] = data[(size_t)row * cols + col];
                (*center_row)++;
                sampled_from_this_class++;
                if (sampled_from_this_class == CLUSTERS_PER_CLASS) break;
            }
        }
        if (only_one_class) break;
    }
    return used;
}

std::string clusters_text(const uint8_t* centers, int num_clusters, int dim) {
    std::ostringstream output_file;
    for (int i = 0; i < num_clusters; ++i) {
        for (int j = 0; j < dim; ++j) output_file << (float)centers[(size_t)i * dim + j] << " ";
        output_file << std::endl;
    }
    return output_file.str();
}

}  // namespace orc
