// TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// Restatement of the bag-of-words assignment of reference
// src/MoFREAK/BagOfWordsRepresentation.cpp:22-37 (bruteForceMatch), :39-72 (hammingDistance),
// :74-138 (buildHistogram) and :140-173 (loadClusters).
#include "oracle.h"

#include <limits.h>
#include <sstream>

namespace orc {

// the reference masks with an `unsigned int` bit that walks all 32 positions (:63)
unsigned hamming_u8(unsigned char a, unsigned char b) {
    unsigned hamming_distance = 0;
    const unsigned xor_result = a ^ b;
    for (unsigned bit = 1; bit != 0; bit <<= 1)
        if ((xor_result & bit) != 0) hamming_distance++;
    return hamming_distance;
}

// strict '<' keeps the lowest index among equal distances (:30)
int brute_force_match(const uint8_t desc[16], const uint8_t* codebook, int K) {
    unsigned shortest_distance = INT_MAX;
    int shortest_index = -1;
    for (int i = 0; i < K; i++) {
        unsigned dist = 0;
        for (int col = 0; col < 16; col++) dist += hamming_u8(desc[col], codebook[(size_t)i * 16 + col]);
        if (dist < shortest_distance) { shortest_distance = dist; shortest_index = i; }
    }
    return shortest_index;
}

// getline consumes a line, then 16 operator>> reads take the NEXT line's numbers: the first line of
// clusters.txt is never used and the final iteration parses nothing.  The reference leaves `elem`
// uninitialised there (formally undefined; in practice the last parsed value is repeated, SURVEY
// App. B #7) — defined here as "repeat the last parsed value".
std::vector<uint8_t> load_clusters_text(const std::string& text) {
    std::vector<uint8_t> out;
    std::istringstream cluster_file(text);
    std::string line;
    float elem = 0;
    while (std::getline(cluster_file, line)) {
        for (unsigned col = 0; col < 16; ++col) {
            float parsed;
            if (cluster_file >> parsed) elem = parsed;
            out.push_back((unsigned char)elem);
        }
    }
    return out;
}

bool build_histogram(const std::string& mofreak_text, const uint8_t* codebook, int K, std::vector<float>& hist) {
    bool success = false;
    hist.assign(K, 0.0f);
    std::istringstream input_file(mofreak_text);
    std::string line;
    while (std::getline(input_file, line)) {
        std::istringstream iss(line);
        double discard;
        for (unsigned i = 0; i < 6; ++i) iss >> discard;
        uint8_t feature_vector[16];
        float elem = 0;
        for (unsigned i = 0; i < 16; ++i) {
            iss >> elem;
            feature_vector[i] = (unsigned char)elem;
        }
        const int best_match = brute_force_match(feature_vector, codebook, K);
        hist[best_match] = hist[best_match] + 1;
        success = true;
    }
    if (!success) return false;
    float histogram_sum = 0;
    for (int col = 0; col < K; ++col) histogram_sum += hist[col];
    for (int col = 0; col < K; ++col) hist[col] = hist[col] / histogram_sum;
    return true;
}

}  // namespace orc
