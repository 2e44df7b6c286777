// TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// Restatement of BagOfWordsRepresentation::computeSlidingBagOfWords, reference
// src/MoFREAK/BagOfWordsRepresentation.cpp:176-378 (TRECVID: one bag-of-words line per window of `alpha` key frames),
// statement by statement on the text of a `.mofreak` file.  What the walk does, as written:
//   * the first line opens the first frame (:188-212); every further line either joins the current frame's feature list or
//     CLOSES the current frame (:242-262): histogram of its features, pushed onto a list; when the list holds `alpha`
//     histograms their sum is L1-normalised and written as a libsvm line, then the oldest histogram is dropped (:268-312);
//   * the LAST frame of the file is never closed (the loop ends with its features still in the list, :214-331);
//   * if no window ever filled, one line is written from whatever histograms exist (:334-377).
// The match (:255-258) goes through cv::BFMatcher(NORM_HAMMING)::match with a CV_32FC1 query against the CV_8U codewords
// added by loadClusters (:170-172) and takes matches[0].imgIdx.  OpenCV's matcher is not vendored; mixing the two
// element types is rejected by its type assertion (recalled from OpenCV 2.4 matchers.cpp, not verifiable offline), so the
// restatement DEFINES the match as what the code evidently intends and what the reference's own bruteForceMatch
// (:22-37) computes for buildHistogram: the Hamming arg-min over the (unsigned char) casts of the 16 values, lowest
// codeword index on ties (BFMatcher keeps the first best match, and imgIdx is the codeword's position in the list).
// PARITY UNPINNED for that one call; everything around it follows the reference's statements.
#include "oracle.h"

#include <list>
#include <sstream>

namespace orc {

namespace {

std::string sliding_line(int label, const std::vector<float>& histogram) {      // :296-309, :361-374
    std::stringstream ss;
    ss << label << " ";
    for (size_t col = 0; col < histogram.size(); ++col) ss << (col + 1) << ":" << histogram[col] << " ";
    return ss.str();
}

std::vector<float> summed_and_normalised(const std::list<std::vector<float> >& histograms_per_frame, int K) {
    std::vector<float> histogram(K, 0.0f);
    for (std::list<std::vector<float> >::const_iterator it = histograms_per_frame.begin(); it != histograms_per_frame.end(); ++it)
        for (int col = 0; col < K; ++col) histogram[col] += (*it)[col];
    float normalizer = 0.0;
    for (int col = 0; col < K; ++col) normalizer += histogram[col];
    for (int col = 0; col < K; ++col) histogram[col] = histogram[col] / normalizer;
    return histogram;
}

}  // namespace

std::string sliding_bag_of_words(const std::string& mofreak_text, const uint8_t* codebook, int K, int alpha, int label) {
    std::string out;
    bool over_alpha_frames = false;
    std::list<std::vector<float> > histograms_per_frame;
    std::vector<std::vector<float> > feature_list;          // for just the current frame

    std::istringstream input_file(mofreak_text);
    std::string line;
    int current_frame = 0;
    std::getline(input_file, line);
    {
        std::istringstream init(line);
        double discard;
        init >> discard; init >> discard;
        init >> current_frame;
        init >> discard; init >> discard; init >> discard;
        std::vector<float> ftr_vector(16, 0.0f);
        float elem = 0;
        for (unsigned i = 0; i < 16; ++i) { init >> elem; ftr_vector[i] = elem; }
        feature_list.push_back(ftr_vector);
    }
    while (std::getline(input_file, line)) {
        std::istringstream iss(line);
        double discard;
        iss >> discard; iss >> discard;
        int frame = 0;
        iss >> frame;
        iss >> discard; iss >> discard; iss >> discard;
        std::vector<float> feature_vector(16, 0.0f);
        float elem = 0;
        if (frame == current_frame) {
            for (unsigned i = 0; i < 16; ++i) { iss >> elem; feature_vector[i] = elem; }
            feature_list.push_back(feature_vector);
        } else {
            // 1: the histogram of the frame that just ended
            std::vector<float> new_histogram(K, 0.0f);
            for (size_t f = 0; f < feature_list.size(); ++f) {
                uint8_t bytes[16];
                for (int i = 0; i < 16; i++) bytes[i] = (unsigned char)feature_list[f][i];
                const int img_idx = brute_force_match(bytes, codebook, K);
                new_histogram[img_idx] = new_histogram[img_idx] + 1;
            }
            // 2: onto the list; at capacity: emit, then drop the oldest
            histograms_per_frame.push_back(new_histogram);
            if ((int)histograms_per_frame.size() == alpha) {
                over_alpha_frames = true;
                // the reference sums first, pops, then normalises and writes (:270-311): same result
                const std::vector<float> histogram = summed_and_normalised(histograms_per_frame, K);
                histograms_per_frame.pop_front();
                out += sliding_line(label, histogram);
                out += "\n";
            }
            feature_list.clear();
            for (int i = 0; i < 16; ++i) { iss >> elem; feature_vector[i] = elem; }
            feature_list.push_back(feature_vector);
            current_frame = frame;
        }
    }
    if (!over_alpha_frames) {
        out += sliding_line(label, summed_and_normalised(histograms_per_frame, K));
        out += "\n";
    }
    return out;
}

}  // namespace orc
