// TEST INFRASTRUCTURE ONLY — the CPU oracle for the MoFREAK extraction hot path.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
// anything built from oracle/.  The product (mofreak_b200/) never includes, links or calls it and has
// no CPU fallback: it fails loudly when its CUDA library is missing.
//
// What is in here
//   brisk_port.cpp       our restatement of the vendored BRISK scale-space detector
//                        (reference src/MoFREAK/brisk.cpp:544-704, 838-1725, 1840-2065 and the
//                        OAST-9/16 / AGAST-5/8 segment tests of oast9_16*.cc, agast5_8*.cc), written in
//                        the phase-structured, order-independent form the CUDA kernels use.  PINNED:
//                        tests compare it against the reference's own object code (oracle/_ref).
//   freak_restate.cpp    cv::FREAK (OpenCV 2.4.2 features2d/src/freak.cpp, NOT vendored by the
//                        reference; call site MoFREAKUtilities.cpp:427-428).  PARITY UNPINNED: no
//                        FREAK implementation exists in this container and the reference holds no
//                        golden vectors; restated from the published algorithm.
//   mip_restate.cpp      MIP motion bytes (MoFREAKUtilities.cpp:46-99, 288-325) and the cv::resize
//                        INTER_LINEAR 8-bit model (OpenCV, un-vendored; pinned against cv2.resize
//                        of OpenCV 4.13 by tests/test_oracle_resize.py).
//   extract_restate.cpp  the computeMoFREAKFromFile loop body on raw gray frames
//                        (MoFREAKUtilities.cpp:374-498) and the .mofreak writer (:691-719).
//   bow_restate.cpp      bruteForceMatch / hammingDistance / buildHistogram / loadClusters
//                        (BagOfWordsRepresentation.cpp:22-173).
//   sliding_restate.cpp  computeSlidingBagOfWords (BagOfWordsRepresentation.cpp:176-378), SURVEY 8(f) N2.
//   cluster_restate.cpp  shuffleCVMat / randomClusters / writeClusters (Clustering.cpp:6-25, 165-213, 233-257) with the
//                        random numbers injected, SURVEY 8(f) N3.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <string>
#include <vector>

namespace orc {

struct Keypoint {      // cv::KeyPoint fields the path uses (brisk.cpp:677-678, 699)
    float x, y, size, response;
    int layer;
};

struct Row {           // one .mofreak row, MoFREAKUtilities.h:23-53 payload; 32 bytes
    float x, y;
    int32_t frame;
    float scale;
    uint8_t appearance[8];
    uint8_t motion[8];
};

// ---- detector ------------------------------------------------------------------------------
typedef void (*DetectFn)(const uint8_t* img, int w, int h, std::vector<Keypoint>& out);

// Our restatement (threshold 30, 3 octaves fixed as at MoFREAKUtilities.cpp:420-423).
void brisk_port_detect(const uint8_t* img, int w, int h, std::vector<Keypoint>& out);
// jacobi_reverse = resolve tie candidates in REVERSE raster order with deferred status updates,
// i.e. the least favourable schedule a parallel fixed point could see; must give identical output.
void brisk_port_detect_ex(const uint8_t* img, int w, int h, int thresh, bool jacobi_reverse,
                          std::vector<Keypoint>& out);
#ifdef MFK_ORACLE_WITH_REF
// The reference's own object code (oracle/_ref build only).
void brisk_ref_detect(const uint8_t* img, int w, int h, std::vector<Keypoint>& out);
#endif

// Stage helpers of the port (pure functions; used for stage-by-stage diffs).
void port_halfsample(const uint8_t* src, int w, int h, uint8_t* dst);       // brisk.cpp:1840-1972
void port_twothirdsample(const uint8_t* src, int w, int h, uint8_t* dst);   // brisk.cpp:1974-2065
int  port_oast_score(const uint8_t* img, int w, int x, int y);              // oast9_16_nms.cc:42, b0 = 0
int  port_agast58_score(const uint8_t* img, int w, int x, int y);           // agast5_8_nms.cc:42, b0 = 0
void port_score_map(const uint8_t* img, int w, int h, uint8_t* S);          // dense S_full, 0 in the 3-px border

// ---- FREAK ----------------------------------------------------------------------------------
struct FreakPattern;                       // 64 x 256 x 43 look-up + pairs (freak.cpp buildPattern)
const FreakPattern& freak_pattern();       // built once ("hoisted")
FreakPattern* freak_build_pattern();       // rebuilt per call ("as written": MoFREAKUtilities.cpp:427)
void freak_free_pattern(FreakPattern*);
int  freak_pattern_size(const FreakPattern&, int scale_idx);
int  freak_scale_index(float kp_size);     // max(int(log(size/7)*sizeCst+0.5),0) clamped to 63
// Keeps OpenCV's semantics: keypoints failing the border test are erased (order preserved);
// desc receives 64 bytes per surviving keypoint.  theta_idx (optional) receives the orientation bin.
void freak_compute(const FreakPattern& P, const uint8_t* img, int w, int h, std::vector<Keypoint>& kps,
                   std::vector<uint8_t>& desc, std::vector<int>* theta_idx);
void integral_u8(const uint8_t* img, int w, int h, int32_t* sum /*(h+1)*(w+1)*/);

// ---- MIP -------------------------------------------------------------------------------------
void resize_linear_u8_19(const uint8_t* src, int src_stride, int s, uint8_t dst[361]);  // s x s ROI -> 19x19
unsigned mip_byte(const uint8_t cur19[361], const uint8_t prev19[361], int x, int y);    // MoFREAKUtilities.cpp:46-99
void mip_extract(const uint8_t* cur, const uint8_t* prev, int w, int h, float size, int x, int y,
                 uint8_t motion[8]);                                                     // :288-325

// ---- extraction loop + writer ---------------------------------------------------------------------
struct ExtractOptions {
    bool as_written;      // true: rebuild the FREAK table per frame and run the discarded
                          // sufficientMotion work (MoFREAKUtilities.cpp:427, 443); results identical
    DetectFn detect;      // brisk_port_detect or brisk_ref_detect
};
// frames: n_frames gray u8 images, each w*h, contiguous.  Appends rows in reference order.
void extract_clip(const uint8_t* frames, int w, int h, int n_frames, const ExtractOptions& opt,
                  std::vector<Row>& rows);
std::string format_mofreak(const Row* rows, size_t n);        // MoFREAKUtilities.cpp:691-719, byte-identical

// ---- bag of words ------------------------------------------------------------------------------
unsigned hamming_u8(unsigned char a, unsigned char b);                          // BagOfWordsRepresentation.cpp:53-72
int  brute_force_match(const uint8_t desc[16], const uint8_t* codebook, int K); // :22-51
// text of clusters.txt -> K x 16 bytes with the reference's off-by-one parse (:140-173).
std::vector<uint8_t> load_clusters_text(const std::string& text);
// .mofreak text -> normalised histogram (:74-138); returns false for an empty file.
bool build_histogram(const std::string& mofreak_text, const uint8_t* codebook, int K, std::vector<float>& hist);

// ---- "next" rows of SURVEY 8(f) -------------------------------------------------------------------
// TRECVID sliding-window bag of words (:176-378): text of the libsvm lines written for one `.mofreak` file
std::string sliding_bag_of_words(const std::string& mofreak_text, const uint8_t* codebook, int K, int alpha, int label);
// codebook selection (Clustering.cpp) with injected random numbers; see cluster_restate.cpp
size_t shuffle_rows(uint8_t* data, int rows, int cols, const int* rands);
size_t random_clusters(uint8_t* data, int rows, int dim, int num_clusters, int num_classes, const int* classes, int n_classes,
                       bool only_one_class, const int* rands, uint8_t* centers, int* center_row);
std::string clusters_text(const uint8_t* centers, int num_clusters, int dim);

}  // namespace orc
