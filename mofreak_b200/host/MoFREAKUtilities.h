// Host-side mirror of the reference's MoFREAKUtilities (src/MoFREAK/MoFREAKUtilities.h:23-104): same class name,
// same public method names, argument meaning and error behaviour for the extraction path, implemented on the
// C ABI of include/mofreak_b200.h (CUDA, sm_100a).  No OpenCV, no Boost, no CPU fallback.
//
// Differences a caller can see, all outside the arithmetic contract:
//   * computeMoFREAKFromFile reads gray frames from a raw ".y8" container (GrayVideo below) instead of decoding a
//     video: decode + cvtColor(BGR2GRAY) (reference MoFREAKUtilities.cpp:380-410) sit before the hot path;
//     mofreak_b200/tools/video_to_y8.py converts any video OpenCV can open.
//   * computeMoFREAKFromFrames is the same call for frames already in memory.
//   * buildMoFREAKFeaturesFromMoSIFT (legacy MoSIFT path, never reached from the reference's main()) is absent.
#ifndef MOFREAK_B200_MOFREAKUTILITIES_H
#define MOFREAK_B200_MOFREAKUTILITIES_H

#include <stdint.h>

#include <deque>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/mofreak_b200.h"

#define MOTION_BYTES 8
#define APPEARANCE_BYTES 8

struct MoFREAKFeature {                       // reference MoFREAKUtilities.h:23-53
    MoFREAKFeature(int motion_bytes, int appearance_bytes)
        : using_image_difference(false), x(0), y(0), scale(0), motion_x(0), motion_y(0), frame_number(0),
          motion(motion_bytes, 0u), appearance(appearance_bytes, 0u), action(0), video_number(0), person(0) {}
    bool using_image_difference;
    float x, y, scale, motion_x, motion_y;
    int frame_number;
    std::vector<unsigned int> motion;
    std::vector<unsigned int> appearance;
    int action, video_number, person;
};

// Raw gray video: 8-byte magic "MFKGRAY1", int32 width, height, frame count (little endian), then dense frames.
struct GrayVideo {
    int width = 0, height = 0, frames = 0;
    std::vector<uint8_t> data;
    bool load(const std::string& path, std::string* error);
    static bool save(const std::string& path, const uint8_t* frames, int width, int height, int n, std::string* error);
};

class MoFREAKUtilities {
public:
    enum datasets { KTH, TRECVID, HOLLYWOOD, UTI1, UTI2, HMDB51, UCF101 };

    explicit MoFREAKUtilities(int dset);
    ~MoFREAKUtilities();
    MoFREAKUtilities(const MoFREAKUtilities&) = delete;
    MoFREAKUtilities& operator=(const MoFREAKUtilities&) = delete;

    // reference call surface (MoFREAKUtilities.h:58-73)
    void readMoFREAKFeatures(std::string filename, int num_to_sample = 0);
    std::deque<MoFREAKFeature> getMoFREAKFeatures();
    void clearFeatures();
    void writeMoFREAKFeaturesToFile(std::string output_file);
    void computeMoFREAKFromFile(std::string video_filename, std::string mofreak_filename, bool clear_features_after_computation);
    void setAllFeaturesToLabel(int label);
    void setCurrentAction(std::string folder_name);
    int current_action;
    std::unordered_map<std::string, int> actions;

    static const int NUMBER_OF_BYTES_FOR_APPEARANCE = APPEARANCE_BYTES;
    static const int NUMBER_OF_BYTES_FOR_MOTION = MOTION_BYTES;

    // same computation on frames already in host memory (gray u8, row pitch = width)
    void computeMoFREAKFromFrames(const uint8_t* gray, int width, int height, int n_frames, const std::string& name_for_metadata,
                                  const std::string& mofreak_filename, bool clear_features_after_computation);
    // which CUDA device the next context is created on (default 0)
    void setDevice(int device) { device_ = device; }
    // rows of the last computeMoFREAKFrom* call in the binary layout of the C ABI
    const std::vector<mfk_row>& lastRows() const { return last_rows_; }
    const std::string& lastText() const { return last_text_; }
    // initial size of the host row/text buffers (rows per frame); too small only costs one more read of the device rows
    void setRowsPerFrameHint(size_t rows) { rows_per_frame_hint_ = rows ? rows : 1; }

private:
    void readMetadata(std::string filename, int& action, int& video_number, int& person);
    void ensureContext(int width, int height, int budget_scale = 0);
    [[noreturn]] void die(const std::string& what);

    std::deque<MoFREAKFeature> features;
    int dataset;
    int device_;
    mfk_ctx* ctx_;
    int ctx_w_, ctx_h_;
    int budget_scale_;     // 1, 4, 16, ... multiplier on the default per-pair candidate/row budgets
    std::vector<mfk_row> last_rows_;
    size_t rows_per_frame_hint_;
    std::string last_text_;                  // `.mofreak` text of the last clip as formatted on the device
};

// .mofreak text exactly as writeMoFREAKFeaturesToFile prints it (MoFREAKUtilities.cpp:691-719)
void formatMoFREAKRow(std::string& out, float x, float y, int frame, float scale, float motion_x, float motion_y,
                      const unsigned* appearance, const unsigned* motion);

#endif
