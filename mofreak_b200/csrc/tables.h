// Host-built constant tables the descriptor kernel reads.  They are computed ONCE per process with the
// host libm — the same libm the reference's CPU path would use on this machine — so that every float in
// them is bit-identical to what cv::FREAK::buildPattern produces there (OpenCV 2.4.2 freak.cpp; the
// reference rebuilds the table on every frame, MoFREAKUtilities.cpp:427).
#pragma once
#include <stdint.h>

namespace mfk {

enum {
    FREAK_SCALES = 64,
    FREAK_ORIENTATIONS = 256,
    FREAK_POINTS = 43,
    FREAK_ORIENT_PAIRS = 45,
    FREAK_USED_PAIRS = 64,     // the reference keeps descriptor bytes 0..7 only (MoFREAKUtilities.cpp:453-456)
};

struct alignas(16) FreakPoint { float x, y, sigma, pad; };   // offset from the keypoint and box half-size; one 128-bit load
struct OrientPair { int16_t i, j; int32_t wdx, wdy; };
struct UsedPair   { uint8_t i, j; };               // index = byte*8 + bit of the 8 appearance bytes

struct Tables {
    FreakPoint* points;                            // [scale][orientation][point], 704512 entries
    int32_t     pattern_size[FREAK_SCALES];        // border a keypoint of that scale must keep
    float       scale_threshold[FREAK_SCALES];     // [k] = smallest keypoint size whose scale index is >= k ([0] unused)
    OrientPair  orient[FREAK_ORIENT_PAIRS];
    UsedPair    used[FREAK_USED_PAIRS];
    // MIP: the pixels of the two 19x19 patches that the 64 SSD tests actually read (51 of the current-frame patch, 225 of
    // the previous-frame patch); entry = dx | dy << 5 | 0x8000 for the previous frame.  Only these are interpolated.
    uint16_t    mip_needed[2 * 361];
    int32_t     n_mip_needed;
};

const Tables& tables();                            // built on first use, then immutable
int scale_index_from_table(float kp_size);         // count of thresholds <= size
int scale_index_direct(float kp_size);             // the libm expression the thresholds were derived from

}  // namespace mfk
