// FREAK retina tables for the descriptor kernel (see tables.h).  Geometry and constants are those of
// cv::FREAK with default parameters (patternScale 22, 4 octaves, 64 scales x 256 orientations x 43
// points; OpenCV 2.4.2 modules/features2d/src/freak.cpp, un-vendored by the reference, call site
// MoFREAKUtilities.cpp:427-428).  Every value is evaluated in the same precision and operand order as
// there, because the floats themselves are part of the bit-exact contract.
#include "tables.h"

#include <math.h>
#include <string.h>
#include <mutex>
#include <vector>

namespace mfk {
namespace {

const double kPi = 3.1415926535897932384626433832795;
const double kLog2 = 0.693147180559945;
const float  kPatternScale = 22.0f;
const int    kOctaves = 4;
const int    kSmallestKeypoint = 7;

// indices (into the i>j enumeration of the 903 point pairs) of the first 128 default description
// pairs; the 8 appearance bytes only ever see pairs p < 128 with p % 16 >= 8
const uint16_t kDefaultPairs128[128] = {
    404, 431, 818, 511, 181, 52,  311, 874, 774, 543, 719, 230, 417, 205, 11,  560,
    149, 265, 39,  306, 165, 857, 250, 8,   61,  15,  55,  717, 44,  412, 592, 134,
    761, 695, 660, 782, 625, 487, 549, 516, 271, 665, 762, 392, 178, 796, 773, 31,
    672, 845, 548, 794, 677, 654, 241, 831, 225, 238, 849, 83,  691, 484, 826, 707,
    122, 517, 583, 731, 328, 339, 571, 475, 394, 472, 580, 381, 137, 93,  380, 327,
    619, 729, 808, 218, 213, 459, 141, 806, 341, 95,  382, 568, 124, 750, 193, 749,
    706, 843, 79,  199, 317, 329, 768, 198, 100, 466, 613, 78,  562, 783, 689, 136,
    838, 94,  142, 164, 679, 219, 419, 366, 418, 423, 77,  89,  523, 259, 683, 312};

struct Ring { int count; double radius, sigma; };

void retina(Ring ring[8]) {
    // seven rings of six receptive fields plus the centre, radius normalised so that the outermost
    // field (radius + sigma) reaches 1
    const double big = 2.0 / 3.0, small = 2.0 / 24.0;
    const double unit = (big - small) / 21.0;
    const int steps[8] = {0, 6, 11, 15, 18, 20, 0, 0};
    for (int r = 0; r < 8; r++) {
        ring[r].count = (r == 7) ? 1 : 6;
        ring[r].radius = (r < 6) ? big - steps[r] * unit : (r == 6 ? small : 0.0);
    }
    for (int r = 0; r < 8; r++) ring[r].sigma = (r == 7 ? ring[6].radius : ring[r].radius) / 2.0;
}

void fill_points(Tables& T) {
    Ring ring[8];
    retina(ring);
    const double step = pow(2.0, (double)(kOctaves) / FREAK_SCALES);
    for (int s = 0; s < FREAK_SCALES; ++s) {
        const double zoom = pow(step, s);
        int reach = 0;
        for (int o = 0; o < FREAK_ORIENTATIONS; ++o) {
            const double theta = double(o) * 2 * kPi / double(FREAK_ORIENTATIONS);
            FreakPoint* out = T.points + ((size_t)s * FREAK_ORIENTATIONS + o) * FREAK_POINTS;
            for (size_t r = 0; r < 8; ++r) {
                for (int k = 0; k < ring[r].count; ++k) {
                    const double stagger = kPi / ring[r].count * (r % 2);
                    const double alpha = double(k) * 2 * kPi / double(ring[r].count) + stagger + theta;
                    out->x = static_cast<float>(ring[r].radius * cos(alpha) * zoom * kPatternScale);
                    out->y = static_cast<float>(ring[r].radius * sin(alpha) * zoom * kPatternScale);
                    out->sigma = static_cast<float>(ring[r].sigma * zoom * kPatternScale);
                    const int need = static_cast<int>(ceil((ring[r].radius + ring[r].sigma) * zoom * kPatternScale)) + 1;
                    if (need > reach) reach = need;
                    ++out;
                }
            }
        }
        T.pattern_size[s] = reach;
    }
}

void fill_orientation_pairs(Tables& T) {
    // within each of the four outer rings: three diameters, then the six chords two fields apart
    static const int8_t chord[9][2] = {{0, 3}, {1, 4}, {2, 5}, {0, 2}, {1, 3}, {2, 4}, {3, 5}, {4, 0}, {5, 1}};
    int m = 0;
    for (int base = 0; base < 24; base += 6)
        for (int q = 0; q < 9; q++, m++) {
            T.orient[m].i = (int16_t)(base + chord[q][0]);
            T.orient[m].j = (int16_t)(base + chord[q][1]);
        }
    // the three inner rings contribute their diameters only
    for (int base = 24; base < 42; base += 6)
        for (int q = 0; q < 3; q++, m++) {
            T.orient[m].i = (int16_t)(base + q);
            T.orient[m].j = (int16_t)(base + q + 3);
        }
    const FreakPoint* p0 = T.points;   // scale 0, orientation 0
    for (int q = FREAK_ORIENT_PAIRS; q--;) {
        const float dx = p0[T.orient[q].i].x - p0[T.orient[q].j].x;
        const float dy = p0[T.orient[q].i].y - p0[T.orient[q].j].y;
        const float norm_sq = (dx * dx + dy * dy);
        T.orient[q].wdx = int((dx / (norm_sq)) * 4096.0 + 0.5);
        T.orient[q].wdy = int((dy / (norm_sq)) * 4096.0 + 0.5);
    }
}

void fill_used_pairs(Tables& T) {
    // descriptor layout of cv::FREAK (SSE2 and scalar builds agree): inside the first block of 128
    // pairs, pair p sets bit p/16 of byte 15 - p%16; appearance byte b, bit t is therefore pair 16t+15-b
    for (int b = 0; b < 8; b++)
        for (int t = 0; t < 8; t++) {
            const int flat = kDefaultPairs128[16 * t + 15 - b];
            // invert flat = i(i-1)/2 + j, 0 <= j < i
            int i = 1;
            while ((i + 1) * i / 2 <= flat) i++;
            T.used[b * 8 + t].i = (uint8_t)i;
            T.used[b * 8 + t].j = (uint8_t)(flat - i * (i - 1) / 2);
        }
}

void fill_mip_needed(Tables& T) {
    // centres MoFREAKUtilities.cpp:307-316, shifts :55-70; each test reads 9 consecutive bytes from the top-left corner of
    // its 3x3 region of the row-major 19x19 buffer (:79-88)
    static const int CX[8] = {5, 5, 5, 9, 9, 13, 13, 13}, CY[8] = {5, 9, 13, 5, 13, 5, 9, 13};
    static const int SX[8] = {-4, -3, 0, 3, 4, 3, 0, -3}, SY[8] = {0, 3, 4, 3, 0, -3, -4, -3};
    bool need[2][361] = {};
    for (int c = 0; c < 8; c++) {
        const int s = (CY[c] - 1) * 19 + CX[c] - 1;
        for (int i = 0; i < 9; i++) need[0][s + i] = true;
        for (int k = 0; k < 8; k++) {
            const int q = (CY[c] + SY[k] - 1) * 19 + CX[c] + SX[k] - 1;
            for (int i = 0; i < 9; i++) need[1][q + i] = true;
        }
    }
    int n = 0;
    for (int fr = 0; fr < 2; fr++)
        for (int p = 0; p < 361; p++)
            if (need[fr][p]) T.mip_needed[n++] = (uint16_t)((p % 19) | ((p / 19) << 5) | (fr << 15));     // dx | dy << 5 | frame << 15
    T.n_mip_needed = n;
}

inline float bits_to_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline uint32_t float_to_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

void fill_scale_thresholds(Tables& T) {
    // scale_index_direct is a non-decreasing step function of the (positive) size; positive floats
    // order like their bit patterns, so each step edge is found by bisection on the bits.
    T.scale_threshold[0] = 0.0f;
    uint32_t lo = float_to_bits(1.0f);
    for (int k = 1; k < FREAK_SCALES; k++) {
        uint32_t hi = float_to_bits(4096.0f);        // index(4096) == 63
        uint32_t l = lo;                              // index(l) < k
        while (hi - l > 1) {
            const uint32_t mid = l + (hi - l) / 2;
            if (scale_index_direct(bits_to_float(mid)) >= k) hi = mid; else l = mid;
        }
        T.scale_threshold[k] = bits_to_float(hi);
        lo = l;
    }
}

Tables* g_tables = nullptr;
std::once_flag g_once;

}  // namespace

// kpScaleIdx of cv::FREAK::computeImpl: max(int(log(size / 7) * sizeCst + 0.5), 0), clamped to 63; log on
// a float argument is logf and the product stays in float until the + 0.5
int scale_index_direct(float kp_size) {
    const float sizeCst = static_cast<float>(FREAK_SCALES / (kLog2 * kOctaves));
    int idx = (int)(logf(kp_size / kSmallestKeypoint) * sizeCst + 0.5);
    if (idx < 0) idx = 0;
    if (idx >= FREAK_SCALES) idx = FREAK_SCALES - 1;
    return idx;
}

int scale_index_from_table(float kp_size) {
    const Tables& T = tables();
    int idx = 0;
    for (int k = 1; k < FREAK_SCALES; k++) idx += (kp_size >= T.scale_threshold[k]);
    return idx;
}

const Tables& tables() {
    std::call_once(g_once, [] {
        Tables* T = new Tables();
        T->points = new FreakPoint[(size_t)FREAK_SCALES * FREAK_ORIENTATIONS * FREAK_POINTS]();
        fill_points(*T);
        fill_orientation_pairs(*T);
        fill_used_pairs(*T);
        fill_mip_needed(*T);
        g_tables = T;                 // scale_index_from_table() is not used while filling
        fill_scale_thresholds(*T);
    });
    return *g_tables;
}

}  // namespace mfk
