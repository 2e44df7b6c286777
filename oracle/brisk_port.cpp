// TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// Restatement of the reference's vendored BRISK scale-space detector, reference
// src/MoFREAK/brisk.cpp:544-704 (driver), :838-934 (isMax2D), :937-1103 (refine3D),
// :1106-1416 (getScoreMaxAbove/Below), :1418-1533 (refine1D*), :1535-1644 (subpixel2D),
// :1647-1725 (layers, lazy scores), :1840-2065 (pyramid samplers), plus the segment tests of
// oast9_16.cc / oast9_16_nms.cc / agast5_8_nms.cc.
//
// It is NOT a transliteration.  The reference walks candidates sequentially and its lazily
// memoised score map makes isMax2D order dependent.  Here the same result is produced in the
// phase-structured form the CUDA kernels use:
//   * scores are the closed form S_full = clamp(max_arc min_i(+-(v_i - p)) - 1, 0, 254) (the value the
//     bisection of cornerScore converges to from b0 = 0), zero inside the 3-px border;
//   * what the reference's raw score map would hold at a given moment is S_full masked by a
//     visibility predicate: detected (S >= thresh), or touched by a getScoreMaxAbove scan from the
//     layer below, or lying in the 3x3 patch of a raster-earlier candidate that survived all tests;
//   * candidates without an equal-score neighbour are decided independently (phase A); the few with
//     ties are resolved by a fixed point that only needs "every raster-earlier undecided tie
//     candidate within 3 px is final" (phase B) — any schedule that respects that gives the same
//     answer, which tests check by also running the least favourable one (jacobi_reverse).
// Pinned against the reference's object code (oracle/_ref) by tests/test_oracle_brisk.py.
#include "oracle.h"

#include <math.h>
#include <string.h>
#include <algorithm>

namespace orc {

// ------------------------------------------------------------------------------------------------
// pyramid samplers
// ------------------------------------------------------------------------------------------------
static inline int avg_up(int a, int b) { return (a + b + 1) >> 1; }   // _mm_avg_epu8

// brisk.cpp:1840-1972.  Per destination pixel, which of three arithmetic rules applies depends only on
// its column: 16-wide SSE blocks are consumed in pairs (avg of avg, both rounding up); an odd
// trailing block is reduced with a truncating pair mean; the remaining (cols % 16) / 2 outputs use a
// 4-tap truncating mean over columns k, k+1 (not 2k, 2k+1) counted from the end of the blocks.
void port_halfsample(const uint8_t* src, int w, int h, uint8_t* dst) {
    const int dw = w / 2, dh = h / 2;
    const int blocks = w / 16;
    const int paired_out = (blocks / 2) * 16;
    const bool odd_block = (blocks & 1) != 0;
    const int tail = (w % 16) / 2;
    for (int y = 0; y < dh; y++) {
        const uint8_t* up = src + (size_t)(2 * y) * w;
        const uint8_t* lo = up + w;
        uint8_t* d = dst + (size_t)y * dw;
        int o = 0;
        for (; o < paired_out; o++)
            d[o] = (uint8_t)avg_up(avg_up(up[2 * o], lo[2 * o]), avg_up(up[2 * o + 1], lo[2 * o + 1]));
        int base = 2 * paired_out;
        if (odd_block) {
            for (int j = 0; j < 8; j++, o++)
                d[o] = (uint8_t)((avg_up(up[base + 2 * j], lo[base + 2 * j]) +
                                  avg_up(up[base + 2 * j + 1], lo[base + 2 * j + 1])) / 2);
            base += 16;
        }
        for (int k = 0; k < tail; k++, o++)
            d[o] = (uint8_t)((up[base + k] + up[base + k + 1] + lo[base + k] + lo[base + k + 1]) / 4);
    }
}

// brisk.cpp:1974-2065.  15 source columns -> 10 outputs through avg(avg(a,b),a) chains, vertically
// then horizontally; the fifth triple of each block takes its middle sample from byte 12 (shuffle
// masks :1982-1983); trailing triples use the (4A+2(B+C)+D)/9 scalar rule.
void port_twothirdsample(const uint8_t* src, int w, int h, uint8_t* dst) {
    const int dw = 2 * (w / 3);
    const int blocks = w / 15;
    const int tail_cols = ((w / 3) * 3) % 15;
    for (int rt = 0; rt < h / 3; rt++) {
        const uint8_t* r1 = src + (size_t)(3 * rt) * w;
        const uint8_t* r2 = r1 + w;
        const uint8_t* r3 = r2 + w;
        uint8_t* d1 = dst + (size_t)(2 * rt) * dw;
        uint8_t* d2 = d1 + dw;
        for (int i = 0; i < blocks; i++) {
            int u[15], l[15];
            for (int c = 0; c < 15; c++) {
                const int a = r1[15 * i + c], b = r2[15 * i + c], cc = r3[15 * i + c];
                u[c] = avg_up(avg_up(a, b), a);
                l[c] = avg_up(avg_up(cc, b), cc);
            }
            for (int t = 0; t < 5; t++) {
                const int mid = (t == 4) ? 12 : 3 * t + 1;
                d1[10 * i + 2 * t]     = (uint8_t)avg_up(avg_up(u[3 * t], u[mid]), u[3 * t]);
                d1[10 * i + 2 * t + 1] = (uint8_t)avg_up(avg_up(u[3 * t + 2], u[mid]), u[3 * t + 2]);
                d2[10 * i + 2 * t]     = (uint8_t)avg_up(avg_up(l[3 * t], l[mid]), l[3 * t]);
                d2[10 * i + 2 * t + 1] = (uint8_t)avg_up(avg_up(l[3 * t + 2], l[mid]), l[3 * t + 2]);
            }
        }
        int sc = 15 * blocks, o = 10 * blocks;
        for (int j = 0; j < tail_cols; j += 3, sc += 3, o += 2) {
            const int A1 = r1[sc], A2 = r1[sc + 1], A3 = r1[sc + 2];
            const int B1 = r2[sc], B2 = r2[sc + 1], B3 = r2[sc + 2];
            const int C1 = r3[sc], C2 = r3[sc + 1], C3 = r3[sc + 2];
            d1[o]     = (uint8_t)((4 * A1 + 2 * (A2 + B1) + B2) / 9);
            d1[o + 1] = (uint8_t)((4 * A3 + 2 * (A2 + B3) + B2) / 9);
            d2[o]     = (uint8_t)((4 * C1 + 2 * (C2 + B1) + B2) / 9);
            d2[o + 1] = (uint8_t)((4 * C3 + 2 * (C2 + B3) + B2) / 9);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// segment-test scores, closed form
// ------------------------------------------------------------------------------------------------
// circle offsets: oast9_16.h:214-232 and agast5_8.h:65-75 (index order = position on the circle)
static const int RING16_X[16] = {-3, -3, -2, -1, 0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3};
static const int RING16_Y[16] = {0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3, 3, 3, 2, 1};
static const int RING8_X[8] = {-1, -1, 0, 1, 1, 1, 0, -1};
static const int RING8_Y[8] = {0, -1, -1, -1, 0, 1, 1, 1};

// largest b in [0,254] for which ARC contiguous ring pixels are all > p+b or all < p-b; 0 if none
template <int N, int ARC>
static inline int arc_score(const int* d) {
    int best = -256;
    for (int s = 0; s < N; s++) {
        int mn = 255, mx = -255;
        for (int k = 0; k < ARC; k++) {
            const int v = d[(s + k) % N];
            mn = std::min(mn, v);
            mx = std::max(mx, v);
        }
        best = std::max(best, std::max(mn, -mx));
    }
    return std::min(254, std::max(0, best - 1));
}

int port_oast_score(const uint8_t* img, int w, int x, int y) {
    const uint8_t* p = img + (size_t)y * w + x;
    int d[16];
    for (int i = 0; i < 16; i++) d[i] = (int)p[RING16_Y[i] * w + RING16_X[i]] - (int)*p;
    return arc_score<16, 9>(d);
}

int port_agast58_score(const uint8_t* img, int w, int x, int y) {
    const uint8_t* p = img + (size_t)y * w + x;
    int d[8];
    for (int i = 0; i < 8; i++) d[i] = (int)p[RING8_Y[i] * w + RING8_X[i]] - (int)*p;
    return arc_score<8, 5>(d);
}

void port_score_map(const uint8_t* img, int w, int h, uint8_t* S) {
    memset(S, 0, (size_t)w * h);
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) S[(size_t)y * w + x] = (uint8_t)port_oast_score(img, w, x, y);
}

// ------------------------------------------------------------------------------------------------
// scale space
// ------------------------------------------------------------------------------------------------
namespace {

enum { ST_PASS2D = 1, ST_SURVIVE = 2, ST_PENDING = 4 };

struct PLayer {
    int w, h;
    float scale, offset;
    std::vector<uint8_t> img;
    std::vector<uint8_t> S;         // S_full, computed on first use (pure, so laziness is unobservable)
    std::vector<uint8_t> known;     // S computed?
    std::vector<uint8_t> seen;      // touched by a getScoreMaxAbove scan issued from the layer below
    std::vector<uint8_t> status;    // per candidate pixel: ST_* bits
    std::vector<int> cand;          // raster indices with S >= thresh, ascending

    void init(int w_, int h_, float scale_) {
        w = w_; h = h_; scale = scale_;
        offset = 0.5 * scale - 0.5;                              // brisk.cpp:1663, 1669
        img.assign((size_t)w * h, 0);
        S.assign((size_t)w * h, 0);
        known.assign((size_t)w * h, 0);
        seen.assign((size_t)w * h, 0);
        status.assign((size_t)w * h, 0);
    }
    // BriskLayer::getAgastScore(int,int,1), brisk.cpp:1691-1700, minus the memo side effect
    inline int score(int x, int y) {
        if (x < 3 || y < 3 || x >= w - 3 || y >= h - 3) return 0;
        const size_t i = (size_t)y * w + x;
        if (!known[i]) { S[i] = (uint8_t)port_oast_score(img.data(), w, x, y); known[i] = 1; }
        return S[i];
    }
    // BriskLayer::getAgastScore_5_8, brisk.cpp:1702-1709
    inline int score58(int x, int y) {
        if (x < 2 || y < 2 || x >= w - 2 || y >= h - 2) return 0;
        return port_agast58_score(img.data(), w, x, y);
    }
};

// Reads of a neighbouring layer's scores.  `mark` records the visibility side effect.
struct Access {
    PLayer& L;
    bool mark;
    inline int at(int x, int y) {
        if (x < 3 || y < 3 || x >= L.w - 3 || y >= L.h - 3) return 0;
        if (mark) L.seen[(size_t)y * L.w + x] = 1;
        return L.score(x, y);
    }
    // BriskLayer::getAgastScore(float,float,1) with scale <= 1, brisk.cpp:1711-1725: bilinear blend in
    // float, truncated to uint8
    inline float atf(float xf, float yf) {
        const int x = int(xf);
        const float rx1 = xf - float(x);
        const float rx = 1.0f - rx1;
        const int y = int(yf);
        const float ry1 = yf - float(y);
        const float ry = 1.0f - ry1;
        const int s00 = at(x, y), s10 = at(x + 1, y), s01 = at(x, y + 1), s11 = at(x + 1, y + 1);
        const float v = rx * ry * s00 + rx1 * ry * s10 + rx * ry1 * s01 + rx1 * ry1 * s11;
        return (float)(uint8_t)v;
    }
};

// brisk.cpp:1535-1644
float subpixel2D(const int s_0_0, const int s_0_1, const int s_0_2, const int s_1_0, const int s_1_1,
                 const int s_1_2, const int s_2_0, const int s_2_1, const int s_2_2, float& delta_x,
                 float& delta_y) {
    // least-squares quadratic fit on the 3x3 patch
    const int tmp1 = s_0_0 + s_0_2 - 2 * s_1_1 + s_2_0 + s_2_2;
    const int coeff1 = 3 * (tmp1 + s_0_1 - ((s_1_0 + s_1_2) << 1) + s_2_1);
    const int coeff2 = 3 * (tmp1 - ((s_0_1 + s_2_1) << 1) + s_1_0 + s_1_2);
    const int tmp2 = s_0_2 - s_2_0;
    const int tmp3 = (s_0_0 + tmp2 - s_2_2);
    const int tmp4 = tmp3 - 2 * tmp2;
    const int coeff3 = -3 * (tmp3 + s_0_1 - s_2_1);
    const int coeff4 = -3 * (tmp4 + s_1_0 - s_1_2);
    const int coeff5 = (s_0_0 - s_0_2 - s_2_0 + s_2_2) << 2;
    const int coeff6 = -(s_0_0 + s_0_2 - ((s_1_0 + s_0_1 + s_1_2 + s_2_1) << 1) - 5 * s_1_1 + s_2_0 + s_2_2) << 1;

    const int H_det = 4 * coeff1 * coeff2 - coeff5 * coeff5;
    if (H_det == 0) {
        delta_x = 0.0f;
        delta_y = 0.0f;
        return float(coeff6) / 18.0;
    }
    if (!(H_det > 0 && coeff1 < 0)) {
        // maximum on one of the four patch corners
        int tmp_max = coeff3 + coeff4 + coeff5;
        delta_x = 1.0f; delta_y = 1.0f;
        int tmp = -coeff3 + coeff4 - coeff5;
        if (tmp > tmp_max) { tmp_max = tmp; delta_x = -1.0f; delta_y = 1.0f; }
        tmp = coeff3 - coeff4 - coeff5;
        if (tmp > tmp_max) { tmp_max = tmp; delta_x = 1.0f; delta_y = -1.0f; }
        tmp = -coeff3 - coeff4 + coeff5;
        if (tmp > tmp_max) { tmp_max = tmp; delta_x = -1.0f; delta_y = -1.0f; }
        return float(tmp_max + coeff1 + coeff2 + coeff6) / 18.0;
    }
    delta_x = float(2 * coeff2 * coeff3 - coeff4 * coeff5) / float(-H_det);
    delta_y = float(2 * coeff1 * coeff4 - coeff3 * coeff5) / float(-H_det);
    bool tx = false, tx_ = false, ty = false, ty_ = false;
    if (delta_x > 1.0) tx = true;
    else if (delta_x < -1.0) tx_ = true;
    if (delta_y > 1.0) ty = true;
    if (delta_y < -1.0) ty_ = true;

    if (tx || tx_ || ty || ty_) {
        float delta_x1 = 0.0f, delta_x2 = 0.0f, delta_y1 = 0.0f, delta_y2 = 0.0f;
        if (tx) {
            delta_x1 = 1.0f;
            delta_y1 = -float(coeff4 + coeff5) / float(2 * coeff2);
            if (delta_y1 > 1.0) delta_y1 = 1.0f; else if (delta_y1 < -1.0) delta_y1 = -1.0f;
        } else if (tx_) {
            delta_x1 = -1.0f;
            delta_y1 = -float(coeff4 - coeff5) / float(2 * coeff2);
            if (delta_y1 > 1.0) delta_y1 = 1.0f; else if (delta_y1 < -1.0) delta_y1 = -1.0f;
        }
        if (ty) {
            delta_y2 = 1.0f;
            delta_x2 = -float(coeff3 + coeff5) / float(2 * coeff1);
            if (delta_x2 > 1.0) delta_x2 = 1.0f; else if (delta_x2 < -1.0) delta_x2 = -1.0f;
        } else if (ty_) {
            delta_y2 = -1.0f;
            delta_x2 = -float(coeff3 - coeff5) / float(2 * coeff1);
            if (delta_x2 > 1.0) delta_x2 = 1.0f; else if (delta_x2 < -1.0) delta_x2 = -1.0f;
        }
        const float max1 = (coeff1 * delta_x1 * delta_x1 + coeff2 * delta_y1 * delta_y1 + coeff3 * delta_x1 +
                            coeff4 * delta_y1 + coeff5 * delta_x1 * delta_y1 + coeff6) / 18.0;
        const float max2 = (coeff1 * delta_x2 * delta_x2 + coeff2 * delta_y2 * delta_y2 + coeff3 * delta_x2 +
                            coeff4 * delta_y2 + coeff5 * delta_x2 * delta_y2 + coeff6) / 18.0;
        // the reference returns delta_y = delta_x here (brisk.cpp:1627-1636); kept
        if (max1 > max2) { delta_x = delta_x1; delta_y = delta_x1; return max1; }
        delta_x = delta_x2; delta_y = delta_x2; return max2;
    }
    return (coeff1 * delta_x * delta_x + coeff2 * delta_y * delta_y + coeff3 * delta_x + coeff4 * delta_y +
            coeff5 * delta_x * delta_y + coeff6) / 18.0;
}

// the three parabola fits over relative scale, brisk.cpp:1418-1533; sample positions lo < 1 < hi and
// integer coefficient rows {a, b, c}, value divisor den
struct Fit1D { int a[3], b[3], c[3]; double lo, hi; double den; };
static const Fit1D FIT_OCTAVE = {{16, -24, 8}, {-40, 54, -14}, {24, -27, 6}, 0.75, 1.5, 3072.0};               // refine1D
static const Fit1D FIT_INTRA  = {{9, -18, 9}, {-21, 36, -15}, {12, -16, 6},
                                 0.6666666666666666666666666667, 1.3333333333333333333333333333, 2048.0};  // refine1D_1
static const Fit1D FIT_BASE   = {{2, -4, 2}, {-5, 8, -3}, {3, -3, 1}, 0.7, 1.5, 1024.0};                    // refine1D_2

float refine1D_fit(const Fit1D& F, const float s_05, const float s0, const float s05, float& max) {
    const int i_05 = int(1024.0 * s_05 + 0.5);
    const int i0 = int(1024.0 * s0 + 0.5);
    const int i05 = int(1024.0 * s05 + 0.5);
    const int qa = F.a[0] * i_05 + F.a[1] * i0 + F.a[2] * i05;
    if (qa >= 0) {   // not concave: best sample
        if (s0 >= s_05 && s0 >= s05) { max = s0; return 1.0f; }
        if (s_05 >= s0 && s_05 >= s05) { max = s_05; return (float)F.lo; }
        if (s05 >= s0 && s05 >= s_05) { max = s05; return (float)F.hi; }
    }
    const int qb = F.b[0] * i_05 + F.b[1] * i0 + F.b[2] * i05;
    float ret_val = -float(qb) / float(2 * qa);
    if (ret_val < F.lo) ret_val = (float)F.lo;
    else if (ret_val > F.hi) ret_val = (float)F.hi;
    const int qc = F.c[0] * i_05 + F.c[1] * i0 + F.c[2] * i05;
    max = float(qc) + float(qa) * ret_val * ret_val + float(qb) * ret_val;
    max = (float)(max / F.den);
    return ret_val;
}

struct Scan { bool ismax; float max, dx, dy; };

// getScoreMaxAbove (below == false, brisk.cpp:1106-1249) / getScoreMaxBelow (below == true, :1251-1416)
// over neighbour layer N for the candidate (x_layer, y_layer) of layer `layer`.
Scan scan_neighbour(Access N, const bool below, const int layer, const int x_layer, const int y_layer,
                    const int threshold) {
    Scan R; R.ismax = false; R.max = 0.0f; R.dx = 0.0f; R.dy = 0.0f;
    float x_1, x1, y_1, y1;
    if (!below) {
        if (layer % 2 == 0) {
            x_1 = float(4 * (x_layer) - 1 - 2) / 6.0;  x1 = float(4 * (x_layer) - 1 + 2) / 6.0;
            y_1 = float(4 * (y_layer) - 1 - 2) / 6.0;  y1 = float(4 * (y_layer) - 1 + 2) / 6.0;
        } else {
            x_1 = float(6 * (x_layer) - 1 - 3) / 8.0f; x1 = float(6 * (x_layer) - 1 + 3) / 8.0f;
            y_1 = float(6 * (y_layer) - 1 - 3) / 8.0f; y1 = float(6 * (y_layer) - 1 + 3) / 8.0f;
        }
    } else {
        if (layer % 2 == 0) {
            x_1 = float(8 * (x_layer) + 1 - 4) / 6.0;  x1 = float(8 * (x_layer) + 1 + 4) / 6.0;
            y_1 = float(8 * (y_layer) + 1 - 4) / 6.0;  y1 = float(8 * (y_layer) + 1 + 4) / 6.0;
        } else {
            x_1 = float(6 * (x_layer) + 1 - 3) / 4.0;  x1 = float(6 * (x_layer) + 1 + 3) / 4.0;
            y_1 = float(6 * (y_layer) + 1 - 3) / 4.0;  y1 = float(6 * (y_layer) + 1 + 3) / 4.0;
        }
    }
    const int xa = int(x_1 + 1), xb = int(x1);      // interior integer columns xa..xb
    const int ya = int(y_1 + 1), yb = int(y1);      // interior integer rows ya..yb
    int max_x = xa, max_y = ya;
    float tmp;
    float max = N.atf(x_1, y_1);
    // first row (fractional y): early exit when any sample beats the candidate
    if (max > threshold) return R;
    for (int x = xa; x <= xb; x++) {
        tmp = N.atf(float(x), y_1);
        if (tmp > threshold) return R;
        if (tmp > max) { max = tmp; max_x = x; }
    }
    tmp = N.atf(x1, y_1);
    if (tmp > threshold) return R;
    if (tmp > max) { max = tmp; max_x = xb; }
    // middle rows (integer y)
    for (int y = ya; y <= yb; y++) {
        tmp = N.atf(x_1, float(y));
        if (tmp > threshold) return R;
        if (tmp > max) { max = tmp; max_x = xa; max_y = y; }
        for (int x = xa; x <= xb; x++) {
            tmp = (float)N.at(x, y);
            if (tmp > threshold) return R;
            if (below && tmp == max) {
                // equal interior maxima: prefer the one with the larger 8-neighbour sum (:1314-1337)
                const int t1 = 2 * (N.at(x - 1, y) + N.at(x + 1, y) + N.at(x, y + 1) + N.at(x, y - 1)) +
                               (N.at(x + 1, y + 1) + N.at(x - 1, y + 1) + N.at(x + 1, y - 1) + N.at(x - 1, y - 1));
                const int t2 = 2 * (N.at(max_x - 1, max_y) + N.at(max_x + 1, max_y) + N.at(max_x, max_y + 1) +
                                    N.at(max_x, max_y - 1)) +
                               (N.at(max_x + 1, max_y + 1) + N.at(max_x - 1, max_y + 1) +
                                N.at(max_x + 1, max_y - 1) + N.at(max_x - 1, max_y - 1));
                if (t1 > t2) { max_x = x; max_y = y; }
            }
            if (tmp > max) { max = tmp; max_x = x; max_y = y; }
        }
        tmp = N.atf(x1, float(y));
        if (tmp > threshold) return R;
        if (tmp > max) { max = tmp; max_x = xb; max_y = y; }
    }
    // bottom row (fractional y): no early exit here (:1186-1206, :1353-1373)
    tmp = N.atf(x_1, y1);
    if (tmp > max) { max = tmp; max_x = xa; max_y = yb; }
    for (int x = xa; x <= xb; x++) {
        tmp = N.atf(float(x), y1);
        if (tmp > max) { max = tmp; max_x = x; max_y = yb; }
    }
    tmp = N.atf(x1, y1);
    if (tmp > max) { max = tmp; max_x = xb; max_y = yb; }

    // sub-pixel refinement around the integer maximum
    const int s_0_0 = N.at(max_x - 1, max_y - 1), s_1_0 = N.at(max_x, max_y - 1), s_2_0 = N.at(max_x + 1, max_y - 1);
    const int s_2_1 = N.at(max_x + 1, max_y), s_1_1 = N.at(max_x, max_y), s_0_1 = N.at(max_x - 1, max_y);
    const int s_0_2 = N.at(max_x - 1, max_y + 1), s_1_2 = N.at(max_x, max_y + 1), s_2_2 = N.at(max_x + 1, max_y + 1);
    float dx_1, dy_1;
    const float refined_max = subpixel2D(s_0_0, s_0_1, s_0_2, s_1_0, s_1_1, s_1_2, s_2_0, s_2_1, s_2_2, dx_1, dy_1);
    const float real_x = float(max_x) + dx_1;
    const float real_y = float(max_y) + dy_1;
    bool returnrefined = true;
    float dx, dy;
    if (!below) {
        if (layer % 2 == 0) {
            dx = (real_x * 6.0f + 1.0f) / 4.0f - float(x_layer);
            dy = (real_y * 6.0f + 1.0f) / 4.0f - float(y_layer);
        } else {
            dx = (real_x * 8.0 + 1.0) / 6.0 - float(x_layer);
            dy = (real_y * 8.0 + 1.0) / 6.0 - float(y_layer);
        }
    } else {
        if (layer % 2 == 0) {
            dx = (real_x * 6.0 + 1.0) / 8.0 - float(x_layer);
            dy = (real_y * 6.0 + 1.0) / 8.0 - float(y_layer);
        } else {
            dx = (real_x * 4.0 - 1.0) / 6.0 - float(x_layer);
            dy = (real_y * 4.0 - 1.0) / 6.0 - float(y_layer);
        }
    }
    if (dx > 1.0f) { dx = 1.0f; returnrefined = false; }
    if (dx < -1.0f) { dx = -1.0f; returnrefined = false; }
    if (dy > 1.0f) { dy = 1.0f; returnrefined = false; }
    if (dy < -1.0f) { dy = -1.0f; returnrefined = false; }
    R.ismax = true;
    R.dx = dx; R.dy = dy;
    R.max = returnrefined ? std::max(refined_max, max) : max;
    return R;
}

struct Space {
    PLayer L[6];
    int thresh;

    // BriskScaleSpace::constructPyramid, brisk.cpp:572-588 with the layer constructors :1647-1674
    void build(const uint8_t* img, int w, int h) {
        L[0].init(w, h, 1.0f);
        memcpy(L[0].img.data(), img, (size_t)w * h);
        L[1].init(2 * (w / 3), 2 * (h / 3), 1.0f * 1.5f);
        port_twothirdsample(L[0].img.data(), w, h, L[1].img.data());
        for (int i = 2; i < 6; i += 2) {
            L[i].init(L[i - 2].w / 2, L[i - 2].h / 2, L[i - 2].scale * 2);
            port_halfsample(L[i - 2].img.data(), L[i - 2].w, L[i - 2].h, L[i].img.data());
            L[i + 1].init(L[i - 1].w / 2, L[i - 1].h / 2, L[i - 1].scale * 2);
            port_halfsample(L[i - 1].img.data(), L[i - 1].w, L[i - 1].h, L[i + 1].img.data());
        }
    }

    // getAgastPoints, brisk.cpp:1678-1690 + oast9_16.cc:46: a hit at threshold b has S_full >= b.
    // An arc of 9 contains one pixel of every antipodal pair, which gives a cheap reject.
    void find_candidates(PLayer& P) {
        P.cand.clear();
        const int w = P.w;
        for (int y = 3; y < P.h - 3; y++) {
            const uint8_t* row = P.img.data() + (size_t)y * w;
            for (int x = 3; x < w - 3; x++) {
                const int p = row[x];
                const int a = abs((int)row[x - 3] - p), b = abs((int)row[x + 3] - p);
                if (a <= thresh && b <= thresh) continue;
                const int c = abs((int)row[x - 3 * w] - p), d = abs((int)row[x + 3 * w] - p);
                if (c <= thresh && d <= thresh) continue;
                if (P.score(x, y) >= thresh) P.cand.push_back(y * w + x);
            }
        }
    }

    // the order-independent part of isMax2D (brisk.cpp:838-869): 0 = not a maximum, 1 = strict,
    // 2 = has equal neighbours
    int max8(PLayer& P, int x, int y) {
        const int c = P.score(x, y);
        bool tie = false;
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
                if (!dx && !dy) continue;
                const int s = P.score(x + dx, y + dy);
                if (s > c) return 0;
                if (s == c) tie = true;
            }
        return tie ? 2 : 1;
    }

    // everything that follows a successful isMax2D up to (not including) the own-layer 3x3 reads
    void after_max2d(int li, int x, int y) {
        PLayer& P = L[li];
        uint8_t& st = P.status[(size_t)y * P.w + x];
        st = ST_PASS2D;
        const int center = P.score(x, y);
        if (li < 5) {
            const Scan up = scan_neighbour(Access{L[li + 1], true}, false, li, x, y, center);
            if (!up.ismax) return;
            if (li > 0) {
                const Scan dn = scan_neighbour(Access{L[li - 1], false}, true, li, x, y, center);
                if (!dn.ismax) return;
            }
        } else {
            const Scan dn = scan_neighbour(Access{L[li - 1], false}, true, li, x, y, center);
            if (!dn.ismax) return;
        }
        st |= ST_SURVIVE;
    }

    // what the reference's raw score map holds at (qx,qy) when candidate `cidx` runs isMax2D
    int raw_at(PLayer& P, int qx, int qy, int cidx) {
        const int s = P.score(qx, qy);
        if (s == 0 || s >= thresh) return s;
        if (P.seen[(size_t)qy * P.w + qx]) return s;
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
                const int nx = qx + dx, ny = qy + dy;
                const int nidx = ny * P.w + nx;
                if (nidx >= cidx) continue;
                if (P.score(nx, ny) < thresh) continue;
                if (P.status[nidx] & ST_SURVIVE) return s;
            }
        return 0;
    }

    // the tie-break of isMax2D, brisk.cpp:870-933
    bool tie_break(PLayer& P, int x, int y) {
        const int cidx = y * P.w + x;
        const int center = P.score(x, y);
        int r[5][5];
        for (int dy = -2; dy <= 2; dy++)
            for (int dx = -2; dx <= 2; dx++) r[dy + 2][dx + 2] = raw_at(P, x + dx, y + dy, cidx);
        auto smooth = [&](int cx, int cy) {   // 1-2-1 x 1-2-1 around window position (cx,cy)
            return 4 * r[cy][cx] + 2 * (r[cy][cx - 1] + r[cy][cx + 1] + r[cy - 1][cx] + r[cy + 1][cx]) +
                   r[cy - 1][cx - 1] + r[cy - 1][cx + 1] + r[cy + 1][cx - 1] + r[cy + 1][cx + 1];
        };
        const int mine = smooth(2, 2);
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
                if (!dx && !dy) continue;
                if (r[dy + 2][dx + 2] != center) continue;
                if (smooth(dx + 2, dy + 2) > mine) return false;
            }
        return true;
    }

    bool earlier_pending_nearby(PLayer& P, int x, int y) {
        const int cidx = y * P.w + x;
        for (int dy = -3; dy <= 3; dy++)
            for (int dx = -3; dx <= 3; dx++) {
                const int nx = x + dx, ny = y + dy;
                if (nx < 0 || ny < 0 || nx >= P.w || ny >= P.h) continue;
                const int nidx = ny * P.w + nx;
                if (nidx >= cidx) continue;
                if (P.status[nidx] & ST_PENDING) return true;
            }
        return false;
    }

    // refine3D (brisk.cpp:937-1103) / the top-layer branch of getKeypoints (:645-680) for a survivor
    bool emit(int li, int x, int y, Keypoint& kp) {
        PLayer& P = L[li];
        const int center = P.score(x, y);
        auto patch = [&](float& ddx, float& ddy) {
            const int s_0_0 = P.score(x - 1, y - 1), s_1_0 = P.score(x, y - 1), s_2_0 = P.score(x + 1, y - 1);
            const int s_2_1 = P.score(x + 1, y), s_1_1 = P.score(x, y), s_0_1 = P.score(x - 1, y);
            const int s_0_2 = P.score(x - 1, y + 1), s_1_2 = P.score(x, y + 1), s_2_2 = P.score(x + 1, y + 1);
            return subpixel2D(s_0_0, s_0_1, s_0_2, s_1_0, s_1_1, s_1_2, s_2_0, s_2_1, s_2_2, ddx, ddy);
        };
        if (li == 5) {
            float delta_x, delta_y;
            const float max = patch(delta_x, delta_y);
            kp.x = (float(x) + delta_x) * P.scale + P.offset;
            kp.y = (float(y) + delta_y) * P.scale + P.offset;
            kp.size = 12.0f * P.scale;
            kp.response = max;
            kp.layer = li;
            return true;
        }
        const Scan up = scan_neighbour(Access{L[li + 1], false}, false, li, x, y, center);
        float max, scale, kx, ky;
        float delta_x_below, delta_y_below, max_below;
        if (li == 0) {
            // no layer below: 5/8 scores on the same image stand in for it (:959-986)
            const int s_0_0 = P.score58(x - 1, y - 1), s_1_0 = P.score58(x, y - 1), s_2_0 = P.score58(x + 1, y - 1);
            const int s_2_1 = P.score58(x + 1, y), s_1_1 = P.score58(x, y), s_0_1 = P.score58(x - 1, y);
            const int s_0_2 = P.score58(x - 1, y + 1), s_1_2 = P.score58(x, y + 1), s_2_2 = P.score58(x + 1, y + 1);
            int m = s_0_0;
            m = std::max(m, s_1_0); m = std::max(m, s_2_0); m = std::max(m, s_2_1); m = std::max(m, s_1_1);
            m = std::max(m, s_0_1); m = std::max(m, s_0_2); m = std::max(m, s_1_2); m = std::max(m, s_2_2);
            subpixel2D(s_0_0, s_0_1, s_0_2, s_1_0, s_1_1, s_1_2, s_2_0, s_2_1, s_2_2, delta_x_below, delta_y_below);
            max_below = (float)(uint8_t)m;
        } else {
            const Scan dn = scan_neighbour(Access{L[li - 1], false}, true, li, x, y, center);
            max_below = dn.max; delta_x_below = dn.dx; delta_y_below = dn.dy;
        }
        float delta_x_layer, delta_y_layer;
        const float max_layer = patch(delta_x_layer, delta_y_layer);
        const float mid = std::max(float(center), max_layer);
        if (li % 2 == 0) {
            scale = refine1D_fit(li == 0 ? FIT_BASE : FIT_OCTAVE, max_below, mid, up.max, max);
            if (scale > 1.0) {
                const float r0 = (1.5 - scale) / .5;
                const float r1 = 1.0 - r0;
                kx = (r0 * delta_x_layer + r1 * up.dx + float(x)) * P.scale + P.offset;
                ky = (r0 * delta_y_layer + r1 * up.dy + float(y)) * P.scale + P.offset;
            } else if (li == 0) {
                const float r0 = (scale - 0.5) / 0.5;
                const float r_1 = 1.0 - r0;
                kx = r0 * delta_x_layer + r_1 * delta_x_below + float(x);
                ky = r0 * delta_y_layer + r_1 * delta_y_below + float(y);
            } else {
                const float r0 = (scale - 0.75) / 0.25;
                const float r_1 = 1.0 - r0;
                kx = (r0 * delta_x_layer + r_1 * delta_x_below + float(x)) * P.scale + P.offset;
                ky = (r0 * delta_y_layer + r_1 * delta_y_below + float(y)) * P.scale + P.offset;
            }
        } else {
            scale = refine1D_fit(FIT_INTRA, max_below, mid, up.max, max);
            if (scale > 1.0) {
                const float r0 = 4.0 - scale * 3.0;
                const float r1 = 1.0 - r0;
                kx = (r0 * delta_x_layer + r1 * up.dx + float(x)) * P.scale + P.offset;
                ky = (r0 * delta_y_layer + r1 * up.dy + float(y)) * P.scale + P.offset;
            } else {
                const float r0 = scale * 3.0 - 2.0;
                const float r_1 = 1.0 - r0;
                kx = (r0 * delta_x_layer + r_1 * delta_x_below + float(x)) * P.scale + P.offset;
                ky = (r0 * delta_y_layer + r_1 * delta_y_below + float(y)) * P.scale + P.offset;
            }
        }
        scale *= P.scale;
        if (!(max > float(thresh))) return false;      // brisk.cpp:698
        kp.x = kx; kp.y = ky; kp.size = 12.0f * scale; kp.response = max; kp.layer = li;
        return true;
    }

    void run(bool jacobi_reverse, std::vector<Keypoint>& out) {
        out.clear();
        for (int li = 0; li < 6; li++) find_candidates(L[li]);
        // phase A: every candidate, any order
        for (int li = 0; li < 6; li++) {
            PLayer& P = L[li];
            for (size_t n = 0; n < P.cand.size(); n++) {
                const int x = P.cand[n] % P.w, y = P.cand[n] / P.w;
                const int m = max8(P, x, y);
                if (m == 1) after_max2d(li, x, y);
                else if (m == 2) P.status[P.cand[n]] = ST_PENDING;
            }
        }
        // phase B: tie candidates, layer by layer (the seen map of layer li is complete once layer li-1 is)
        for (int li = 0; li < 6; li++) {
            PLayer& P = L[li];
            std::vector<int> pending;
            for (size_t n = 0; n < P.cand.size(); n++)
                if (P.status[P.cand[n]] & ST_PENDING) pending.push_back(P.cand[n]);
            if (!jacobi_reverse) {
                for (size_t n = 0; n < pending.size(); n++) {
                    const int x = pending[n] % P.w, y = pending[n] / P.w;
                    P.status[pending[n]] = 0;
                    if (tie_break(P, x, y)) after_max2d(li, x, y);
                }
            } else {
                while (!pending.empty()) {
                    std::vector<int> ready, rest;
                    for (size_t n = pending.size(); n--;) {
                        const int x = pending[n] % P.w, y = pending[n] / P.w;
                        (earlier_pending_nearby(P, x, y) ? rest : ready).push_back(pending[n]);
                    }
                    // decisions of one round see only the statuses of the previous round
                    std::vector<char> pass(ready.size());
                    for (size_t n = 0; n < ready.size(); n++)
                        pass[n] = tie_break(P, ready[n] % P.w, ready[n] / P.w);
                    for (size_t n = 0; n < ready.size(); n++) {
                        P.status[ready[n]] = 0;
                        if (pass[n]) after_max2d(li, ready[n] % P.w, ready[n] / P.w);
                    }
                    std::reverse(rest.begin(), rest.end());
                    pending.swap(rest);
                }
            }
        }
        // phase C: survivors in layer-major raster order
        for (int li = 0; li < 6; li++) {
            PLayer& P = L[li];
            for (size_t n = 0; n < P.cand.size(); n++) {
                if (!(P.status[P.cand[n]] & ST_SURVIVE)) continue;
                Keypoint kp;
                if (emit(li, P.cand[n] % P.w, P.cand[n] / P.w, kp)) out.push_back(kp);
            }
        }
    }
};

}  // namespace

void brisk_port_detect_ex(const uint8_t* img, int w, int h, int thresh, bool jacobi_reverse,
                          std::vector<Keypoint>& out) {
    out.clear();
    // every layer must hold at least one detectable pixel row/column for the arithmetic above
    if (w < 48 || h < 48) return;
    Space* S = new Space();
    S->thresh = thresh;
    S->build(img, w, h);
    S->run(jacobi_reverse, out);
    delete S;
}

void brisk_port_detect(const uint8_t* img, int w, int h, std::vector<Keypoint>& out) {
    brisk_port_detect_ex(img, w, h, 30, false, out);
}

}  // namespace orc
