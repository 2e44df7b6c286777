// Device-side BRISK scale-space arithmetic: what BriskScaleSpace does per candidate in reference
// src/MoFREAK/brisk.cpp — getScoreMaxAbove/Below (:1106-1416), subpixel2D (:1535-1644), refine1D*
// (:1418-1533), refine3D (:937-1103), the float score sampler (:1711-1725) — on dense closed-form
// score maps instead of the reference's lazily memoised ones.
//
// The float results are part of the bit-exact contract (they become x, y, scale in the .mofreak rows and
// decide the FREAK scale index), so this file is compiled with -fmad=false and every expression keeps
// the reference's precision and association: float where it computes in float, double where a double
// literal promotes the expression.
#pragma once
#include "geom.h"

namespace mfk {

// One layer of one pair's score map, with the optional "seen" side-effect map of getAgastScore's
// write-back (only meaningful for reads issued by getScoreMaxAbove, see kernels_sparse.cu).
struct ScoreView {
    const uint8_t* S;
    uint8_t* seen;     // nullptr: pure read
    int w, h;

    // BriskLayer::getAgastScore(int, int, 1), brisk.cpp:1691-1700
    __device__ __forceinline__ int at(int x, int y) const {
        if (x < 3 || y < 3 || x >= w - 3 || y >= h - 3) return 0;
        const int i = y * w + x;
        if (seen) seen[i] = 1;
        return S[i];
    }
    // BriskLayer::getAgastScore(float, float, 1, scale <= 1), brisk.cpp:1711-1725
    __device__ __forceinline__ float atf(float xf, float yf) const {
        const int x = int(xf);
        const float rx1 = xf - float(x);
        const float rx = 1.0f - rx1;
        const int y = int(yf);
        const float ry1 = yf - float(y);
        const float ry = 1.0f - ry1;
        const int s00 = at(x, y), s10 = at(x + 1, y), s01 = at(x, y + 1), s11 = at(x + 1, y + 1);
        const float v = rx * ry * s00 + rx1 * ry * s10 + rx * ry1 * s01 + rx1 * ry1 * s11;
        return (float)(unsigned char)v;
    }
};

// brisk.cpp:1535-1644, including the delta_y = delta_x quirk of the out-of-range branch (:1627-1636)
__device__ inline float subpixel2D(const int s_0_0, const int s_0_1, const int s_0_2, const int s_1_0,
                                   const int s_1_1, const int s_1_2, const int s_2_0, const int s_2_1,
                                   const int s_2_2, float& delta_x, float& delta_y) {
    const int tmp1 = s_0_0 + s_0_2 - 2 * s_1_1 + s_2_0 + s_2_2;
    const int coeff1 = 3 * (tmp1 + s_0_1 - ((s_1_0 + s_1_2) << 1) + s_2_1);
    const int coeff2 = 3 * (tmp1 - ((s_0_1 + s_2_1) << 1) + s_1_0 + s_1_2);
    const int tmp2 = s_0_2 - s_2_0;
    const int tmp3 = (s_0_0 + tmp2 - s_2_2);
    const int tmp4 = tmp3 - 2 * tmp2;
    const int coeff3 = -3 * (tmp3 + s_0_1 - s_2_1);
    const int coeff4 = -3 * (tmp4 + s_1_0 - s_1_2);
    const int coeff5 = (s_0_0 - s_0_2 - s_2_0 + s_2_2) * 4;
    const int coeff6 = -(s_0_0 + s_0_2 - ((s_1_0 + s_0_1 + s_1_2 + s_2_1) << 1) - 5 * s_1_1 + s_2_0 + s_2_2) * 2;

    const int H_det = 4 * coeff1 * coeff2 - coeff5 * coeff5;
    if (H_det == 0) {
        delta_x = 0.0f;
        delta_y = 0.0f;
        return (float)((double)float(coeff6) / 18.0);
    }
    if (!(H_det > 0 && coeff1 < 0)) {
        int tmp_max = coeff3 + coeff4 + coeff5;
        delta_x = 1.0f; delta_y = 1.0f;
        int tmp = -coeff3 + coeff4 - coeff5;
        if (tmp > tmp_max) { tmp_max = tmp; delta_x = -1.0f; delta_y = 1.0f; }
        tmp = coeff3 - coeff4 - coeff5;
        if (tmp > tmp_max) { tmp_max = tmp; delta_x = 1.0f; delta_y = -1.0f; }
        tmp = -coeff3 - coeff4 + coeff5;
        if (tmp > tmp_max) { tmp_max = tmp; delta_x = -1.0f; delta_y = -1.0f; }
        return (float)((double)float(tmp_max + coeff1 + coeff2 + coeff6) / 18.0);
    }
    delta_x = float(2 * coeff2 * coeff3 - coeff4 * coeff5) / float(-H_det);
    delta_y = float(2 * coeff1 * coeff4 - coeff3 * coeff5) / float(-H_det);
    bool tx = false, tx_ = false, ty = false, ty_ = false;
    if (delta_x > 1.0f) tx = true;
    else if (delta_x < -1.0f) tx_ = true;
    if (delta_y > 1.0f) ty = true;
    if (delta_y < -1.0f) ty_ = true;

    if (tx || tx_ || ty || ty_) {
        float delta_x1 = 0.0f, delta_x2 = 0.0f, delta_y1 = 0.0f, delta_y2 = 0.0f;
        if (tx) {
            delta_x1 = 1.0f;
            delta_y1 = -float(coeff4 + coeff5) / float(2 * coeff2);
            if (delta_y1 > 1.0f) delta_y1 = 1.0f; else if (delta_y1 < -1.0f) delta_y1 = -1.0f;
        } else if (tx_) {
            delta_x1 = -1.0f;
            delta_y1 = -float(coeff4 - coeff5) / float(2 * coeff2);
            if (delta_y1 > 1.0f) delta_y1 = 1.0f; else if (delta_y1 < -1.0f) delta_y1 = -1.0f;
        }
        if (ty) {
            delta_y2 = 1.0f;
            delta_x2 = -float(coeff3 + coeff5) / float(2 * coeff1);
            if (delta_x2 > 1.0f) delta_x2 = 1.0f; else if (delta_x2 < -1.0f) delta_x2 = -1.0f;
        } else if (ty_) {
            delta_y2 = -1.0f;
            delta_x2 = -float(coeff3 - coeff5) / float(2 * coeff1);
            if (delta_x2 > 1.0f) delta_x2 = 1.0f; else if (delta_x2 < -1.0f) delta_x2 = -1.0f;
        }
        const float q1 = coeff1 * delta_x1 * delta_x1 + coeff2 * delta_y1 * delta_y1 + coeff3 * delta_x1 +
                         coeff4 * delta_y1 + coeff5 * delta_x1 * delta_y1 + coeff6;
        const float q2 = coeff1 * delta_x2 * delta_x2 + coeff2 * delta_y2 * delta_y2 + coeff3 * delta_x2 +
                         coeff4 * delta_y2 + coeff5 * delta_x2 * delta_y2 + coeff6;
        const float max1 = (float)((double)q1 / 18.0);
        const float max2 = (float)((double)q2 / 18.0);
        if (max1 > max2) { delta_x = delta_x1; delta_y = delta_x1; return max1; }
        delta_x = delta_x2; delta_y = delta_x2; return max2;
    }
    const float q = coeff1 * delta_x * delta_x + coeff2 * delta_y * delta_y + coeff3 * delta_x + coeff4 * delta_y +
                    coeff5 * delta_x * delta_y + coeff6;
    return (float)((double)q / 18.0);
}

// The three parabola fits over relative scale (brisk.cpp:1418-1533).  kind 0: refine1D (octave layers
// 2, 4), 1: refine1D_1 (intra layers), 2: refine1D_2 (layer 0).
__device__ inline float refine1D_fit(const int kind, const float s_05, const float s0, const float s05, float& max) {
    int a0, a1, a2, b0, b1, b2, c0, c1, c2;
    double lo, hi, den;
    if (kind == 0)      { a0 = 16; a1 = -24; a2 = 8;  b0 = -40; b1 = 54; b2 = -14; c0 = 24; c1 = -27; c2 = 6; lo = 0.75; hi = 1.5; den = 3072.0; }
    else if (kind == 1) { a0 = 9;  a1 = -18; a2 = 9;  b0 = -21; b1 = 36; b2 = -15; c0 = 12; c1 = -16; c2 = 6;
                          lo = 0.6666666666666666666666666667; hi = 1.3333333333333333333333333333; den = 2048.0; }
    else                { a0 = 2;  a1 = -4;  a2 = 2;  b0 = -5;  b1 = 8;  b2 = -3;  c0 = 3;  c1 = -3;  c2 = 1; lo = 0.7; hi = 1.5; den = 1024.0; }
    const int i_05 = int(1024.0 * (double)s_05 + 0.5);
    const int i0 = int(1024.0 * (double)s0 + 0.5);
    const int i05 = int(1024.0 * (double)s05 + 0.5);
    const int qa = a0 * i_05 + a1 * i0 + a2 * i05;
    if (qa >= 0) {
        if (s0 >= s_05 && s0 >= s05) { max = s0; return 1.0f; }
        if (s_05 >= s0 && s_05 >= s05) { max = s_05; return (float)lo; }
        if (s05 >= s0 && s05 >= s_05) { max = s05; return (float)hi; }
    }
    const int qb = b0 * i_05 + b1 * i0 + b2 * i05;
    float ret_val = -float(qb) / float(2 * qa);
    if ((double)ret_val < lo) ret_val = (float)lo;
    else if ((double)ret_val > hi) ret_val = (float)hi;
    const int qc = c0 * i_05 + c1 * i0 + c2 * i05;
    max = float(qc) + float(qa) * ret_val * ret_val + float(qb) * ret_val;
    max = (float)((double)max / den);
    return ret_val;
}

struct Scan { bool ismax; float max, dx, dy; };

// getScoreMaxAbove (below == false) / getScoreMaxBelow (below == true) of candidate (x_layer, y_layer)
// on layer `layer` over the neighbouring layer N.
__device__ inline Scan scan_neighbour(const ScoreView N, const bool below, const int layer, const int x_layer,
                                      const int y_layer, const int threshold) {
    Scan R; R.ismax = false; R.max = 0.0f; R.dx = 0.0f; R.dy = 0.0f;
    float x_1, x1, y_1, y1;
    if (!below) {
        if (layer % 2 == 0) {
            x_1 = (float)((double)float(4 * x_layer - 1 - 2) / 6.0); x1 = (float)((double)float(4 * x_layer - 1 + 2) / 6.0);
            y_1 = (float)((double)float(4 * y_layer - 1 - 2) / 6.0); y1 = (float)((double)float(4 * y_layer - 1 + 2) / 6.0);
        } else {
            x_1 = float(6 * x_layer - 1 - 3) / 8.0f; x1 = float(6 * x_layer - 1 + 3) / 8.0f;
            y_1 = float(6 * y_layer - 1 - 3) / 8.0f; y1 = float(6 * y_layer - 1 + 3) / 8.0f;
        }
    } else {
        if (layer % 2 == 0) {
            x_1 = (float)((double)float(8 * x_layer + 1 - 4) / 6.0); x1 = (float)((double)float(8 * x_layer + 1 + 4) / 6.0);
            y_1 = (float)((double)float(8 * y_layer + 1 - 4) / 6.0); y1 = (float)((double)float(8 * y_layer + 1 + 4) / 6.0);
        } else {
            x_1 = (float)((double)float(6 * x_layer + 1 - 3) / 4.0); x1 = (float)((double)float(6 * x_layer + 1 + 3) / 4.0);
            y_1 = (float)((double)float(6 * y_layer + 1 - 3) / 4.0); y1 = (float)((double)float(6 * y_layer + 1 + 3) / 4.0);
        }
    }
    const int xa = int(x_1 + 1), xb = int(x1);
    const int ya = int(y_1 + 1), yb = int(y1);
    int max_x = xa, max_y = ya;
    const float thr = (float)threshold;
    float tmp;
    float max = N.atf(x_1, y_1);
    // first (fractional) row: leave as soon as a sample beats the candidate
    if (max > thr) return R;
    for (int x = xa; x <= xb; x++) {
        tmp = N.atf(float(x), y_1);
        if (tmp > thr) return R;
        if (tmp > max) { max = tmp; max_x = x; }
    }
    tmp = N.atf(x1, y_1);
    if (tmp > thr) return R;
    if (tmp > max) { max = tmp; max_x = xb; }
    // integer rows
    for (int y = ya; y <= yb; y++) {
        tmp = N.atf(x_1, float(y));
        if (tmp > thr) return R;
        if (tmp > max) { max = tmp; max_x = xa; max_y = y; }
        for (int x = xa; x <= xb; x++) {
            tmp = (float)N.at(x, y);
            if (tmp > thr) return R;
            if (below && tmp == max) {
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
        if (tmp > thr) return R;
        if (tmp > max) { max = tmp; max_x = xb; max_y = y; }
    }
    // last (fractional) row: no early exit
    tmp = N.atf(x_1, y1);
    if (tmp > max) { max = tmp; max_x = xa; max_y = yb; }
    for (int x = xa; x <= xb; x++) {
        tmp = N.atf(float(x), y1);
        if (tmp > max) { max = tmp; max_x = x; max_y = yb; }
    }
    tmp = N.atf(x1, y1);
    if (tmp > max) { max = tmp; max_x = xb; max_y = yb; }

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
            dx = (float)(((double)real_x * 8.0 + 1.0) / 6.0 - (double)float(x_layer));
            dy = (float)(((double)real_y * 8.0 + 1.0) / 6.0 - (double)float(y_layer));
        }
    } else {
        if (layer % 2 == 0) {
            dx = (float)(((double)real_x * 6.0 + 1.0) / 8.0 - (double)float(x_layer));
            dy = (float)(((double)real_y * 6.0 + 1.0) / 8.0 - (double)float(y_layer));
        } else {
            dx = (float)(((double)real_x * 4.0 - 1.0) / 6.0 - (double)float(x_layer));
            dy = (float)(((double)real_y * 4.0 - 1.0) / 6.0 - (double)float(y_layer));
        }
    }
    if (dx > 1.0f) { dx = 1.0f; returnrefined = false; }
    if (dx < -1.0f) { dx = -1.0f; returnrefined = false; }
    if (dy > 1.0f) { dy = 1.0f; returnrefined = false; }
    if (dy < -1.0f) { dy = -1.0f; returnrefined = false; }
    R.ismax = true;
    R.dx = dx; R.dy = dy;
    R.max = returnrefined ? fmaxf(refined_max, max) : max;
    return R;
}

// AGAST-5/8 closed-form score with b0 = 0 (agast5_8_nms.cc:42; ring offsets agast5_8.h:65-75), 2-px border
// (BriskLayer::getAgastScore_5_8, brisk.cpp:1702-1709).  Same packed scheme as the 9/16 score: every ring
// pixel is the pair (v, 255 - v) in one word, so one 3-input packed minimum serves the brighter and the darker
// test at once.  (A scalar max(best, max(mn, -mx)) formulation is miscompiled by nvcc 12.9 for sm_100a: ptxas
// folds the negation into a VIMNMX3 operand and drops it; see DESIGN.md "toolchain notes".)
__device__ inline int agast58_score(const uint8_t* img, int w, int h, int x, int y) {
    if (x < 2 || y < 2 || x >= w - 2 || y >= h - 2) return 0;
    const uint8_t* p = img + y * w + x;
    const int c = *p;
    const int ofs[8] = {-1, -w - 1, -w, -w + 1, 1, w + 1, w, w - 1};
    uint32_t r[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { const uint32_t v = p[ofs[i]]; r[i] = v | ((255u - v) << 16); }
    uint32_t best = 0;
#pragma unroll
    for (int s = 0; s < 8; s++) {
        uint32_t m = __vimin3_u16x2(r[s], r[(s + 1) & 7], r[(s + 2) & 7]);
        m = __vimin3_u16x2(m, r[(s + 3) & 7], r[(s + 4) & 7]);
        best = __vimax3_u16x2(best, m, m);
    }
    const int brighter = (int)(best & 0xffffu) - c;
    const int darker = (int)(best >> 16) - (255 - c);
    const int s = (brighter > darker ? brighter : darker) - 1;
    return s < 0 ? 0 : (s > 254 ? 254 : s);
}

}  // namespace mfk
